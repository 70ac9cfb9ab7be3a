// pass_kernel.cu — CUDA wrapper of the fused tile pass (logic in pass_core.cuh).
// Bound: HBM (one read + one write of the state per launch, 16 B per amplitude).
#include "common.cuh"
#include "pass_core.cuh"
#include "../../include/tcb200.h"

namespace tcb {

constexpr int PASS_THREADS = 256;
constexpr int LOAD_UNROLL = 8;

template <int R>
__global__ void __launch_bounds__(PASS_THREADS, 2)
pass_kernel(const float2* src, float2* dst, int nbits, const int32_t* __restrict__ prog,
            int prog_words, const float2* __restrict__ gatebuf, long long gate_bstride,
            unsigned long long index_base, unsigned tiles_per_state) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;

  // ---- stage the program (T is needed to carve shared memory: read it from global) ----
  const int T = __ldg(prog + H_T);
  const int L = __ldg(prog + H_L);
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  uint64_t* hi_flat = reinterpret_cast<uint64_t*>(smem_raw + ((size_t)8 << T));
  int32_t* sprog = reinterpret_cast<int32_t*>(smem_raw + ((size_t)8 << T) + ((size_t)8 << (T - L)));
  for (int w = tid; w < prog_words; w += PASS_THREADS) sprog[w] = __ldg(prog + w);
  __syncthreads();
  const int32_t* hdr = sprog;

  // flat offsets of the high tile bits: hi_flat[h] for h = t >> L
  for (int h = tid; h < (1 << (T - L)); h += PASS_THREADS) hi_flat[h] = tile_to_flat(h << L, hdr);

  const unsigned tile_id = blockIdx.x % tiles_per_state;
  const unsigned batch = blockIdx.x / tiles_per_state;
  const uint64_t base = tile_base(tile_id, hdr);
  const float2* src_b = src + ((size_t)batch << nbits);
  float2* dst_b = dst + ((size_t)batch << nbits);
  const float2* gates = gatebuf + (size_t)batch * gate_bstride;
  __syncthreads();

  // ---- load: LDG.128 (two amplitudes), LOAD_UNROLL requests in flight per thread ----
  const int nvec = 1 << (T - 1);
  const int lowmask = (1 << L) - 1;
  for (int v0 = tid; v0 < nvec; v0 += PASS_THREADS * LOAD_UNROLL) {
    float4 x[LOAD_UNROLL];
#pragma unroll
    for (int u = 0; u < LOAD_UNROLL; ++u) {
      const int v = v0 + u * PASS_THREADS;
      if (v < nvec) {
        const int t = 2 * v;
        const uint64_t g = base | hi_flat[t >> L] | (uint64_t)(t & lowmask);
        x[u] = ldg_stream(reinterpret_cast<const float4*>(src_b + g));
      }
    }
#pragma unroll
    for (int u = 0; u < LOAD_UNROLL; ++u) {
      const int v = v0 + u * PASS_THREADS;
      if (v < nvec) {
        const int t = 2 * v;
        tile[swz(t)] = make_float2(x[u].x, x[u].y);
        tile[swz(t + 1)] = make_float2(x[u].z, x[u].w);
      }
    }
  }
  __syncthreads();

  // ---- sub-passes ----
  const int nsub = hdr[H_NSUB];
  const int32_t* sp = hdr + HDR_WORDS;
  const uint64_t cta_base = base | index_base;
  for (int s = 0; s < nsub; ++s) {
    if (sp[S_KIND] == SUB_REG) {
      const int ngroups = 1 << (T - R);
      for (int g = tid; g < ngroups; g += PASS_THREADS)
        run_reg_subpass<R>(tile, hdr, sp, gates, g, cta_base, hi_flat);
    } else {
      run_smem_dense(tile, hdr, sp, gates, tid, PASS_THREADS);
    }
    __syncthreads();
    sp += sp[S_WORDS];
  }

  // ---- store ----
  for (int v0 = tid; v0 < nvec; v0 += PASS_THREADS * LOAD_UNROLL) {
#pragma unroll
    for (int u = 0; u < LOAD_UNROLL; ++u) {
      const int v = v0 + u * PASS_THREADS;
      if (v < nvec) {
        const int t = 2 * v;
        const uint64_t g = base | hi_flat[t >> L] | (uint64_t)(t & lowmask);
        const float2 a = tile[swz(t)], b = tile[swz(t + 1)];
        stg_stream(reinterpret_cast<float4*>(dst_b + g), make_float4(a.x, a.y, b.x, b.y));
      }
    }
  }
}

static size_t pass_smem_bytes(int T, int L, int prog_words) {
  return ((size_t)8 << T) + ((size_t)8 << (T - L)) + (size_t)prog_words * 4;
}

int launch_pass(const void* src, void* dst, int nbits, int64_t batch, const int32_t* program,
                int32_t program_words, int tile_bits, int low_bits, const void* gatebuf,
                int64_t gate_batch_stride, uint64_t index_base, cudaStream_t stream) {
  TCB_REQUIRE(tile_bits >= PASS_R + 5 && tile_bits <= PASS_MAX_T && tile_bits <= nbits,
              "tcb_sv_run_pass: tile_bits=%d out of range [%d,%d] (nbits=%d)", tile_bits,
              PASS_R + 5, PASS_MAX_T, nbits);
  TCB_REQUIRE(low_bits >= 1 && low_bits <= tile_bits, "tcb_sv_run_pass: bad low_bits=%d", low_bits);
  TCB_REQUIRE(program_words >= HDR_WORDS && program_words <= PASS_MAX_WORDS,
              "tcb_sv_run_pass: program_words=%d out of range", program_words);
  TCB_REQUIRE(batch >= 1, "tcb_sv_run_pass: batch must be >= 1");
  const uint64_t tiles = 1ull << (nbits - tile_bits);
  const uint64_t grid = tiles * (uint64_t)batch;
  TCB_REQUIRE(grid < (1ull << 31), "tcb_sv_run_pass: grid too large");
  const size_t smem = pass_smem_bytes(tile_bits, low_bits, program_words);
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  TCB_REQUIRE(smem <= 100 * 1024, "tcb_sv_run_pass: shared memory %zu too large", smem);
  pass_kernel<PASS_R><<<(unsigned)grid, PASS_THREADS, smem, stream>>>(
      reinterpret_cast<const float2*>(src), reinterpret_cast<float2*>(dst), nbits, program,
      program_words, reinterpret_cast<const float2*>(gatebuf), (long long)gate_batch_stride,
      (unsigned long long)index_base, (unsigned)tiles);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tcb

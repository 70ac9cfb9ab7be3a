// pass_kernel.cu — CUDA shell of the fused tile pass (logic in pass_core.cuh).
//
// Bound: HBM (one read + one write of the state per launch, 16 B per amplitude).
// Structure: PERSISTENT CTAs with a two-stage shared-memory pipeline.
//   * grid = (#SMs x CTAs per SM); a CTA walks tiles  blockIdx.x, + gridDim.x, ...
//   * while tile k is being computed in buffer k&1, tile k+1 streams into the other buffer with
//     cp.async (LDGSTS: no registers are tied up, so the 2^R amplitudes per thread and the next
//     tile's 64 KB in flight coexist);  cp.async.wait_group + one barrier hands the buffer over.
//   * once per CTA: the program is staged in shared memory and the gate tensors the pass needs are
//     copied into a shared-memory pool; per tile, warps resolve the fill records (fused 1q
//     products, diagonal gates against this tile's constant bits) from the pool — no global
//     latency on the per-tile critical path.
//   * blockDim = 2^(T-R) threads: 256 for 2^13-amplitude tiles (1 CTA/SM, 2 x 64 KB buffers),
//     128 for 2^12 tiles (2 CTAs/SM).
#include "common.cuh"
#include "pass_core.cuh"
#include "../../include/tcb200.h"

namespace tcb {

#ifdef PASS_PROFILE
__device__ unsigned long long g_pass_prof[16];
#define PROF_DECL unsigned long long _pt = clock64();
#define PROF_MARK(slot)                                            \
  do {                                                             \
    if (threadIdx.x == 0) {                                        \
      const unsigned long long _now = clock64();                   \
      atomicAdd(&g_pass_prof[slot], _now - _pt);                   \
      _pt = _now;                                                  \
    }                                                              \
  } while (0)
#else
#define PROF_DECL
#define PROF_MARK(slot)
#endif

__device__ __forceinline__ void cp_async8(float2* smem_dst, const float2* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shuffle a whole Mat2 from lane (lane + d)
__device__ __forceinline__ Mat2 shfl_down_mat2(const Mat2& m, int d) {
  Mat2 r;
  r.a.x = __shfl_down_sync(0xffffffffu, m.a.x, d);
  r.a.y = __shfl_down_sync(0xffffffffu, m.a.y, d);
  r.b.x = __shfl_down_sync(0xffffffffu, m.b.x, d);
  r.b.y = __shfl_down_sync(0xffffffffu, m.b.y, d);
  r.c.x = __shfl_down_sync(0xffffffffu, m.c.x, d);
  r.c.y = __shfl_down_sync(0xffffffffu, m.c.y, d);
  r.d.x = __shfl_down_sync(0xffffffffu, m.d.x, d);
  r.d.y = __shfl_down_sync(0xffffffffu, m.d.y, d);
  return r;
}

// per-tile prologue: a warp owns a fill record, lane i loads source i from the shared-memory gate
// pool, the ordered product is a shuffle tree (after step k lane i holds sources [i, i + 2^k)).
__device__ __forceinline__ void prologue_fill(int32_t* sprog, int prog_words, const float2* pool,
                                              uint64_t cta_bits, int warp, int lane, int nwarps, int first,
                                              int last) {
  const int nfill = sprog[H_NFILL];
  const int32_t* filltab = sprog + prog_words - nfill;
  for (int r = first + warp; r < last; r += nwarps) {
    const int32_t* rec = sprog + filltab[r];
    const int kind = rec[1], count = rec[2];
    if (count == 1) {  // nothing to multiply
      if (lane == 0) store_fill_result(sprog, rec, load_fill_source(rec + 4, kind, pool, cta_bits));
      continue;
    }
    Mat2 acc = mat2_identity();
    for (int c0 = 0; c0 < count; c0 += 32) {  // (records longer than a warp: chunks, in order)
      Mat2 e = mat2_identity();
      if (c0 + lane < count) e = load_fill_source(rec + 4 + 4 * (c0 + lane), kind, pool, cta_bits);
      const int span = count - c0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        if (d < span) {  // warp-uniform: skip tree levels that only see identities
          const Mat2 later = shfl_down_mat2(e, d);
          e = mat2_mul(later, e);
        }
      }
      acc = mat2_mul(e, acc);  // lane 0 holds the chunk product
    }
    if (lane == 0) store_fill_result(sprog, rec, acc);
  }
}

struct PassArgs {
  const float2* src;
  float2* dst;
  const int32_t* prog;
  const float2* gatebuf;
  long long gate_bstride;
  unsigned long long index_base;
  unsigned long long total_tiles;  // tiles_per_state * batch
  int log_tiles_per_state;
  int nbits, prog_words;
};

// R register bits, LT = log2(threads per CTA) (512 threads for a 2^13 tile at R = 4); tile bits T = LT + R are compile-time, so every
// swizzle constant of the streaming phases is a literal.
template <int R, int LT>
__global__ void __launch_bounds__(1 << LT, 1) pass_kernel(const PassArgs A) {
  constexpr int T = LT + R;
  constexpr int NT = 1 << LT;
  constexpr int N_LD = (1 << T) / NT;        // 8-byte cp.async per thread
  constexpr int N_ST = (1 << (T - 1)) / NT;  // STG.128 per thread
  constexpr size_t TILE_BYTES = (size_t)8 << T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NWARPS = NT / 32;

  // ---- once per CTA: carve shared memory, stage the program, build hi_flat ----
  const int L = __ldg(A.prog + H_L);
  const int poolsize = __ldg(A.prog + H_POOLSIZE);
  uint64_t* hi_flat = reinterpret_cast<uint64_t*>(smem_raw + 2 * TILE_BYTES);
  float2* pool = reinterpret_cast<float2*>(smem_raw + 2 * TILE_BYTES + ((size_t)8 << (T - L)));
  int32_t* sprog = reinterpret_cast<int32_t*>(smem_raw + 2 * TILE_BYTES + ((size_t)8 << (T - L)) +
                                              (((size_t)poolsize * 8 + 15) & ~(size_t)15));
  for (int w = tid; w < A.prog_words; w += NT) sprog[w] = __ldg(A.prog + w);
  __syncthreads();
  const int32_t* hdr = sprog;
  for (int h = tid; h < (1 << (T - L)); h += NT) hi_flat[h] = tile_to_flat(h << L, hdr);
  const int npool = hdr[H_NPOOL];
  const int nfill = hdr[H_NFILL];
  const int nstatic = hdr[H_NFILL_STATIC];
  const int32_t* pooltab = sprog + A.prog_words - nfill - 3 * npool;

  // element mapping of the streaming phases
  //   load : t = tid + NT*u        (8 B per lane: a warp instruction covers 256 contiguous bytes)
  //   store: t = 2*tid + 2*NT*u    (two amplitudes -> one STG.128)
  // tid and NT*u have disjoint bits, so swizzle and hi_flat index split into thread + literal parts.
  const int lowmask = (1 << L) - 1;
  const int ld_s0 = swz(tid), st_s0 = swz(2 * tid);
  const unsigned long long tps_mask = (1ull << A.log_tiles_per_state) - 1ull;

  long long cur_batch = -1;
  auto issue_load = [&](unsigned long long tile_global, int which) {
    float2* buf = reinterpret_cast<float2*>(smem_raw + (size_t)which * TILE_BYTES);
    const unsigned long long b = tile_global >> A.log_tiles_per_state;
    const uint64_t base = tile_base(tile_global & tps_mask, hdr);
    const float2* src_b = A.src + ((size_t)b << A.nbits) + (base | (uint64_t)(tid & lowmask));
    const uint64_t* hf = hi_flat + (tid >> L);
    const int hstep = NT >> L;  // hi_flat entries per u step
    uint64_t off[N_LD];
#pragma unroll
    for (int u = 0; u < N_LD; ++u) off[u] = hf[u * hstep];
#pragma unroll
    for (int u = 0; u < N_LD; ++u) cp_async8(buf + (ld_s0 ^ swz(NT * u)), src_b + off[u]);
    cp_async_commit();
  };

  unsigned long long tg = blockIdx.x;
  __syncthreads();  // hi_flat ready
  if (tg < A.total_tiles) issue_load(tg, 0);

  for (int k = 0; tg < A.total_tiles; ++k, tg += gridDim.x) {
    float2* tile = reinterpret_cast<float2*>(smem_raw + (size_t)(k & 1) * TILE_BYTES);
    const long long b = (long long)(tg >> A.log_tiles_per_state);
    const uint64_t base = tile_base(tg & tps_mask, hdr);
    const uint64_t cta_bits = base | A.index_base;
    const float2* gates = A.gatebuf + (size_t)b * A.gate_bstride;
    PROF_DECL
    if (b != cur_batch) {  // (re)stage the gate pool of this batch element
      for (int e = 0; e < npool; ++e) {
        const int goff = pooltab[3 * e], cnt = pooltab[3 * e + 1], poff = pooltab[3 * e + 2];
        for (int i = tid; i < cnt; i += NT) pool[poff + i] = gates[goff + i];
      }
      cur_batch = b;
      __syncthreads();
      // records that do not depend on tile bits (fused 1q products, ...): once per batch element
      prologue_fill(sprog, A.prog_words, pool, 0, warp, lane, NWARPS, 0, nstatic);
    }
    // resolve the tile-dependent fill records (shared memory only)
    PROF_MARK(0);
    prologue_fill(sprog, A.prog_words, pool, cta_bits, warp, lane, NWARPS, nstatic, nfill);
    PROF_MARK(1);
    // tile k has landed; everybody is done with the other buffer (barrier at the end of k-1)
    cp_async_wait_all();
    __syncthreads();
    PROF_MARK(2);
    if (tg + gridDim.x < A.total_tiles) issue_load(tg + gridDim.x, (k + 1) & 1);
    PROF_MARK(3);

    // ---- sub-passes ----
    const int nsub = hdr[H_NSUB];
    const int32_t* sp = hdr + HDR_WORDS;
    for (int s = 0; s < nsub; ++s) {
      if (sp[S_KIND] == SUB_REG) {
        run_reg_subpass<R>(tile, hdr, sp, tid, cta_bits, hi_flat);  // 2^(T-R) groups == NT threads
      } else {
        run_smem_dense(tile, hdr, sp, gates, tid, NT);
      }
      __syncthreads();
      sp += sp[S_WORDS];
    }
    PROF_MARK(4);

    // ---- store ----
    {
      float2* dst_b = A.dst + ((size_t)b << A.nbits) + (base | (uint64_t)((2 * tid) & lowmask));
      const uint64_t* hf = hi_flat + ((2 * tid) >> L);
      const int hstep = (2 * NT) >> L;
      uint64_t off[N_ST];
      float2 x[N_ST], y[N_ST];
#pragma unroll
      for (int u = 0; u < N_ST; ++u) {
        const int sa = st_s0 ^ swz(2 * NT * u);
        off[u] = hf[u * hstep];
        x[u] = tile[sa];
        y[u] = tile[sa ^ 1];
      }
#pragma unroll
      for (int u = 0; u < N_ST; ++u)
        stg_stream(reinterpret_cast<float4*>(dst_b + off[u]), make_float4(x[u].x, x[u].y, y[u].x, y[u].y));
    }
    PROF_MARK(5);
    __syncthreads();  // the buffer and sprog may be overwritten from here on
    PROF_MARK(6);
  }
}

static size_t pass_smem_bytes(int T, int L, int prog_words, int poolsize) {
  return ((size_t)16 << T) + ((size_t)8 << (T - L)) + (((size_t)poolsize * 8 + 15) & ~(size_t)15) +
         (size_t)prog_words * 4;
}

int launch_pass(const void* src, void* dst, int nbits, int64_t batch, const int32_t* program,
                int32_t program_words, int tile_bits, int low_bits, const void* gatebuf,
                int64_t gate_batch_stride, uint64_t index_base, cudaStream_t stream) {
  TCB_REQUIRE(tile_bits >= PASS_R + 5 && tile_bits <= PASS_MAX_T && tile_bits <= nbits,
              "tcb_sv_run_pass: tile_bits=%d out of range [%d,%d] (nbits=%d)", tile_bits,
              PASS_R + 5, PASS_MAX_T, nbits);
  TCB_REQUIRE(low_bits >= 1 && low_bits <= 5, "tcb_sv_run_pass: bad low_bits=%d (1..5)", low_bits);
  TCB_REQUIRE(program_words >= HDR_WORDS && program_words <= PASS_MAX_WORDS,
              "tcb_sv_run_pass: program_words=%d out of range", program_words);
  TCB_REQUIRE(batch >= 1, "tcb_sv_run_pass: batch must be >= 1");
  const uint64_t tiles = 1ull << (nbits - tile_bits);
  const uint64_t total = tiles * (uint64_t)batch;
  const int lt = tile_bits - PASS_R;
  // worst-case pool (the exact size is in the program header, which lives on the device)
  const size_t smem = pass_smem_bytes(tile_bits, low_bits, program_words, PASS_MAX_POOL);
  TCB_REQUIRE(smem <= 200 * 1024, "tcb_sv_run_pass: shared memory %zu too large", smem);
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
  uint64_t grid = (uint64_t)sm_count() * (ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm));
  if (grid > total) grid = total;
  PassArgs a;
  a.src = reinterpret_cast<const float2*>(src);
  a.dst = reinterpret_cast<float2*>(dst);
  a.prog = program;
  a.gatebuf = reinterpret_cast<const float2*>(gatebuf);
  a.gate_bstride = (long long)gate_batch_stride;
  a.index_base = (unsigned long long)index_base;
  a.total_tiles = total;
  a.log_tiles_per_state = nbits - tile_bits;
  a.nbits = nbits;
  a.prog_words = program_words;
  switch (lt) {
    case 9: pass_kernel<PASS_R, 9><<<(unsigned)grid, 512, smem, stream>>>(a); break;
    case 8: pass_kernel<PASS_R, 8><<<(unsigned)grid, 256, smem, stream>>>(a); break;
    case 7: pass_kernel<PASS_R, 7><<<(unsigned)grid, 128, smem, stream>>>(a); break;
    case 6: pass_kernel<PASS_R, 6><<<(unsigned)grid, 64, smem, stream>>>(a); break;
    default: pass_kernel<PASS_R, 5><<<(unsigned)grid, 32, smem, stream>>>(a); break;
  }
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef PASS_PROFILE
extern "C" int tcb_debug_pass_prof(unsigned long long* out16_host, int reset) {
  cudaMemcpyFromSymbol(out16_host, g_pass_prof, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_pass_prof, z, sizeof(z));
  }
  return 0;
}
#endif

}  // namespace tcb

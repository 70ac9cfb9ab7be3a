// pass_kernel.cu — CUDA shell of the fused tile pass (logic in pass_core.cuh).
//
// Bound: HBM (one read + one write of the state per launch, 16 B per amplitude).
// Structure: PERSISTENT CTAs, several per SM, each streaming one tile at a time.
//   * grid = (#SMs x CTAs per SM); a CTA walks tiles  blockIdx.x, + gridDim.x, ...
//   * blockDim = 2^(T-R) threads, R = 5: 128 threads for 2^12-amplitude tiles, FOUR CTAs per SM
//     (4 x 32 KB tiles + side tables, 4 x 128 x 128 registers = the whole register file), 256 threads
//     and two CTAs per SM for 2^13 tiles.  The register sub-passes are FP32-pipe bound (a 2x2 complex
//     matvec per amplitude pair and gate), the streaming phases HBM bound; with four CTAs per SM in
//     different phases the warp schedulers overlap one CTA's load / store with another's arithmetic
//     (measured: a CTA pair with in-CTA double buffering left the FMA pipe 2/3 idle — too few warps).
//   * load: 16-byte cp.async (LDGSTS) straight into the swizzled tile, so no registers are tied up
//     while the tile is in flight; the per-tile fill records are resolved meanwhile.
//   * once per CTA: the program is staged in shared memory and the gate tensors the pass needs are
//     copied into a shared-memory pool; per tile, threads resolve the fill records (fused 1q
//     products, diagonal gates against this tile's constant bits) from the pool — no global
//     latency on the per-tile critical path.
#include <stdlib.h>

#include "common.cuh"
#ifdef PASS_PROFILE
// sub-phase stamps of warp 0 (LDS phase / rounds / STS phase of every register sub-pass), see tools/phase_prof.py
namespace tcb {
__device__ unsigned long long g_pass_prof[16];
}
#define TCB_SUBPROF_DECL unsigned long long _st = clock64();
#define TCB_SUBPROF(slot)                                                        \
  do {                                                                           \
    if (threadIdx.x == 0) {                                                      \
      const unsigned long long _n = clock64();                                   \
      atomicAdd(&::tcb::g_pass_prof[slot], _n - _st);                            \
      _st = _n;                                                                  \
    }                                                                            \
  } while (0)
#endif
#include "pass_core.cuh"
#include "../../include/tcb200.h"

namespace tcb {

#ifdef PASS_PROFILE
// timeline of the CTAs resident on SM 0: [cta slot][tile][phase stamp] (globaltimer ns)
__device__ unsigned long long g_trace[8][64][6];
__device__ int g_trace_slots;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#define TRACE_DECL int _slot = -1, _tk = 0; if (threadIdx.x == 0 && smid() == 0) _slot = atomicAdd(&g_trace_slots, 1);
#define TRACE(ph) do { if (_slot >= 0 && _slot < 8 && _tk < 64) g_trace[_slot][_tk][ph] = gtime(); } while (0)
#define TRACE_NEXT ++_tk
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
#define TRACE_DECL
#define TRACE(ph)
#define TRACE_NEXT
#endif

__device__ __forceinline__ void cp_async16(float2* smem_dst, const float2* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
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

// fill records [first, last) with ONE THREAD per record (short records: <= FILL_SHORT sources)
__device__ __forceinline__ void prologue_fill_threads(int32_t* sprog, int prog_words, const float2* pool,
                                                      uint64_t cta_bits, int tid, int nthreads, int first,
                                                      int last) {
  const int nfill = sprog[H_NFILL];
  const int32_t* filltab = sprog + prog_words - nfill;
  for (int r = first + tid; r < last; r += nthreads) run_fill_record(sprog, filltab[r], pool, cta_bits);
}

// fill records [first, last) with ONE WARP per record: lane i loads source i from the shared-memory
// gate pool, the ordered product is a shuffle tree (after step k lane i holds sources [i, i + 2^k)).
__device__ __forceinline__ void prologue_fill_warps(int32_t* sprog, int prog_words, const float2* pool,
                                                    uint64_t cta_bits, int warp, int lane, int nwarps,
                                                    int first, int last) {
  const int nfill = sprog[H_NFILL];
  const int32_t* filltab = sprog + prog_words - nfill;
  for (int r = first + warp; r < last; r += nwarps) {
    const int32_t* rec = sprog + filltab[r];
    const int kind = rec[1], count = rec[2];
    Mat2 acc = fill_identity(kind);
    for (int c0 = 0; c0 < count; c0 += 32) {  // (records longer than a warp: chunks, in order)
      Mat2 e = fill_identity(kind);
      if (c0 + lane < count) e = load_fill_source(rec + 4 + 4 * (c0 + lane), kind, pool, cta_bits);
      const int span = count - c0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        if (d < span) {  // warp-uniform: skip tree levels that only see identities
          const Mat2 later = shfl_down_mat2(e, d);
          e = fill_combine(kind, later, e);
        }
      }
      acc = fill_combine(kind, e, acc);  // lane 0 holds the chunk product
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
  int stagger_ns, n_sm;
};

// shared-memory carve-up (bytes, every region 16-byte aligned)
constexpr int GRP_STRIDE = 40;  // per sub-pass: 32 lane parts + up to 8 warp parts
struct PassSmem {
  size_t tile, hi_flat, grp, pool, prog, total;
};
__host__ __device__ inline PassSmem pass_smem_layout(int T, int L, int prog_words, int poolsize) {
  PassSmem s;
  s.tile = 0;
  s.hi_flat = (size_t)8 << T;
  s.grp = s.hi_flat + ((((size_t)4 << (T - L)) + 15) & ~(size_t)15);
  s.pool = s.grp + (size_t)PASS_MAX_SUB * GRP_STRIDE * 4;
  s.prog = s.pool + (((size_t)poolsize * 8 + 15) & ~(size_t)15);
  s.total = s.prog + (((size_t)prog_words * 4 + 15) & ~(size_t)15);
  return s;
}

#define PASS_KERNEL_NAME pass_kernel
#define PASS_EXTRA_PARAM
#include "pass_kernel_body.cuh"
#undef PASS_KERNEL_NAME
#undef PASS_EXTRA_PARAM

// source of the generating variant: vecs[p][x] = factor of flat bit position p (all total_bits of them: the local
// bits and, for a sharded state, the rank bits above them) of the initial product state
struct PassGen {
  const float2* vecs;
  int total_bits;
};
#define PASS_GENERATE
#define PASS_KERNEL_NAME pass_kernel_generate
#define PASS_EXTRA_PARAM , const PassGen GEN
#include "pass_kernel_body.cuh"
#undef PASS_KERNEL_NAME
#undef PASS_EXTRA_PARAM
#undef PASS_GENERATE

template <int LT>
static int launch_pass_lt(const PassArgs& a, uint64_t tiles_total, size_t smem, cudaStream_t stream) {
#ifdef TCB_PASS_MINB  // (experiments: another CTAs-per-SM / registers-per-thread trade for every tile size)
  constexpr int MINB = TCB_PASS_MINB;
#else
  constexpr int MINB = (512 >> LT) < 1 ? 1 : ((512 >> LT) > 8 ? 8 : (512 >> LT));  // <= 128 registers / thread
#endif
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel<PASS_R, LT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
    attr_set = true;
  }
  // persistent CTAs: as many per SM as shared memory and the register file (MINB) allow
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > MINB ? MINB : per_sm);
  uint64_t grid = (uint64_t)sm_count() * per_sm;
  if (grid > tiles_total) grid = tiles_total;
  pass_kernel<PASS_R, LT, MINB><<<(unsigned)grid, 1 << LT, smem, stream>>>(a);
  return 0;
}

int launch_pass(const void* src, void* dst, int nbits, int64_t batch, const int32_t* program,
                int32_t program_words, int tile_bits, int low_bits, int pool_elems, const void* gatebuf,
                int64_t gate_batch_stride, uint64_t index_base, cudaStream_t stream) {
  TCB_REQUIRE(tile_bits >= PASS_R + 5 && tile_bits <= PASS_MAX_T && tile_bits <= nbits,
              "tcb_sv_run_pass: tile_bits=%d out of range [%d,%d] (nbits=%d)", tile_bits,
              PASS_R + 5, PASS_MAX_T, nbits);
  TCB_REQUIRE(low_bits >= 1 && low_bits <= 5, "tcb_sv_run_pass: bad low_bits=%d (1..5)", low_bits);
  TCB_REQUIRE(program_words >= HDR_WORDS && program_words <= PASS_MAX_WORDS,
              "tcb_sv_run_pass: program_words=%d out of range", program_words);
  TCB_REQUIRE(pool_elems >= 0 && pool_elems <= PASS_MAX_POOL, "tcb_sv_run_pass: pool_elems=%d out of range",
              pool_elems);
  TCB_REQUIRE(batch >= 1, "tcb_sv_run_pass: batch must be >= 1");
  TCB_REQUIRE(nbits - low_bits <= 32, "tcb_sv_run_pass: nbits - low_bits must be <= 32");
  const uint64_t tiles = 1ull << (nbits - tile_bits);
  const uint64_t total = tiles * (uint64_t)batch;
  const int lt = tile_bits - PASS_R;
  const size_t smem = pass_smem_layout(tile_bits, low_bits, program_words, pool_elems).total;
  TCB_REQUIRE(smem <= 200 * 1024, "tcb_sv_run_pass: shared memory %zu too large", smem);
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
  {
    const char* e = getenv("TCB_STAGGER_NS");
    a.stagger_ns = e ? atoi(e) : 0;
    a.n_sm = sm_count();
  }
  int rc = 0;
  switch (lt) {
    case 8: rc = launch_pass_lt<8>(a, total, smem, stream); break;
    case 7: rc = launch_pass_lt<7>(a, total, smem, stream); break;
    case 6: rc = launch_pass_lt<6>(a, total, smem, stream); break;
    default: rc = launch_pass_lt<5>(a, total, smem, stream); break;
  }
  if (rc) return rc;
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// the FIRST pass of a circuit that starts from a product state  prod_p vecs[p][x_p]  (|0...0> included): the tile is
// generated in shared memory, `dst` is only written.  Replaces tcb_sv_init_product + the read half of the first
// pass.  Only the default tile size (T = 12) and batch 1; the caller falls back to init + pass otherwise.
int launch_pass_generate(void* dst, int nbits, const int32_t* program, int32_t program_words, int tile_bits,
                         int low_bits, int pool_elems, const void* gatebuf, uint64_t index_base, const void* vecs,
                         int total_bits, cudaStream_t stream) {
  constexpr int LT = 7;
  TCB_REQUIRE(tile_bits == PASS_R + LT && tile_bits <= nbits, "tcb_sv_run_pass_generate: tile_bits=%d (only %d)",
              tile_bits, PASS_R + LT);
  TCB_REQUIRE(low_bits >= 1 && low_bits <= 5, "tcb_sv_run_pass_generate: bad low_bits=%d (1..5)", low_bits);
  TCB_REQUIRE(program_words >= HDR_WORDS && program_words <= PASS_MAX_WORDS,
              "tcb_sv_run_pass_generate: program_words=%d out of range", program_words);
  TCB_REQUIRE(pool_elems >= 0 && pool_elems <= PASS_MAX_POOL, "tcb_sv_run_pass_generate: pool_elems=%d out of range",
              pool_elems);
  TCB_REQUIRE(nbits - low_bits <= 32, "tcb_sv_run_pass_generate: nbits - low_bits must be <= 32");
  TCB_REQUIRE(total_bits >= nbits && total_bits <= 48, "tcb_sv_run_pass_generate: total_bits=%d (nbits=%d)", total_bits,
              nbits);
  const uint64_t total = 1ull << (nbits - tile_bits);
  const size_t smem = pass_smem_layout(tile_bits, low_bits, program_words, pool_elems).total;
  TCB_REQUIRE(smem <= 200 * 1024, "tcb_sv_run_pass_generate: shared memory %zu too large", smem);
  PassArgs a;
  a.src = nullptr;
  a.dst = reinterpret_cast<float2*>(dst);
  a.prog = program;
  a.gatebuf = reinterpret_cast<const float2*>(gatebuf);
  a.gate_bstride = 0;
  a.index_base = (unsigned long long)index_base;
  a.total_tiles = total;
  a.log_tiles_per_state = nbits - tile_bits;
  a.nbits = nbits;
  a.prog_words = program_words;
  a.stagger_ns = 0;
  a.n_sm = sm_count();
  PassGen g;
  g.vecs = reinterpret_cast<const float2*>(vecs);
  g.total_bits = total_bits;
  constexpr int MINB = 4;
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(pass_kernel_generate<PASS_R, LT, MINB>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int per_sm = (int)((227 * 1024) / (smem + 9 * 1024 + 1024));  // (+ the static product tables)
  per_sm = per_sm < 1 ? 1 : (per_sm > MINB ? MINB : per_sm);
  uint64_t grid = (uint64_t)sm_count() * per_sm;
  if (grid > total) grid = total;
  pass_kernel_generate<PASS_R, LT, MINB><<<(unsigned)grid, 1 << LT, smem, stream>>>(a, g);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef PASS_PROFILE
extern "C" int tcb_debug_pass_trace(unsigned long long* out_host /* [8][64][6] */, int* nslots) {
  cudaMemcpyFromSymbol(out_host, g_trace, sizeof(unsigned long long) * 8 * 64 * 6);
  cudaMemcpyFromSymbol(nslots, g_trace_slots, sizeof(int));
  int z = 0;
  cudaMemcpyToSymbol(g_trace_slots, &z, sizeof(int));
  return 0;
}
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

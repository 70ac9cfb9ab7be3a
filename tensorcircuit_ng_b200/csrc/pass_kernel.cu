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

// R register bits, LT = log2(threads per CTA); tile bits T = LT + R are compile-time, so every
// swizzle constant of the streaming phases is a literal.
template <int R, int LT, int MINB>
__global__ void __launch_bounds__(1 << LT, MINB) pass_kernel(const PassArgs A) {
  constexpr int T = LT + R;
  constexpr int NT = 1 << LT;
  constexpr int N_IO = (1 << (T - 1)) / NT;  // 16-byte chunks per thread (load and store)
  constexpr int NWARPS = (NT + 31) / 32;
  static_assert(NWARPS <= 8, "grp table holds 8 warp parts");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- once per CTA: carve shared memory, stage the program, build the static tables ----
  const int L = __ldg(A.prog + H_L);
  const int poolsize = __ldg(A.prog + H_POOLSIZE);
  const PassSmem lay = pass_smem_layout(T, L, A.prog_words, poolsize);
  float2* tile = reinterpret_cast<float2*>(smem_raw + lay.tile);
  uint32_t* hi_flat = reinterpret_cast<uint32_t*>(smem_raw + lay.hi_flat);  // tile_to_flat(h << L) >> L
  int32_t* grp = reinterpret_cast<int32_t*>(smem_raw + lay.grp);
  float2* pool = reinterpret_cast<float2*>(smem_raw + lay.pool);
  int32_t* sprog = reinterpret_cast<int32_t*>(smem_raw + lay.prog);
  for (int w = tid; w < A.prog_words; w += NT) sprog[w] = __ldg(A.prog + w);
  __syncthreads();
  const int32_t* hdr = sprog;
  for (int h = tid; h < (1 << (T - L)); h += NT) hi_flat[h] = (uint32_t)(tile_to_flat(h << L, hdr) >> L);
  // tile number -> CTA-constant flat bits through 6-bit lookup tables (the bit-deposit loop over the 18+ non-tile
  // bits ran once per tile in every thread)
  __shared__ unsigned long long tb_lut[7][64];
  const int tb_chunks = (hdr[H_NNONTILE] + 5) / 6;
  for (int e = tid; e < tb_chunks * 64; e += NT) {
    const int c = e >> 6, v = e & 63;
    unsigned long long g = 0;
    for (int i = 0; i < 6; ++i)
      if (c * 6 + i < hdr[H_NNONTILE]) g |= (unsigned long long)((v >> i) & 1) << hdr[H_NONTILEPOS + c * 6 + i];
    tb_lut[c][v] = g;
  }
  const int npool = hdr[H_NPOOL];
  const int nfill = hdr[H_NFILL];
  const int nstatic = hdr[H_NFILL_STATIC];
  const int nshort_end = hdr[H_NFILL_SHORT_END];
  const int nsub = hdr[H_NSUB];
  const int32_t* pooltab = sprog + A.prog_words - nfill - 3 * npool;
  {
    // tile index of a thread = lane part | warp part, per register sub-pass
    const int32_t* sp = hdr + HDR_WORDS;
    for (int s = 0; s < nsub; ++s) {
      if (sp[S_KIND] == SUB_REG) {
        for (int i = tid; i < GRP_STRIDE; i += NT) {
          const int g = i < 32 ? i : ((i - 32) << 5);
          grp[s * GRP_STRIDE + i] = group_to_tile(g & (NT - 1), T, R, sp);
        }
      }
      sp += sp[S_WORDS];
    }
  }

  // element mapping of the streaming phases: chunk c = tid + NT*u holds amplitudes t = 2c, 2c+1
  // (16 bytes per lane: a warp instruction covers 512 contiguous bytes of the tile index space).
  // tid and NT*u have disjoint bits, so swizzle and hi_flat index split into thread + literal parts.
  const int lowmask = (1 << L) - 1;
  const int io_s0 = swz(2 * tid);
  const unsigned long long tps_mask = (1ull << A.log_tiles_per_state) - 1ull;
  const uint32_t* hf = hi_flat + ((2 * tid) >> L);
  const int hstep = (2 * NT) >> L;  // hi_flat entries per u step
  __syncthreads();                  // hi_flat and grp ready

  if (A.stagger_ns > 0) {  // experiment: de-phase the CTAs that share an SM
    const int phase = blockIdx.x / A.n_sm;
    for (int i = 0; i < phase; ++i) __nanosleep(A.stagger_ns);
  }
  TRACE_DECL
  long long cur_batch = -1;
  for (unsigned long long tg = blockIdx.x; tg < A.total_tiles; tg += gridDim.x) {
    const long long b = (long long)(tg >> A.log_tiles_per_state);
    uint64_t base = 0;  // = tile_base(tg & tps_mask, hdr)
    {
      const unsigned long long tnum = tg & tps_mask;
      for (int c = 0; c < tb_chunks; ++c) base |= tb_lut[c][(tnum >> (6 * c)) & 63ull];
    }
    const uint64_t cta_bits = base | A.index_base;
    const float2* gates = A.gatebuf + (size_t)b * A.gate_bstride;
    PROF_DECL
    TRACE(0);
    // ---- load: 16-byte cp.async straight into the swizzled tile (no registers held) ----
    {
      const float2* src_b = A.src + ((size_t)b << A.nbits) + (base | (uint64_t)((2 * tid) & lowmask));
      uint64_t off[N_IO];
#pragma unroll
      for (int u = 0; u < N_IO; ++u) off[u] = (uint64_t)hf[u * hstep] << L;
#pragma unroll
      for (int u = 0; u < N_IO; ++u) cp_async16(tile + (io_s0 ^ swz(2 * NT * u)), src_b + off[u]);
      cp_async_commit();
    }
    PROF_MARK(0);
    // ---- while the tile streams in: resolve the fill records (shared memory only) ----
    if (b != cur_batch) {  // (re)stage the gate pool of this batch element
      for (int e = 0; e < npool; ++e) {
        const int goff = pooltab[3 * e], cnt = pooltab[3 * e + 1], poff = pooltab[3 * e + 2];
        for (int i = tid; i < cnt; i += NT) pool[poff + i] = gates[goff + i];
      }
      cur_batch = b;
      __syncthreads();
      // records that do not depend on tile bits (fused 1q products, tables, ...): once per batch element
      prologue_fill_warps(sprog, A.prog_words, pool, 0, warp, lane, NWARPS, 0, nstatic);
    }
    prologue_fill_threads(sprog, A.prog_words, pool, cta_bits, tid, NT, nstatic, nshort_end);
    prologue_fill_warps(sprog, A.prog_words, pool, cta_bits, warp, lane, NWARPS, nshort_end, nfill);
    PROF_MARK(1);
    TRACE(1);
    cp_async_wait_all();
    __syncthreads();
    PROF_MARK(2);
    TRACE(2);

    // ---- sub-passes ----
    const int32_t* sp = hdr + HDR_WORDS;
    for (int s = 0; s < nsub; ++s) {
      if (sp[S_KIND] == SUB_REG) {
        const int tbase = grp[s * GRP_STRIDE + lane] | grp[s * GRP_STRIDE + 32 + warp];
        run_reg_subpass<R>(tile, sp, tbase, cta_bits);  // 2^(T-R) groups == NT threads
      } else {
        run_smem_dense(tile, hdr, sp, gates, tid, NT);
      }
#ifdef PASS_PROFILE
      const unsigned long long _bt = clock64();
#endif
      __syncthreads();
#ifdef PASS_PROFILE
      if (threadIdx.x == 0) atomicAdd(&g_pass_prof[11], clock64() - _bt);
#endif
      sp += sp[S_WORDS];
    }
    PROF_MARK(4);
    TRACE(3);

    // ---- store ----
    {
      float2* dst_b = A.dst + ((size_t)b << A.nbits) + (base | (uint64_t)((2 * tid) & lowmask));
      uint64_t off[N_IO];
      float4 v[N_IO];
#pragma unroll
      for (int u = 0; u < N_IO; ++u) {
        off[u] = (uint64_t)hf[u * hstep] << L;
        v[u] = *reinterpret_cast<const float4*>(tile + (io_s0 ^ swz(2 * NT * u)));
      }
#pragma unroll
      for (int u = 0; u < N_IO; ++u) stg_stream(reinterpret_cast<float4*>(dst_b + off[u]), v[u]);
    }
    PROF_MARK(5);
    TRACE(4);
    __syncthreads();  // the tile and sprog may be overwritten from here on
    PROF_MARK(6);
    TRACE(5);
    TRACE_NEXT;
  }
}

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

// tn_gemm_tc.cu — pairwise tensor-network contraction on the 5th-generation tensor cores.
//
//   C[b, m, n] (+)= sum_k A[b, m, k] * B[b, k, n]        complex64, every mode of extent 2,
//
// with m / n / k / b bit-deposited into the operands' flat addresses (no transposed copies: the
// permutation that `tensordot` / cotengra's contract_core pay for as a separate transpose kernel,
// tensorcircuit/cons.py:948 and tensorcircuit/experimental.py:1008, is folded into the gather).
// Replaces the reference's torch.tensordot -> cuBLAS cgemm + ATen permute kernels.
//
// Arithmetic: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM, 3xTF32 error compensation
// (x = hi + lo, x*y ~ hi*hi + hi*lo + lo*hi) so the result keeps FP32-level accuracy (|d| ~ 1e-6
// relative), and the complex product as four real GEMMs that share two accumulators:
//     Cr = Ar Br + (-Ai) Bi      (the minus sign is the instruction descriptor's a_negate bit)
//     Ci = Ar Bi +   Ai  Br
// => 12 MMAs of 128 x BN x 8 per 8 complex k.
//
// Structure (one CTA per SM, persistent over output tiles of 128 (m) x BN (n)):
//   warps 0-3  producers + epilogue: gather A / B from global memory with the bit-deposit
//              addressing, split every component into TF32 hi / lo, store the eight operand tiles
//              of a stage in the UMMA canonical K-major (no swizzle) core-matrix layout, publish
//              them to the async proxy (fence.proxy.async) and arrive on the stage's `full`
//              mbarrier; after the last k block: tcgen05.ld the two accumulators of their own TMEM
//              lane quarter and scatter C (thread i owns row i of the tile).
//   warp 4     TMEM allocation; one elected lane issues the MMAs and tcgen05.commit's each stage
//              back to the producers (`empty`) and the finished tile to the epilogue (`tmem_full`).
// Bound: HBM for the skinny steps of a contraction tree (K, N <~ 64), the TF32 pipe / 3 beyond.
#include "common.cuh"
#include "tn_common.cuh"
#include "../../include/tcb200.h"

namespace tcb {

namespace tc {

constexpr int BM = 128;       // rows of an output tile = TMEM lanes
constexpr int KB = 16;        // complex k per stage (two MMA k-steps of 8)
constexpr int STAGES = 3;
constexpr int N_PROD = 128;   // producer / epilogue threads (warps 0-3)
constexpr int N_THREADS = N_PROD + 32;
constexpr uint32_t LBO = 128;       // bytes between core matrices adjacent in K
constexpr uint32_t SBO = 4 * 128;   // bytes between 8-row groups (KB / 4 core matrices each)

__device__ __forceinline__ uint64_t deposit(uint64_t v, const int8_t* pos, int n) {
  uint64_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1ull) << pos[i];
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem],  kind::tf32, one CTA
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: start address, leading (K) and stride (M/N) byte offsets in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __host__ constexpr uint32_t make_idesc(int n, int a_neg) {
  return (1u << 4)      /* D = F32 */
         | (2u << 7)    /* A = TF32 */
         | (2u << 10)   /* B = TF32 */
         | ((uint32_t)a_neg << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// operand tiles of one stage: [Ar_hi, Ar_lo, Ai_hi, Ai_lo] (BM rows) then [Br_hi, Br_lo, Bi_hi, Bi_lo] (BN rows)
template <int BN>
struct Stage {
  static constexpr uint32_t A_TILE = BM * KB * 4, B_TILE = BN * KB * 4;
  static constexpr uint32_t BYTES = 4 * A_TILE + 4 * B_TILE;
};
__device__ __forceinline__ uint32_t tile_off(int row, int kgroup) {  // 16-byte slot of (row, 4 k's)
  return (uint32_t)(row >> 3) * SBO + (uint32_t)kgroup * LBO + (uint32_t)(row & 7) * 16;
}

template <int BN>
__global__ void __launch_bounds__(N_THREADS, 1)
gemm_tc_kernel(const float2* __restrict__ A, const float2* __restrict__ B, float2* C, ContractParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  using St = Stage<BN>;
  unsigned char* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * St::BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  uint64_t* offBn = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // [BN]
  uint64_t* offCn = offBn + BN;                                  // [BN]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;  // Cr: [0, BN), Ci: [BN, 2 BN)

  if (tid == N_PROD) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, N_PROD);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, N_PROD);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint64_t Mtot = 1ull << p.nm, Ntot = 1ull << p.nn, Ktot = 1ull << p.nk;
  const uint64_t tiles_m = (Mtot + BM - 1) / BM, tiles_n = (Ntot + BN - 1) / BN;
  const uint64_t tiles_per_batch = tiles_m * tiles_n;
  const uint64_t total_tiles = tiles_per_batch << p.nb;
  const uint32_t nkb = (uint32_t)((Ktot + KB - 1) / KB);

  uint32_t it = 0;      // k blocks processed by this CTA so far (stage ring position)
  uint32_t tcount = 0;  // tiles processed by this CTA so far
  if (warp < 4) {
    // ===================== producers + epilogue: thread `tid` owns tile row `tid` =====================
    for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint64_t bt = tile / tiles_per_batch, tr = tile % tiles_per_batch;
      const uint64_t m0 = (tr / tiles_n) * BM, n0 = (tr % tiles_n) * BN;
      const uint64_t a_b = deposit(bt, p.batch_a, p.nb), b_b = deposit(bt, p.batch_b, p.nb),
                     c_b = deposit(bt, p.batch_c, p.nb);
      const bool row_ok = m0 + tid < Mtot;
      const uint64_t offAm = a_b | deposit(m0 + tid, p.m_a, p.nm);
      const uint64_t offCm = c_b | deposit(m0 + tid, p.m_c, p.nm);
      // the previous tile's epilogue is done with the column tables (barrier among the 128 producers)
      asm volatile("bar.sync 1, %0;" ::"n"(N_PROD));
      if (tid < BN) {
        offBn[tid] = b_b | deposit(n0 + tid, p.n_b, p.nn);
        offCn[tid] = deposit(n0 + tid, p.n_c, p.nn);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(N_PROD));

      for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        unsigned char* st = stages + (size_t)s * St::BYTES;
        const uint64_t k0 = (uint64_t)kb * KB;
        // k offsets of this block (every thread needs all 16: computed redundantly per group of 4)
#pragma unroll
        for (int g = 0; g < KB / 4; ++g) {
          // ---- A: row tid, k = 4g .. 4g+3 ----
          float hr[4], lr[4], hi_[4], li[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t k = k0 + 4 * g + j;
            float2 v = make_float2(0.f, 0.f);
            if (row_ok && k < Ktot) v = A[offAm | deposit(k, p.k_a, p.nk)];
            if (p.conj_a) v.y = -v.y;
            hr[j] = tf32_round(v.x);
            lr[j] = tf32_round(v.x - hr[j]);
            hi_[j] = tf32_round(v.y);
            li[j] = tf32_round(v.y - hi_[j]);
          }
          const uint32_t o = tile_off(tid, g);
          *reinterpret_cast<float4*>(st + 0 * St::A_TILE + o) = make_float4(hr[0], hr[1], hr[2], hr[3]);
          *reinterpret_cast<float4*>(st + 1 * St::A_TILE + o) = make_float4(lr[0], lr[1], lr[2], lr[3]);
          *reinterpret_cast<float4*>(st + 2 * St::A_TILE + o) = make_float4(hi_[0], hi_[1], hi_[2], hi_[3]);
          *reinterpret_cast<float4*>(st + 3 * St::A_TILE + o) = make_float4(li[0], li[1], li[2], li[3]);
        }
        // ---- B: units (row n, k group g), BN * 4 of them over 128 threads ----
        for (int u = tid; u < BN * (KB / 4); u += N_PROD) {
          const int n = u % BN, g = u / BN;
          const bool col_ok = n0 + n < Ntot;
          const uint64_t ob = offBn[n];
          float hr[4], lr[4], hi_[4], li[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t k = k0 + 4 * g + j;
            float2 v = make_float2(0.f, 0.f);
            if (col_ok && k < Ktot) v = B[ob | deposit(k, p.k_b, p.nk)];
            if (p.conj_b) v.y = -v.y;
            hr[j] = tf32_round(v.x);
            lr[j] = tf32_round(v.x - hr[j]);
            hi_[j] = tf32_round(v.y);
            li[j] = tf32_round(v.y - hi_[j]);
          }
          const uint32_t o = 4 * St::A_TILE + tile_off(n, g);
          *reinterpret_cast<float4*>(st + 0 * St::B_TILE + o) = make_float4(hr[0], hr[1], hr[2], hr[3]);
          *reinterpret_cast<float4*>(st + 1 * St::B_TILE + o) = make_float4(lr[0], lr[1], lr[2], lr[3]);
          *reinterpret_cast<float4*>(st + 2 * St::B_TILE + o) = make_float4(hi_[0], hi_[1], hi_[2], hi_[3]);
          *reinterpret_cast<float4*>(st + 3 * St::B_TILE + o) = make_float4(li[0], li[1], li[2], li[3]);
        }
        fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(full + s);
      }

      // ---- epilogue: the accumulators of this tile ----
      mbar_wait(tmem_full, tcount & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 8) {
        float cr[8], ci[8];
        tmem_ld8(trow + c0, cr);
        tmem_ld8(trow + BN + c0, ci);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (n0 + c0 + j < Ntot) {
              const uint64_t addr = offCm | offCn[c0 + j];
              float2 v = make_float2(cr[j], ci[j]);
              if (p.accumulate) {
                const float2 old = C[addr];
                v.x += old.x;
                v.y += old.y;
              }
              C[addr] = v;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
    }
  } else {
    // ===================== MMA issuer (warp 4) =====================
    constexpr uint32_t ID_POS = make_idesc(BN, 0), ID_NEG = make_idesc(BN, 1);
    const uint32_t d_cr = tmem_base, d_ci = tmem_base + BN;
    for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      mbar_wait(tmem_empty, (tcount & 1) ^ 1);  // the epilogue has drained the accumulators
      tc_fence_after();
      for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stages + (size_t)s * St::BYTES);
          const uint32_t sb = sa + 4 * St::A_TILE;
#pragma unroll
          for (int j = 0; j < KB / 8; ++j) {
            const uint32_t ko = (uint32_t)j * 2 * LBO;  // two core matrices = 8 tf32 along K
            const uint64_t ar_h = make_desc(sa + 0 * St::A_TILE + ko), ar_l = make_desc(sa + 1 * St::A_TILE + ko);
            const uint64_t ai_h = make_desc(sa + 2 * St::A_TILE + ko), ai_l = make_desc(sa + 3 * St::A_TILE + ko);
            const uint64_t br_h = make_desc(sb + 0 * St::B_TILE + ko), br_l = make_desc(sb + 1 * St::B_TILE + ko);
            const uint64_t bi_h = make_desc(sb + 2 * St::B_TILE + ko), bi_l = make_desc(sb + 3 * St::B_TILE + ko);
            const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
            // Cr = Ar Br - Ai Bi   (corrections first, leading term last)
            tc_mma(d_cr, ar_l, br_h, ID_POS, acc);
            tc_mma(d_cr, ar_h, br_l, ID_POS, 1u);
            tc_mma(d_cr, ai_l, bi_h, ID_NEG, 1u);
            tc_mma(d_cr, ai_h, bi_l, ID_NEG, 1u);
            tc_mma(d_cr, ar_h, br_h, ID_POS, 1u);
            tc_mma(d_cr, ai_h, bi_h, ID_NEG, 1u);
            // Ci = Ar Bi + Ai Br
            tc_mma(d_ci, ar_l, bi_h, ID_POS, acc);
            tc_mma(d_ci, ar_h, bi_l, ID_POS, 1u);
            tc_mma(d_ci, ai_l, br_h, ID_POS, 1u);
            tc_mma(d_ci, ai_h, br_l, ID_POS, 1u);
            tc_mma(d_ci, ar_h, bi_h, ID_POS, 1u);
            tc_mma(d_ci, ai_h, br_h, ID_POS, 1u);
          }
          tc_commit(empty + s);                      // the stage is free once these MMAs have read it
          if (kb + 1 == nkb) tc_commit(tmem_full);   // ... and the tile is complete
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

template <int BN>
static size_t smem_bytes() {
  return (size_t)STAGES * Stage<BN>::BYTES + (2 * STAGES + 2) * 8 + 16 + (size_t)2 * BN * 8 + 128;
}

template <int BN>
static int launch(const float2* a, const float2* b, float2* c, const ContractParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = smem_bytes<BN>();
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const uint64_t Mtot = 1ull << p.nm, Ntot = 1ull << p.nn;
  const uint64_t tiles = (((Mtot + BM - 1) / BM) * ((Ntot + BN - 1) / BN)) << p.nb;
  uint64_t grid = tiles;
  const uint64_t cap = (uint64_t)sm_count();
  if (grid > cap) grid = cap;
  gemm_tc_kernel<BN><<<(unsigned)grid, N_THREADS, smem, stream>>>(a, b, c, p);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tc

// N tile: the smallest of 16 / 32 / 64 that covers N (M = 128 needs N % 16 == 0; more N tiles beyond 64)
int launch_contract_tc(const float2* a, const float2* b, float2* c, const ContractParams& p, cudaStream_t stream) {
  const uint64_t Ntot = 1ull << p.nn;
  if (Ntot <= 16) return tc::launch<16>(a, b, c, p, stream);
  if (Ntot <= 32) return tc::launch<32>(a, b, c, p, stream);
  return tc::launch<64>(a, b, c, p, stream);
}

}  // namespace tcb

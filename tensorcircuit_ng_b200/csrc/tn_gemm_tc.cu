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
//   warps 0-7   producers: gather A / B from global memory with the bit-deposit addressing (32-byte
//               vector loads when the two lowest k modes are the two lowest address bits), split
//               every component into TF32 hi / lo, store the eight operand tiles of a stage in the
//               UMMA canonical K-major (no swizzle) core-matrix layout, publish them to the async
//               proxy (fence.proxy.async) and arrive on the stage's `full` mbarrier.
//   warp 8      TMEM allocation; one elected lane issues the MMAs and tcgen05.commit's each stage
//               back to the producers (`empty`) and the finished tile to the epilogue (`tmem_full`).
//   warps 9-12  epilogue: tcgen05.ld the two accumulators of their TMEM lane quarter (warp % 4),
//               scatter C, hand the accumulator stage back (`tmem_empty`).
//   The accumulators are double buffered in TMEM (4 BN columns), so the producers already stream
//   tile t+1 while the tensor core works on t and the epilogue drains t-1: the skinny steps of a
//   contraction tree (one k block per tile) are latency-bound without that overlap.
// Bound: HBM for the skinny steps of a contraction tree (K, N <~ 64), the TF32 pipe / 3 beyond.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tn_common.cuh"
#include "../../include/tcb200.h"

namespace tcb {

namespace tc {

constexpr int BM = 128;       // rows of an output tile = TMEM lanes
constexpr int KB = 16;        // complex k per stage (two MMA k-steps of 8)
constexpr int STAGES = 3;
constexpr int N_PROD = 256;   // producer threads (warps 0-7)
constexpr int N_EPI = 128;    // epilogue threads (warps 9-12)
constexpr int N_THREADS = N_PROD + 32 + N_EPI;
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
// TMA engine, non-tensor form: one elected thread moves a contiguous block global -> shared and the bytes are
// accounted on the stage's mbarrier (SASS: UBLKCP + SYNCS.ARRIVE.TRANS64)
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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

// split four complex values into the hi / lo TF32 planes and store one 16-byte slot per plane
__device__ __forceinline__ void split_store(unsigned char* st, uint32_t plane_bytes, uint32_t o, const float2 (&v)[4]) {
  float hr[4], lr[4], hi_[4], li[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hr[j] = tf32_round(v[j].x);
    lr[j] = tf32_round(v[j].x - hr[j]);
    hi_[j] = tf32_round(v[j].y);
    li[j] = tf32_round(v[j].y - hi_[j]);
  }
  *reinterpret_cast<float4*>(st + 0 * plane_bytes + o) = make_float4(hr[0], hr[1], hr[2], hr[3]);
  *reinterpret_cast<float4*>(st + 1 * plane_bytes + o) = make_float4(lr[0], lr[1], lr[2], lr[3]);
  *reinterpret_cast<float4*>(st + 2 * plane_bytes + o) = make_float4(hi_[0], hi_[1], hi_[2], hi_[3]);
  *reinterpret_cast<float4*>(st + 3 * plane_bytes + o) = make_float4(li[0], li[1], li[2], li[3]);
}
// four consecutive k of one row: two 16-byte loads when they are contiguous in memory
__device__ __forceinline__ void load4(const float2* __restrict__ X, uint64_t base, const uint64_t* dlo, int g, bool vec,
                                      bool ok, uint64_t k0, uint64_t Ktot, bool conj, float2 (&v)[4]) {
  if (!ok) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = make_float2(0.f, 0.f);
    return;
  }
  if (vec) {
    const float4* q = reinterpret_cast<const float4*>(X + (base | dlo[4 * g]));
    const float4 x = q[0], y = q[1];
    v[0] = make_float2(x.x, x.y);
    v[1] = make_float2(x.z, x.w);
    v[2] = make_float2(y.x, y.y);
    v[3] = make_float2(y.z, y.w);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (k0 + 4 * g + j < Ktot) ? X[base | dlo[4 * g + j]] : make_float2(0.f, 0.f);
  }
  if (conj) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j].y = -v[j].y;
  }
}

// PACKED: A and B are pre-split operand images (pack_operand_kernel below): per (row tile, k block) the four TF32
// planes of a stage in the exact shared-memory layout, so the producer side is one thread issuing two bulk copies
// per stage — no gather, no conversions, no generic-proxy stores between the memory system and the tensor core.
template <int BN, bool PACKED>
__global__ void __launch_bounds__(N_THREADS, 1)
gemm_tc_kernel(const float2* __restrict__ A, const float2* __restrict__ B, float2* C, ContractParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  using St = Stage<BN>;
  unsigned char* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * St::BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* offBn = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // [BN]   (producers)
  uint64_t* offCn = offBn + BN;                                  // [BN]   (epilogue)
  uint64_t* dAlo = offCn + BN;                                   // [KB] deposit of the low 4 k bits into A
  uint64_t* dBlo = dAlo + KB;                                    // [KB] ... into B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t ACC_COLS = 2 * BN;        // one accumulator stage: Cr [0, BN), Ci [BN, 2 BN)
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  constexpr int MMA_WARP = N_PROD / 32;

  if (tid == N_PROD) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, PACKED ? 1 : N_PROD);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full + a, 1);
      mbar_init(tmem_empty + a, N_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  const int nklo = p.nk < 4 ? p.nk : 4;
  if (tid < KB) {
    dAlo[tid] = deposit((uint64_t)tid, p.k_a, nklo);
    dBlo[tid] = deposit((uint64_t)tid, p.k_b, nklo);
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
  const bool vecA = p.nk >= 2 && p.k_a[0] == 0 && p.k_a[1] == 1;
  const bool vecB = p.nk >= 2 && p.k_b[0] == 0 && p.k_b[1] == 1;
  const bool vecC = p.nn >= 1 && p.n_c[0] == 0;  // (N is then even, so a pair never straddles the edge)

  uint32_t it = 0;      // k blocks processed by this CTA so far (stage ring position)
  uint32_t tcount = 0;  // tiles processed by this CTA so far
  if (PACKED && warp < MMA_WARP) {
    // ===================== producer: bulk copies of pre-split stage images =====================
    if (tid == 0) {
      const unsigned char* PA = reinterpret_cast<const unsigned char*>(A);
      const unsigned char* PB = reinterpret_cast<const unsigned char*>(B);
      for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const uint64_t bt = tile / tiles_per_batch, tr = tile % tiles_per_batch;
        const uint64_t mt = tr / tiles_n, nt = tr % tiles_n;
        const unsigned char* a_src = PA + ((bt * tiles_m + mt) * nkb) * (uint64_t)(4 * St::A_TILE);
        const unsigned char* b_src = PB + ((bt * tiles_n + nt) * nkb) * (uint64_t)(4 * St::B_TILE);
        for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty + s, ph ^ 1);
          unsigned char* st = stages + (size_t)s * St::BYTES;
          mbar_expect_tx(full + s, St::BYTES);
          bulk_g2s(st, a_src + (uint64_t)kb * (4 * St::A_TILE), 4 * St::A_TILE, full + s);
          bulk_g2s(st + 4 * St::A_TILE, b_src + (uint64_t)kb * (4 * St::B_TILE), 4 * St::B_TILE, full + s);
        }
      }
    }
  } else if (warp < MMA_WARP) {
    // ===================== producers =====================
    // Work items = (tile, k block), flattened; the global loads of item w+1 are issued before item w
    // is split and stored, so every thread keeps two batches in flight (a skinny step has ONE k block
    // per tile: without the prefetch each tile pays a full DRAM latency).
    // A unit = (row, group of 4 k) of the 128 x 16 stage tile, two units per thread.  When the four lowest k modes
    // are the four lowest address bits of the operand (tnengine's layout planning makes the contracted modes of
    // every intermediate the lowest bits), the 16 k of a row are one contiguous 128-byte run and a warp takes one
    // 8-row group with all four k groups (lane = 8 * group + row): its load instruction touches 8 cache lines
    // instead of 32 (the gathers kept the L1 data pipe 45 % busy and the tensor pipe at 33 %), and every quarter
    // warp still writes 8 different rows of one core-matrix column, i.e. 8 distinct 16-byte bank groups.
    const bool contigA = p.nk >= 4 && p.k_a[0] == 0 && p.k_a[1] == 1 && p.k_a[2] == 2 && p.k_a[3] == 3;
    const bool contigB = p.nk >= 4 && p.k_b[0] == 0 && p.k_b[1] == 1 && p.k_b[2] == 2 && p.k_b[3] == 3;
    int rowA[2], grpA[2];
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      const int u = tid + N_PROD * gg;
      rowA[gg] = contigA ? (((u >> 5) << 3) | (u & 7)) : (tid & (BM - 1));
      grpA[gg] = contigA ? ((u >> 3) & 3) : (2 * (tid >> 7) + gg);
    }
    constexpr int UB = (BN * (KB / 4) + N_PROD - 1) / N_PROD;  // B units (row n, k group g) per thread
    int rowB[UB], grpB[UB];
#pragma unroll
    for (int i = 0; i < UB; ++i) {
      const int u = tid + i * N_PROD;
      rowB[i] = contigB ? (((u >> 5) << 3) | (u & 7)) : (u % BN);
      grpB[i] = contigB ? ((u >> 3) & 3) : (u / BN);
    }
    struct Batch {
      float2 va[2][4];
      float2 vb[UB][4];
    };
    uint64_t cur_tile = ~0ull, offAm[2] = {0, 0}, offBn_r[UB];
    bool row_ok[2] = {false, false}, col_ok[UB];
    auto issue = [&](uint64_t tile, uint32_t kb, Batch& r) {
      if (tile != cur_tile) {  // new tile: row / column bases
        cur_tile = tile;
        const uint64_t bt = tile / tiles_per_batch, tr = tile % tiles_per_batch;
        const uint64_t m0 = (tr / tiles_n) * BM, n0 = (tr % tiles_n) * BN;
        const uint64_t a_b = deposit(bt, p.batch_a, p.nb);
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          row_ok[gg] = m0 + rowA[gg] < Mtot;
          offAm[gg] = a_b | deposit(m0 + rowA[gg], p.m_a, p.nm);
        }
        const uint64_t b_b = deposit(bt, p.batch_b, p.nb);
#pragma unroll
        for (int i = 0; i < UB; ++i) {
          const int u = tid + i * N_PROD;
          col_ok[i] = u < BN * (KB / 4) && n0 + rowB[i] < Ntot;
          offBn_r[i] = b_b | deposit(n0 + rowB[i], p.n_b, p.nn);
        }
      }
      const uint64_t k0 = (uint64_t)kb * KB;
      // k = k0 + j with k0 a multiple of 16: deposit(k) = deposit(high bits) | deposit(j)
      const uint64_t hiA = p.nk > 4 ? deposit((uint64_t)kb, p.k_a + 4, p.nk - 4) : 0;
      const uint64_t hiB = p.nk > 4 ? deposit((uint64_t)kb, p.k_b + 4, p.nk - 4) : 0;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg)
        load4(A, offAm[gg] | hiA, dAlo, grpA[gg], vecA, row_ok[gg] && k0 + 4 * grpA[gg] < Ktot, k0, Ktot,
              p.conj_a != 0, r.va[gg]);
#pragma unroll
      for (int i = 0; i < UB; ++i)
        load4(B, offBn_r[i] | hiB, dBlo, grpB[i], vecB, col_ok[i] && k0 + 4 * grpB[i] < Ktot, k0, Ktot,
              p.conj_b != 0, r.vb[i]);
    };
    auto commit = [&](const Batch& r) {
      const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(empty + s, ph ^ 1);
      unsigned char* st = stages + (size_t)s * St::BYTES;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) split_store(st, St::A_TILE, tile_off(rowA[gg], grpA[gg]), r.va[gg]);
#pragma unroll
      for (int i = 0; i < UB; ++i) {
        const int u = tid + i * N_PROD;
        if (u < BN * (KB / 4)) split_store(st + 4 * St::A_TILE, St::B_TILE, tile_off(rowB[i], grpB[i]), r.vb[i]);
      }
      fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
      mbar_arrive(full + s);
      ++it;
    };
    Batch b0, b1;
    uint64_t tile = blockIdx.x;
    uint32_t kb = 0;
    bool have = tile < total_tiles;
    if (have) issue(tile, kb, b0);
    while (have) {
      // next item
      uint64_t ntile = tile;
      uint32_t nkb_i = kb + 1;
      if (nkb_i == nkb) {
        nkb_i = 0;
        ntile = tile + gridDim.x;
      }
      const bool more = ntile < total_tiles;
      if (more) issue(ntile, nkb_i, b1);
      commit(b0);
      if (!more) break;
      tile = ntile;
      kb = nkb_i + 1;
      if (kb == nkb) {
        kb = 0;
        tile = ntile + gridDim.x;
      }
      const bool more2 = tile < total_tiles;
      if (more2) issue(tile, kb, b0);
      commit(b1);
      have = more2;
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    constexpr uint32_t ID_POS = make_idesc(BN, 0), ID_NEG = make_idesc(BN, 1);
    for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
      const uint32_t d_cr = tmem_base + a * ACC_COLS, d_ci = d_cr + BN;
      mbar_wait(tmem_empty + a, aph ^ 1);  // the epilogue has drained this accumulator stage
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
          tc_commit(empty + s);                          // the stage is free once these MMAs have read it
          if (kb + 1 == nkb) tc_commit(tmem_full + a);   // ... and the tile is complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: row = TMEM lane = 32 (warp % 4) + lane =====================
    const int row = (warp & 3) * 32 + lane;
    const int etid = tid - (N_PROD + 32);
    for (uint64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
      const uint64_t bt = tile / tiles_per_batch, tr = tile % tiles_per_batch;
      const uint64_t m0 = (tr / tiles_n) * BM, n0 = (tr % tiles_n) * BN;
      const uint64_t c_b = deposit(bt, p.batch_c, p.nb);
      const bool row_ok = m0 + row < Mtot;
      const uint64_t offCm = c_b | deposit(m0 + row, p.m_c, p.nm);
      asm volatile("bar.sync 2, %0;" ::"n"(N_EPI));
      if (etid < BN) offCn[etid] = deposit(n0 + etid, p.n_c, p.nn);
      asm volatile("bar.sync 2, %0;" ::"n"(N_EPI));
      mbar_wait(tmem_full + a, aph);
      tc_fence_after();
      const uint32_t trow = tmem_base + a * ACC_COLS + ((uint32_t)((warp & 3) * 32) << 16);
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
              if (vecC && !(j & 1)) {  // n, n+1 adjacent in C: one 16-byte store for the pair
                float2 w = make_float2(cr[j + 1], ci[j + 1]);
                if (p.accumulate) {
                  const float2 old = C[addr + 1];
                  w.x += old.x;
                  w.y += old.y;
                }
                *reinterpret_cast<float4*>(C + addr) = make_float4(v.x, v.y, w.x, w.y);
              } else if (!vecC) {
                C[addr] = v;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty + a);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// Pre-split operand image for the PACKED kernel.  X is A (rows = m) or B (rows = n) in its bit-deposited layout;
// the image is [batch][row tile of ROWS][k block of KB] x 4 planes (re hi, re lo, im hi, im lo) x ROWS x KB TF32
// values in the UMMA canonical K-major core-matrix layout.  One CTA per stage image; HBM bound (8 B read, 16 B
// written per element).  Rows / k beyond the edge are zero.
template <int ROWS>
__global__ void __launch_bounds__(256) pack_operand_kernel(const float2* __restrict__ X, unsigned char* __restrict__ out,
                                                           int nb, int nrow, int nk, const int8_t* __restrict__ pos,
                                                           int conj) {
  // pos: [32] batch, [32] row, [32] k bit positions
  __shared__ int8_t sp[96];
  __shared__ uint64_t dlo[KB];
  if (threadIdx.x < 96) sp[threadIdx.x] = pos[threadIdx.x];
  __syncthreads();
  const int nklo = nk < 4 ? nk : 4;
  if (threadIdx.x < KB) dlo[threadIdx.x] = deposit((uint64_t)threadIdx.x, sp + 64, nklo);
  __syncthreads();
  const uint64_t Rtot = 1ull << nrow, Ktot = 1ull << nk;
  const uint64_t tiles_r = (Rtot + ROWS - 1) / ROWS;
  const uint32_t nkb = (uint32_t)((Ktot + KB - 1) / KB);
  const uint64_t total = (tiles_r << nb) * nkb;
  const bool vec = nk >= 2 && sp[64] == 0 && sp[65] == 1;
  constexpr uint32_t TILE = ROWS * KB * 4;
  constexpr int UNITS = ROWS * (KB / 4);
  for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const uint32_t kb = (uint32_t)(w % nkb);
    const uint64_t rt = (w / nkb) % tiles_r, bt = (w / nkb) / tiles_r;
    const uint64_t k0 = (uint64_t)kb * KB;
    const uint64_t hi = (nk > 4 ? deposit((uint64_t)kb, sp + 68, nk - 4) : 0) | deposit(bt, sp, nb);
    unsigned char* img = out + w * (uint64_t)(4 * TILE);
    for (int u = threadIdx.x; u < UNITS; u += 256) {
      const int row = ((u >> 5) << 3) | (u & 7), g = (u >> 3) & 3;  // a warp = one 8-row group x 4 k groups
      const uint64_t r = rt * ROWS + row;
      float2 v[4];
      load4(X, hi | deposit(r, sp + 32, nrow), dlo, g, vec, r < Rtot && k0 + 4 * g < Ktot, k0, Ktot, conj != 0, v);
      split_store(img, TILE, tile_off(row, g), v);
    }
  }
}

template <int BN>
static size_t smem_bytes() {
  return (size_t)STAGES * Stage<BN>::BYTES + (2 * STAGES + 4) * 8 + 16 + (size_t)2 * BN * 8 + (size_t)2 * KB * 8 + 128;
}

template <int BN, bool PACKED>
static int launch(const float2* a, const float2* b, float2* c, const ContractParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  const size_t smem = smem_bytes<BN>();
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    attr_set = true;
  }
  const uint64_t Mtot = 1ull << p.nm, Ntot = 1ull << p.nn;
  const uint64_t tiles = (((Mtot + BM - 1) / BM) * ((Ntot + BN - 1) / BN)) << p.nb;
  uint64_t grid = tiles;
  const uint64_t cap = (uint64_t)sm_count();
  if (grid > cap) grid = cap;
  gemm_tc_kernel<BN, PACKED><<<(unsigned)grid, N_THREADS, smem, stream>>>(a, b, c, p);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// pre-split both operands into a stream-ordered scratch allocation, then run the bulk-copy fed kernel
static int launch_packed(const float2* a, const float2* b, float2* c, const ContractParams& p, cudaStream_t stream) {
  constexpr int BN = 128;
  const uint64_t Mtot = 1ull << p.nm, Ntot = 1ull << p.nn, Ktot = 1ull << p.nk;
  const uint64_t tiles_m = (Mtot + BM - 1) / BM, tiles_n = (Ntot + BN - 1) / BN, nkb = (Ktot + KB - 1) / KB;
  const size_t a_bytes = (size_t)((tiles_m << p.nb) * nkb) * 4 * Stage<BN>::A_TILE;
  const size_t b_bytes = (size_t)((tiles_n << p.nb) * nkb) * 4 * Stage<BN>::B_TILE;
  static bool pool_set = false;
  if (!pool_set) {  // keep freed scratch in the pool: the next contraction of the tree reuses it
    int dev = 0;
    cudaMemPool_t pool;
    TCB_CHECK_CUDA(cudaGetDevice(&dev));
    TCB_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = ~0ull;
    TCB_CHECK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_set = true;
  }
  unsigned char* scratch = nullptr;
  int8_t* d_pos = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), a_bytes + b_bytes + 256, stream) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;  // no room for the operand images: the caller falls back to the gathering kernel
  }
  d_pos = reinterpret_cast<int8_t*>(scratch + a_bytes + b_bytes);
  int8_t h_pos[192];
  memcpy(h_pos, p.batch_a, 32);
  memcpy(h_pos + 32, p.m_a, 32);
  memcpy(h_pos + 64, p.k_a, 32);
  memcpy(h_pos + 96, p.batch_b, 32);
  memcpy(h_pos + 128, p.n_b, 32);
  memcpy(h_pos + 160, p.k_b, 32);
  TCB_CHECK_CUDA(cudaMemcpyAsync(d_pos, h_pos, 192, cudaMemcpyHostToDevice, stream));
  const unsigned cap = (unsigned)sm_count() * 8;
  const uint64_t wa = (tiles_m << p.nb) * nkb, wb = (tiles_n << p.nb) * nkb;
  pack_operand_kernel<BM><<<(unsigned)(wa < cap ? wa : cap), 256, 0, stream>>>(a, scratch, p.nb, p.nm, p.nk, d_pos,
                                                                              p.conj_a);
  pack_operand_kernel<BN><<<(unsigned)(wb < cap ? wb : cap), 256, 0, stream>>>(b, scratch + a_bytes, p.nb, p.nn, p.nk,
                                                                              d_pos + 96, p.conj_b);
  TCB_CHECK_CUDA(cudaGetLastError());
  const int rc = launch<BN, true>(reinterpret_cast<const float2*>(scratch),
                                  reinterpret_cast<const float2*>(scratch + a_bytes), c, p, stream);
  TCB_CHECK_CUDA(cudaFreeAsync(scratch, stream));
  return rc;
}

}  // namespace tc

// N tile: the smallest of 16 / 32 / 64 / 128 that covers N (M = 128 needs N % 16 == 0; more N tiles
// beyond 128: every extra column of a tile amortises the hi / lo split of the A rows)
int launch_contract_tc(const float2* a, const float2* b, float2* c, const ContractParams& p, cudaStream_t stream) {
  const uint64_t Ntot = 1ull << p.nn;
  // fat steps (every operand element is reused >= 128 times): split the operands once, feed the tensor core by
  // bulk copies.  TCB_TN_PACKED=0 keeps the gathering kernel (parity tests run both).
  static const int packed_mode = [] {
    const char* e = getenv("TCB_TN_PACKED");
    return e ? atoi(e) : 1;
  }();
  if (packed_mode && p.nn >= 7 && p.nm >= 7 && p.nk >= 6) {
    const int rc = tc::launch_packed(a, b, c, p, stream);
    if (rc >= 0) return rc;
  }
  if (Ntot <= 16) return tc::launch<16, false>(a, b, c, p, stream);
  if (Ntot <= 32) return tc::launch<32, false>(a, b, c, p, stream);
  if (Ntot <= 64) return tc::launch<64, false>(a, b, c, p, stream);
  return tc::launch<128, false>(a, b, c, p, stream);
}

}  // namespace tcb

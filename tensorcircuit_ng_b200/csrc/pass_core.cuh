// pass_core.cuh — the fused statevector "tile pass": one HBM pass applies a whole
// program of gates.  Replaces the per-gate tensordot + transpose loop of the
// reference (tensorcircuit/cons.py:937-953) for circuit-shaped networks.
//
// Execution model (B200):
//   * the state (2^n complex64) is cut into tiles of 2^T amplitudes (T <= 13, 64 KB):
//     the T "tile bits" are the L lowest index bits (coalesced 128 B+ runs) plus T-L
//     arbitrary higher bits chosen by the host planner, so a tile sees every
//     combination of the qubits the program acts on;  all other index bits are
//     constant per CTA.
//   * one CTA = one tile.  load (LDG.128, coalesced) -> shared memory (XOR-swizzled
//     so that any choice of register bits is conflict free) -> a sequence of
//     *register sub-passes* -> store.
//   * a register sub-pass: every thread pulls 2^R amplitudes (R "register bits" of
//     the tile) into registers and applies every op of the sub-pass with compile-time
//     register indices (no local memory): dense 1q, controlled-1q, and diagonal 1q/2q
//     gates.  Diagonal gates may touch ANY qubit of the register (their non-register
//     bits are thread constants) and are accumulated as separable factors, applied
//     lazily — so e.g. a whole QAOA cost layer (45 ZZ gates) costs ~1 multiply per
//     amplitude per register bit.
//   * k-qubit dense gates (k = 2..5) run as shared-memory sub-passes inside the same
//     tile pass.
//
// The code below is host/device neutral (TCB_DEV): the CUDA kernel in pass_kernel.cu
// and the CPU logic emulator in tests/emu/ (test infrastructure, never loaded by the
// package) both include it, so index arithmetic is validated without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TCB_DEV __device__ __forceinline__
#define TCB_UNROLL _Pragma("unroll")
#define TCB_NOUNROLL _Pragma("unroll 1")
#else
#define TCB_DEV inline
#define TCB_UNROLL
#define TCB_NOUNROLL
struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
static inline float2 make_float2(float a, float b) {
  float2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif

namespace tcb {

// ---- program layout (int32 words) -------------------------------------------
constexpr int PASS_MAGIC = 0x7CB20001;
constexpr int PASS_MAX_T = 13;
constexpr int PASS_R = 5;  // register bits per sub-pass of the compiled kernel
constexpr int PASS_MAX_WORDS = 3072;
// header
constexpr int H_MAGIC = 0, H_T = 1, H_L = 2, H_NSUB = 3, H_WORDS = 4, H_NNONTILE = 5, H_R = 6;
constexpr int H_TILEPOS = 8;      // [16] flat-index bit position of tile bit t
constexpr int H_NONTILEPOS = 24;  // [56] flat-index bit positions of the non-tile bits, ascending
constexpr int HDR_WORDS = 80;
// sub-pass header
constexpr int S_NOPS = 0, S_KIND = 1, S_REGBITS = 2 /*[8]*/, S_GRPBITS = 10 /*[12]*/, S_WORDS = 22;
constexpr int SUB_HDR_WORDS = 24;
constexpr int SUB_REG = 0, SUB_SMEM_DENSE = 1;
// ops
constexpr int OP_WORDS = 8;
constexpr int O_CODE = 0, O_A = 1, O_B = 2, O_MAT = 3, O_AUX0 = 4, O_AUX1 = 5, O_AUX2 = 6, O_AUX3 = 7;
enum OpCode : int {
  OP_1Q = 1,     // A = reg index j;            MAT -> 2x2 row-major, AUX1 = row stride
  OP_C1Q = 2,    // A = ctrl qref, B = reg j (target), AUX0 = polarity bits, AUX1 = row stride,
                 // AUX2 = second ctrl qref or -1;   MAT -> top-left of the active 2x2 block
  OP_DIAG1 = 3,  // A = qref;  MAT -> d0, d1 at MAT + x*AUX1
  OP_DIAG2 = 4,  // A,B = qrefs; MAT -> d[xa][xb] at MAT + (2*xa+xb)*AUX1
  OP_DENSE = 5,  // (SUB_SMEM_DENSE only) AUX0 = k, tile-bit indices of gate qubits in
                 // words A,B,AUX1,AUX2,AUX3 ; MAT -> 2^k x 2^k row-major
};
// qref: < 32 -> register-bit index j of the sub-pass;  >= 32 -> flat-index bit (value - 32)
constexpr int QREF_BIT = 32;

// ---- complex helpers ------------------------------------------------------------
TCB_DEV float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
TCB_DEV float2 cfma(float2 a, float2 b, float2 c) {  // a*b + c
  return make_float2(a.x * b.x - a.y * b.y + c.x, a.x * b.y + a.y * b.x + c.y);
}
TCB_DEV float2 csel(bool p, float2 a, float2 b) { return p ? a : b; }

// XOR-fold swizzle of the low 4 bits (8-byte words -> 16 bank pairs per half warp)
TCB_DEV int swz(int t) { return (t & ~15) | ((t ^ (t >> 4) ^ (t >> 8) ^ (t >> 12)) & 15); }

// flat index of tile element t (without the CTA-constant part)
TCB_DEV uint64_t tile_to_flat(int t, const int32_t* hdr) {
  const int T = hdr[H_T], L = hdr[H_L];
  uint64_t g = (uint64_t)(t & ((1 << L) - 1));
  for (int i = L; i < T; ++i) g |= (uint64_t)((t >> i) & 1) << hdr[H_TILEPOS + i];
  return g;
}

// CTA-constant part of the flat index for tile number `tile`
TCB_DEV uint64_t tile_base(uint64_t tile, const int32_t* hdr) {
  const int nn = hdr[H_NNONTILE];
  uint64_t g = 0;
  for (int i = 0; i < nn; ++i) g |= ((tile >> i) & 1ull) << hdr[H_NONTILEPOS + i];
  return g;
}

// ---- register sub-pass -------------------------------------------------------------
// 2x2 complex matrix on register bit J of the 2^R amplitudes held by this thread.  Controls that
// live in registers are given as (cmask, cwant) over the register index (0,0 = uncontrolled); the
// test is on compile-time indices, so uncontrolled gates pay nothing for it after unrolling only
// when the compiler can see cmask == 0 — hence the two instantiations below.
template <int R, int J, bool CTRL>
TCB_DEV void apply_1q(float2 (&a)[1 << R], float2 m00, float2 m01, float2 m10, float2 m11,
                      int cmask, int cwant) {
  TCB_UNROLL
  for (int p = 0; p < (1 << (R - 1)); ++p) {
    const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
    const int i1 = i0 | (1 << J);
    const float2 x = a[i0], y = a[i1];
    const float2 b0 = cfma(m01, y, cmul(m00, x));
    const float2 b1 = cfma(m11, y, cmul(m10, x));
    if (CTRL) {
      const bool on = ((i0 & cmask) == cwant);
      a[i0] = csel(on, b0, x);
      a[i1] = csel(on, b1, y);
    } else {
      a[i0] = b0;
      a[i1] = b1;
    }
  }
}

// multiply by a per-bit diagonal factor (u0 for bit J = 0, u1 for bit J = 1)
template <int R, int J>
TCB_DEV void apply_bitdiag(float2 (&a)[1 << R], float2 u0, float2 u1) {
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) a[i] = cmul(a[i], ((i >> J) & 1) ? u1 : u0);
}

// diagonal on two register bits J (static, first gate qubit) and k (runtime): d[xj][xk]
template <int R, int J>
TCB_DEV void apply_pairdiag(float2 (&a)[1 << R], int k, float2 d00, float2 d01, float2 d10,
                            float2 d11) {
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    const bool xk = (i >> k) & 1;
    const float2 f = ((i >> J) & 1) ? csel(xk, d11, d10) : csel(xk, d01, d00);
    a[i] = cmul(a[i], f);
  }
}

template <int R>
struct RegState {
  float2 a[1 << R];
  float2 u0[R], u1[R];  // pending separable diagonal factors per register bit
  float2 c;             // pending thread-constant factor
  unsigned dirty;       // bit j: u[j] pending; bit 31: c pending
};

#define TCB_SWITCH_J(R, j, ...)                                               \
  switch (j) {                                                                  \
    case 0: { constexpr int J = 0; __VA_ARGS__; } break;                        \
    case 1: if constexpr (R > 1) { constexpr int J = 1; __VA_ARGS__; } break;   \
    case 2: if constexpr (R > 2) { constexpr int J = 2; __VA_ARGS__; } break;   \
    case 3: if constexpr (R > 3) { constexpr int J = 3; __VA_ARGS__; } break;   \
    case 4: if constexpr (R > 4) { constexpr int J = 4; __VA_ARGS__; } break;   \
    default: break;                                                             \
  }

// take (and clear) the pending diagonal factor of register bit j, with c folded in
template <int R>
TCB_DEV void take_pending(RegState<R>& s, int j, float2& f0, float2& f1) {
  f0 = make_float2(1.f, 0.f);
  f1 = make_float2(1.f, 0.f);
  if ((s.dirty >> j) & 1u) {
    TCB_SWITCH_J(R, j, {
      f0 = s.u0[J];
      f1 = s.u1[J];
      s.u0[J] = make_float2(1.f, 0.f);
      s.u1[J] = make_float2(1.f, 0.f);
    })
    s.dirty &= ~(1u << j);
  }
  if (s.dirty >> 31) {
    f0 = cmul(f0, s.c);
    f1 = cmul(f1, s.c);
    s.c = make_float2(1.f, 0.f);
    s.dirty &= 0x7fffffffu;
  }
}

template <int R>
TCB_DEV void mul_u(RegState<R>& s, int j, float2 d0, float2 d1) {
  TCB_SWITCH_J(R, j, {
    s.u0[J] = cmul(s.u0[J], d0);
    s.u1[J] = cmul(s.u1[J], d1);
  })
  s.dirty |= (1u << j);
}

// XOR-fold of the bits above the low nibble (linear over GF(2)): swz(t) = (t & ~15) | ((t ^ fold_hi(t)) & 15)
TCB_DEV int fold_hi(int t) { return ((t >> 4) ^ (t >> 8) ^ (t >> 12)) & 15; }

// one thread's share of a register sub-pass.
//   tile : shared-memory tile (swizzled), sp : sub-pass header, gates : gate buffer of this
//   batch element, group : which 2^R-amplitude group this thread owns, cta_base : CTA-constant
//   flat-index bits (already OR-ed with index_base), hi_flat : optional table of tile_to_flat(h<<L)
template <int R>
TCB_DEV void run_reg_subpass(float2* tile, const int32_t* hdr, const int32_t* sp,
                             const float2* __restrict__ gates, int group, uint64_t cta_base,
                             const uint64_t* hi_flat = nullptr) {
  const int T = hdr[H_T];
  int tbase = 0;
  for (int b = 0; b < T - R; ++b) tbase |= ((group >> b) & 1) << sp[S_GRPBITS + b];
  uint64_t gidx;
  if (hi_flat != nullptr) {
    const int L = hdr[H_L];
    gidx = cta_base | hi_flat[tbase >> L] | (uint64_t)(tbase & ((1 << L) - 1));
  } else {
    gidx = cta_base | tile_to_flat(tbase, hdr);
  }

  // shared-memory address of amplitude i:  tbase and the register offsets have disjoint bits, and
  // the swizzle is linear, so   swz(tbase | off_i) = base_hi | off_i(hi part) | (low ^ fold) & 15
  int rb[R];   // register bit j -> tile bit mask
  int rsw[R];  // its contribution to the swizzled address: mask with the low nibble replaced by the fold
  TCB_UNROLL
  for (int j = 0; j < R; ++j) {
    const int m = 1 << sp[S_REGBITS + j];
    rb[j] = m;
    rsw[j] = (m & ~15) | ((m ^ fold_hi(m)) & 15);
  }
  const int sbase = (tbase & ~15) | ((tbase ^ fold_hi(tbase)) & 15);

  RegState<R> s;
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    s.a[i] = tile[ad];
  }
  TCB_UNROLL
  for (int j = 0; j < R; ++j) {
    s.u0[j] = make_float2(1.f, 0.f);
    s.u1[j] = make_float2(1.f, 0.f);
  }
  s.c = make_float2(1.f, 0.f);
  s.dirty = 0;

  const int nops = sp[S_NOPS];
  const int32_t* op = sp + SUB_HDR_WORDS;
  TCB_NOUNROLL
  for (int o = 0; o < nops; ++o, op += OP_WORDS) {
    const int code = op[O_CODE];
    const float2* m = gates + op[O_MAT];
    if (code == OP_1Q || code == OP_C1Q) {
      // dense (optionally controlled) 2x2 on register bit j.  Pending diagonal factors on that bit
      // (and the thread-constant factor) are folded into the matrix columns:  M' = M diag(f0, f1).
      const int rs = op[O_AUX1];
      int j, cmask = 0, cwant = 0;
      bool active = true;
      if (code == OP_1Q) {
        j = op[O_A];
      } else {
        j = op[O_B];
        const int pol = op[O_AUX0];
        const int qc[2] = {op[O_A], op[O_AUX2]};
        TCB_UNROLL
        for (int c = 0; c < 2; ++c) {
          const int q = qc[c];
          if (q < 0) continue;
          const int want = (pol >> c) & 1;
          if (q >= QREF_BIT) {
            active = active && ((int)((gidx >> (q - QREF_BIT)) & 1ull) == want);
          } else {
            cmask |= 1 << q;
            cwant |= want << q;
          }
        }
      }
      float2 m00 = make_float2(1.f, 0.f), m01 = make_float2(0.f, 0.f), m10 = m01, m11 = m00;
      if (active) {
        m00 = m[0];
        m01 = m[1];
        m10 = m[rs];
        m11 = m[rs + 1];
      }
      if (cmask == 0) {
        // (for an inactive thread-level control the matrix is the identity, so this still applies
        //  the pending diagonal — correct, and keeps one code path)
        const bool pend = ((s.dirty >> j) & 1u) || (s.dirty >> 31);
        if (pend || active) {
          float2 f0, f1;
          take_pending<R>(s, j, f0, f1);
          m00 = cmul(m00, f0);
          m10 = cmul(m10, f0);
          m01 = cmul(m01, f1);
          m11 = cmul(m11, f1);
          TCB_SWITCH_J(R, j, (apply_1q<R, J, false>(s.a, m00, m01, m10, m11, 0, 0)))
        }
      } else {
        // register-resident controls: the pending factor of bit j applies to every pair, the gate
        // only to the selected ones -> flush the factor first, then apply the masked gate
        if (((s.dirty >> j) & 1u) || (s.dirty >> 31)) {
          float2 f0, f1;
          take_pending<R>(s, j, f0, f1);
          TCB_SWITCH_J(R, j, (apply_bitdiag<R, J>(s.a, f0, f1)))
        }
        if (active) {
          TCB_SWITCH_J(R, j, (apply_1q<R, J, true>(s.a, m00, m01, m10, m11, cmask, cwant)))
        }
      }
    } else if (code == OP_DIAG1) {
      const int q = op[O_A];
      const int st = op[O_AUX1];
      const float2 d0 = m[0], d1 = m[st];
      if (q >= QREF_BIT) {
        const bool bit = (gidx >> (q - QREF_BIT)) & 1ull;
        s.c = cmul(s.c, bit ? d1 : d0);
        s.dirty |= 0x80000000u;
      } else {
        mul_u<R>(s, q, d0, d1);
      }
    } else if (code == OP_DIAG2) {
      const int qa = op[O_A], qb = op[O_B];
      const int st = op[O_AUX1];
      const bool a_bit = qa >= QREF_BIT, b_bit = qb >= QREF_BIT;
      if (a_bit || b_bit) {
        // resolve the thread-constant qubit(s): the gate collapses to a 1q diagonal or a scalar
        const bool xa = a_bit ? (bool)((gidx >> (qa - QREF_BIT)) & 1ull) : false;
        const bool xb = b_bit ? (bool)((gidx >> (qb - QREF_BIT)) & 1ull) : false;
        if (a_bit && b_bit) {
          s.c = cmul(s.c, m[(2 * (int)xa + (int)xb) * st]);
          s.dirty |= 0x80000000u;
        } else if (b_bit) {
          mul_u<R>(s, qa, m[(int)xb * st], m[(2 + (int)xb) * st]);
        } else {
          mul_u<R>(s, qb, m[(2 * (int)xa) * st], m[(2 * (int)xa + 1) * st]);
        }
      } else {
        // both register bits: apply directly (commutes with every pending diagonal factor)
        const float2 d00 = m[0], d01 = m[st], d10 = m[2 * st], d11 = m[3 * st];
        TCB_SWITCH_J(R, qa, (apply_pairdiag<R, J>(s.a, qb, d00, d01, d10, d11)))
      }
    }
  }
  // flush what is still pending
  TCB_NOUNROLL
  for (int j = 0; j < R; ++j) {
    if ((s.dirty >> j) & 1u) {
      float2 f0, f1;
      take_pending<R>(s, j, f0, f1);
      TCB_SWITCH_J(R, j, (apply_bitdiag<R, J>(s.a, f0, f1)))
    }
  }
  if (s.dirty >> 31) {
    float2 f0, f1;
    take_pending<R>(s, 0, f0, f1);
    apply_bitdiag<R, 0>(s.a, f0, f1);
  }

  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    tile[ad] = s.a[i];
  }
}

// ---- shared-memory dense sub-pass (k = 1..5 qubits, all inside the tile) ------------
template <int K>
TCB_DEV void run_smem_dense_group(float2* tile, const int32_t* op, const float2* __restrict__ m,
                                  int group) {
  int q[K];   // tile-bit index of gate qubit i (qubit 0 = matrix MSB)
  int qs[K];  // same, ascending
  const int src[5] = {O_A, O_B, O_AUX1, O_AUX2, O_AUX3};
  TCB_UNROLL
  for (int i = 0; i < K; ++i) q[i] = op[src[i]];
  TCB_UNROLL
  for (int i = 0; i < K; ++i) qs[i] = q[i];
  // insertion sort (K <= 5)
  TCB_UNROLL
  for (int i = 1; i < K; ++i) {
    TCB_UNROLL
    for (int j = i; j > 0; --j) {
      if (qs[j - 1] > qs[j]) {
        const int tmp = qs[j];
        qs[j] = qs[j - 1];
        qs[j - 1] = tmp;
      }
    }
  }
  int base = group;
  TCB_UNROLL
  for (int i = 0; i < K; ++i) base = ((base >> qs[i]) << (qs[i] + 1)) | (base & ((1 << qs[i]) - 1));
  int off[1 << K];
  float2 v[1 << K];
  TCB_UNROLL
  for (int c = 0; c < (1 << K); ++c) {
    int o = 0;
    TCB_UNROLL
    for (int i = 0; i < K; ++i)
      if ((c >> (K - 1 - i)) & 1) o |= 1 << q[i];
    off[c] = o;
    v[c] = tile[swz(base | o)];
  }
  TCB_UNROLL
  for (int r = 0; r < (1 << K); ++r) {
    float2 acc = make_float2(0.f, 0.f);
    TCB_UNROLL
    for (int c = 0; c < (1 << K); ++c) acc = cfma(m[r * (1 << K) + c], v[c], acc);
    tile[swz(base | off[r])] = acc;
  }
}

TCB_DEV void run_smem_dense(float2* tile, const int32_t* hdr, const int32_t* sp,
                            const float2* __restrict__ gates, int tid, int nthreads) {
  const int T = hdr[H_T];
  const int32_t* op = sp + SUB_HDR_WORDS;
  const int k = op[O_AUX0];
  const float2* m = gates + op[O_MAT];
  const int ngroups = 1 << (T - k);
  for (int g = tid; g < ngroups; g += nthreads) {
    switch (k) {
      case 1: run_smem_dense_group<1>(tile, op, m, g); break;
      case 2: run_smem_dense_group<2>(tile, op, m, g); break;
      case 3: run_smem_dense_group<3>(tile, op, m, g); break;
      case 4: run_smem_dense_group<4>(tile, op, m, g); break;
      default: break;
    }
  }
}

}  // namespace tcb

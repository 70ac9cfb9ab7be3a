// pass_core.cuh — the fused statevector "tile pass": one HBM pass applies a whole
// program of gates.  Replaces the per-gate tensordot + transpose loop of the
// reference (tensorcircuit/cons.py:937-953) for circuit-shaped networks.
//
// Execution model (B200):
//   * the state (2^n complex64) is cut into tiles of 2^T amplitudes (T <= 13, 64 KB):
//     the T "tile bits" are the L lowest index bits (coalesced runs) plus T-L arbitrary
//     higher bits chosen by the host planner; all other index bits are constant per CTA.
//   * one CTA = one tile at a time: load -> shared memory -> a sequence of sub-passes -> store.
//   * a *register sub-pass* pulls 2^R amplitudes per thread (R = 5) into registers.  Register
//     slot 0 is ALWAYS tile bit 0, so every shared-memory access is a 128-bit LDS/STS of an
//     adjacent amplitude pair; slots 1..4 are tile bits chosen by the planner.  The sub-pass
//     executes a list of ROUNDS.  A round is, per register slot J (compile-time J):
//         1. an optional diagonal factor (f0, f1) on that bit.  Diagonal 1q/2q gates are
//            scheduled LAZILY by the planner: each one is attached to a round in which one of its
//            qubits is a register bit, so the factor is per-THREAD data (the partner bit is a
//            CTA constant, a thread constant or another register bit) — never per amplitude;
//         2. an optional fused 2x2 (optionally controlled) gate; the factor is folded into the
//            matrix columns (M' = M diag(f0, f1)), so a whole QAOA cost layer costs a handful of
//            multiplies per thread.
//   * k-qubit dense gates (k = 2..4) run as shared-memory sub-passes inside the same pass.
//
// The code is host/device neutral (TCB_DEV): the CUDA kernel in pass_kernel.cu and the CPU
// logic emulator in tests/emu/ (test infrastructure, never loaded by the package) both include
// it, so index arithmetic is validated without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TCB_DEV __device__ __forceinline__
#define TCB_UNROLL _Pragma("unroll")
#define TCB_NOUNROLL _Pragma("unroll 1")
#define TCB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#include <cmath>
#define TCB_DEV inline
#define TCB_UNROLL
#define TCB_NOUNROLL
#define TCB_FMA(a, b, c) std::fmaf((a), (b), (c))
struct float2 {
  float x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float2 make_float2(float a, float b) {
  float2 r;
  r.x = a;
  r.y = b;
  return r;
}
static inline float4 make_float4(float a, float b, float c, float d) {
  float4 r;
  r.x = a;
  r.y = b;
  r.z = c;
  r.w = d;
  return r;
}
#endif

#ifndef TCB_SUBPROF
#define TCB_SUBPROF(slot)  // (PASS_PROFILE build of pass_kernel.cu: cycle stamps inside a register sub-pass)
#define TCB_SUBPROF_DECL
#endif

namespace tcb {

// ---- program layout (int32 words) -------------------------------------------
constexpr int PASS_MAGIC = 0x7CB20004;
constexpr int PASS_MAX_T = 13;
constexpr int PASS_R = 5;  // register bits per sub-pass (32 amplitudes / thread); slot 0 = tile bit 0
constexpr int PASS_MAX_WORDS = 6144;
constexpr int PASS_MAX_POOL = 1536;  // complex elements (12 KB)
constexpr int PASS_MAX_SUB = 16;     // sub-passes per pass
// header
constexpr int H_MAGIC = 0, H_T = 1, H_L = 2, H_NSUB = 3, H_WORDS = 4, H_NNONTILE = 5, H_R = 6;
constexpr int H_NFILL = 7;        // number of fill records; their word offsets are the last H_NFILL words
// gate pool: the gate tensors a pass needs are copied once per CTA from the gate buffer into shared
// memory; fill sources address the pool.  Pool table = H_NPOOL triples (gatebuf offset, #elements,
// pool offset) stored right before the fill-offset table.
constexpr int H_NPOOL = 64, H_POOLSIZE = 65;
// fill records are ordered [static | dynamic short | dynamic long]: static ones do not depend on
// CTA bits (once per CTA); short ones (<= FILL_SHORT sources) are resolved by one thread each,
// long ones by a warp each (shuffle-tree product)
constexpr int H_NFILL_STATIC = 66, H_NFILL_SHORT_END = 67;
constexpr int FILL_SHORT = 4;
constexpr int H_TILEPOS = 8;      // [16] flat-index bit position of tile bit t
constexpr int H_NONTILEPOS = 24;  // [40] flat-index bit positions of the non-tile bits, ascending
constexpr int HDR_WORDS = 80;
// sub-pass header
constexpr int S_NROUNDS = 0, S_KIND = 1, S_REGBITS = 2 /*[8]*/, S_GRPBITS = 10 /*[12]*/, S_WORDS = 22;
constexpr int SUB_HDR_WORDS = 24;
constexpr int SUB_REG = 0, SUB_SMEM_DENSE = 1;
// round record (register sub-pass); every float block is 16-byte aligned
constexpr int RD_FLAGS = 0;   // bit J: gate on register slot J; bit 8+J: that gate has register-resident
                              // controls; bit 16+J: diagonal factor on slot J; bit 30: the round needs
                              // the general path (a controlled gate, a factor without gate, an RR table)
constexpr int RD_GENERAL = 1 << 30;
constexpr int RD_NRR = 1;     // # two-bit diagonal tables with both bits in registers
constexpr int RD_WORDS = 2;   // total words of this round (fixed part + tables)
constexpr int RD_NRJ = 3;     // [5] # tables with register slot J and a thread-constant tile bit
constexpr int RD_CTRL = 8;    // [5][2] control words of the gate on slot J
constexpr int RD_M = 20;      // [5][4] float2 fused matrices m00 m01 m10 m11      (device filled)
constexpr int RD_F = 60;      // [5][2] float2 diagonal factor f0 f1 of slot J     (device filled)
constexpr int RD_FIXED = 80;
// two-bit table entry.  RJ tables: A = register slot, B = tile-bit index of the partner,
//   W = [x_B = 0: (d[x_A=0], d[x_A=1])], [x_B = 1: (d[x_A=0], d[x_A=1])]   (one LDS.128 per lookup)
// RR tables: A, B = register slots, W[2 * x_A + x_B]
constexpr int TT_A = 0, TT_B = 1, TT_W = 4, TT_WORDS = 12;
// control words:  w0 = cmask | cwant << 8 | n_thread_ctrl << 16
//                 w1 = c0 | c1 << 16,  c = p | in_tile << 7 | pol << 8   (p: tile-bit index if in_tile,
//                                                                         else flat bit position)
// dense shared-memory op (SUB_SMEM_DENSE): one 16-word record after the sub-pass header
constexpr int OP_WORDS = 16;
constexpr int O_CODE = 0, O_A = 1, O_B = 2, O_MAT = 3, O_AUX0 = 4, O_AUX1 = 5, O_AUX2 = 6, O_AUX3 = 7;
constexpr int OP_DENSE = 5;  // AUX0 = k, tile-bit indices of the gate qubits in A,B,AUX1,AUX2; MAT -> 2^k x 2^k
// fill records (device prologue):  [dst word offset, kind, count, 0] + count x [mat_off, form | stride << 8, p, q]
constexpr int FK_PAIR = 1, FK_TABLE = 2, FK_MATRIX = 3;
enum FillForm : int {
  FF_D1 = 0,         // pair  (m[0], m[st])                          1q diagonal on the slot's bit
  FF_D2_FIRST = 1,   // pair  (m[x*st], m[(2+x)*st]),   x = bit p    slot bit is the gate's first qubit
  FF_D2_SECOND = 2,  // pair  (m[2x*st], m[(2x+1)*st]), x = bit p    slot bit is the gate's second qubit
  FF_T = 5,          // table (m[0], m[st], m[2st], m[3st])
  FF_T_SWAP = 6,     // table transposed (m[0], m[2st], m[st], m[3st])
  FF_M = 7,          // matrix (m[0], m[1], m[st], m[st+1])          dense 2x2, row stride st
  FF_MD = 8,         // matrix diag(m[0], m[st])
};

// ---- complex helpers ------------------------------------------------------------
TCB_DEV float2 cmul(float2 a, float2 b) {
  return make_float2(TCB_FMA(a.x, b.x, -(a.y * b.y)), TCB_FMA(a.x, b.y, a.y * b.x));
}
TCB_DEV float2 cfma(float2 a, float2 b, float2 c) {  // a*b + c, as two dependent FMAs per component
  return make_float2(TCB_FMA(-a.y, b.y, TCB_FMA(a.x, b.x, c.x)), TCB_FMA(a.y, b.x, TCB_FMA(a.x, b.y, c.y)));
}
TCB_DEV float2 csel(bool p, float2 a, float2 b) { return p ? a : b; }
TCB_DEV float2 c_one() { return make_float2(1.f, 0.f); }

// ---- packed FP32x2 (Blackwell FFMA2 / FMUL2) -----------------------------------------------------
// A complex number is one aligned 64-bit register pair.  The product m * x is formed by broadcasting
// the two COMPONENTS of the amplitude against the matrix entry kept as two pairs,
//     m * x = x.re * (m.re, m.im) + x.im * (-m.im, m.re),
// so that every FFMA2 takes its scalar through the .F32 broadcast modifier and its pair operand as a
// plain (or LO_HI-swapped) register pair: no per-amplitude MOVs (the mirror formulation, broadcasting
// the matrix entry against (x, swap(x)), makes ptxas rematerialise an (im, im) pair before each use).
#if defined(__CUDACC__)
struct PackedM {
  float2 m00, n00, m01, n01, m10, n10, m11, n11;  // m = (re, im), n = (-im, re)
};
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }
// n = (-im, re) is produced by ONE packed multiply of the swapped entry with (-1, 1): a genuine
// register pair.  (Built from two scalars, ptxas re-creates the pair with 2 MOVs in front of almost
// every use instead of keeping it live: +50% instructions in the gate loop.)
__device__ __forceinline__ float2 rot90(float2 m) { return __fmul2_rn(make_float2(m.y, m.x), make_float2(-1.f, 1.f)); }
__device__ __forceinline__ PackedM pack_matrix(float2 m00, float2 m01, float2 m10, float2 m11) {
  PackedM p;
  p.m00 = m00; p.n00 = rot90(m00);
  p.m01 = m01; p.n01 = rot90(m01);
  p.m10 = m10; p.n10 = rot90(m10);
  p.m11 = m11; p.n11 = rot90(m11);
  return p;
}
// (b0, b1) = M (x, y)
__device__ __forceinline__ void packed_2x2(const PackedM& p, float2 x, float2 y, float2& b0, float2& b1) {
  float2 r0 = __fmul2_rn(bcast2(x.x), p.m00);
  float2 r1 = __fmul2_rn(bcast2(x.x), p.m10);
  r0 = __ffma2_rn(bcast2(x.y), p.n00, r0);
  r1 = __ffma2_rn(bcast2(x.y), p.n10, r1);
  r0 = __ffma2_rn(bcast2(y.x), p.m01, r0);
  r1 = __ffma2_rn(bcast2(y.x), p.m11, r1);
  r0 = __ffma2_rn(bcast2(y.y), p.n01, r0);
  r1 = __ffma2_rn(bcast2(y.y), p.n11, r1);
  b0 = r0;
  b1 = r1;
}
// x * f (complex) as two packed instructions; fn = (-f.im, f.re)
__device__ __forceinline__ float2 packed_cmul(float2 x, float2 f, float2 fn) {
  return __ffma2_rn(bcast2(x.y), fn, __fmul2_rn(bcast2(x.x), f));
}
#endif

// ---- shared-memory tile layout -------------------------------------------------------------------
// The tile is addressed in 8-byte amplitudes; bit 0 (the pair bit) is never swizzled, so a pair is
// one aligned 16-byte chunk.  The 3 bits that pick the 16-byte chunk inside a 128-byte row are
// XOR-folded with the higher index bits (GF(2)-linear): tile bits 1,4,7,10 land on chunk bit 0,
// bits 2,5,8,11 on chunk bit 1, bits 3,6,9,12 on chunk bit 2.  A quarter-warp LDS.128 is
// conflict-free when its 3 lane bits fall into 3 distinct classes (planner: _order_group_bits).
TCB_DEV int fold_hi(int t) { return ((t >> 4) ^ (t >> 7) ^ (t >> 10)) & 7; }
TCB_DEV int swz(int t) { return t ^ (fold_hi(t) << 1); }

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

// tile index of the thread `group` of a register sub-pass (register bits zero)
TCB_DEV int group_to_tile(int group, int T, int R, const int32_t* sp) {
  int tbase = 0;
  for (int b = 0; b < T - R; ++b) tbase |= ((group >> b) & 1) << sp[S_GRPBITS + b];
  return tbase;
}

// ---- CTA prologue: fill records -----------------------------------------------------
// Every source of a fill record is turned into a 2x2 complex matrix E (diagonal pair ->
// diag(d0, d1), dense 1q gate -> M, two-bit table -> its 4 entries as is), and a record's value is
// the ordered product  E_{count-1} ... E_1 E_0  (later gates multiply from the left; diagonal kinds
// commute elementwise).  Short records are resolved by one thread, long ones by a warp (lane i
// loads source i, the product is a shuffle tree, pass_kernel.cu); the emulator runs the serial form.
struct Mat2 {
  float2 a, b, c, d;  // [[a, b], [c, d]]
};
TCB_DEV Mat2 mat2_identity() {
  Mat2 m;
  m.a = c_one();
  m.b = make_float2(0.f, 0.f);
  m.c = m.b;
  m.d = c_one();
  return m;
}
// later * earlier
TCB_DEV Mat2 mat2_mul(const Mat2& l, const Mat2& e) {
  Mat2 r;
  r.a = cfma(l.b, e.c, cmul(l.a, e.a));
  r.b = cfma(l.b, e.d, cmul(l.a, e.b));
  r.c = cfma(l.d, e.c, cmul(l.c, e.a));
  r.d = cfma(l.d, e.d, cmul(l.c, e.b));
  return r;
}
// combine for a record of kind `kind`: tables and pairs multiply entrywise, matrices as matrices
TCB_DEV Mat2 fill_combine(int kind, const Mat2& later, const Mat2& earlier) {
  if (kind == FK_MATRIX) return mat2_mul(later, earlier);
  Mat2 r;
  r.a = cmul(later.a, earlier.a);
  r.d = cmul(later.d, earlier.d);
  if (kind == FK_TABLE) {
    r.b = cmul(later.b, earlier.b);
    r.c = cmul(later.c, earlier.c);
  } else {
    r.b = make_float2(0.f, 0.f);
    r.c = r.b;
  }
  return r;
}
TCB_DEV Mat2 fill_identity(int kind) {
  Mat2 m = mat2_identity();
  if (kind == FK_TABLE) {
    m.b = c_one();
    m.c = c_one();
  }
  return m;
}
// load source `src` (4 words) of a record of kind `kind`
TCB_DEV Mat2 load_fill_source(const int32_t* src, int kind, const float2* __restrict__ gates,
                              uint64_t cta_bits) {
  const float2* m = gates + src[0];
  const int form = src[1] & 0xff, st = src[1] >> 8;
  const int x = (int)((cta_bits >> src[2]) & 1ull);
  Mat2 e = mat2_identity();
  if (kind == FK_PAIR) {
    if (form == FF_D1) {
      e.a = m[0];
      e.d = m[st];
    } else if (form == FF_D2_FIRST) {
      e.a = m[x * st];
      e.d = m[(2 + x) * st];
    } else {
      e.a = m[(2 * x) * st];
      e.d = m[(2 * x + 1) * st];
    }
  } else if (kind == FK_TABLE) {
    e.a = m[0];
    e.d = m[3 * st];
    e.b = (form == FF_T) ? m[st] : m[2 * st];
    e.c = (form == FF_T) ? m[2 * st] : m[st];
  } else {
    if (form == FF_M) {
      e.a = m[0];
      e.b = m[1];
      e.c = m[st];
      e.d = m[st + 1];
    } else {
      e.a = m[0];
      e.d = m[st];
    }
  }
  return e;
}
TCB_DEV void store_fill_result(int32_t* prog, const int32_t* rec, const Mat2& r) {
  float2* dst = reinterpret_cast<float2*>(prog + rec[0]);
  const int kind = rec[1];
  if (kind == FK_PAIR) {
    dst[0] = r.a;
    dst[1] = r.d;
  } else {
    dst[0] = r.a;
    dst[1] = r.b;
    dst[2] = r.c;
    dst[3] = r.d;
  }
}
// serial form (one thread per record; also the emulator)
TCB_DEV void run_fill_record(int32_t* prog, int rec_off, const float2* __restrict__ gates,
                             uint64_t cta_bits) {
  const int32_t* rec = prog + rec_off;
  const int kind = rec[1], count = rec[2];
  Mat2 acc = load_fill_source(rec + 4, kind, gates, cta_bits);
  for (int i = 1; i < count; ++i)
    acc = fill_combine(kind, load_fill_source(rec + 4 + 4 * i, kind, gates, cta_bits), acc);
  store_fill_result(prog, rec, acc);
}

// ---- register sub-pass -------------------------------------------------------------
template <int R, int J, bool CTRL>
TCB_DEV void apply_1q(float2 (&a)[1 << R], float2 m00, float2 m01, float2 m10, float2 m11, int cmask,
                      int cwant) {
#if defined(__CUDACC__)
  const PackedM pm = pack_matrix(m00, m01, m10, m11);
#endif
  TCB_UNROLL
  for (int p = 0; p < (1 << (R - 1)); ++p) {
    const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
    const int i1 = i0 | (1 << J);
    const float2 x = a[i0], y = a[i1];
    float2 b0, b1;
#if defined(__CUDACC__)
    packed_2x2(pm, x, y, b0, b1);
#else
    b0 = cfma(m01, y, cmul(m00, x));
    b1 = cfma(m11, y, cmul(m10, x));
#endif
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

template <int R, int J>
TCB_DEV void apply_bitdiag(float2 (&a)[1 << R], float2 u0, float2 u1) {
#if defined(__CUDACC__)
  const float2 n0 = rot90(u0), n1 = rot90(u1);
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) a[i] = ((i >> J) & 1) ? packed_cmul(a[i], u1, n1) : packed_cmul(a[i], u0, n0);
#else
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) a[i] = cmul(a[i], ((i >> J) & 1) ? u1 : u0);
#endif
}

// diagonal table on two register slots JA < JB (compile-time): a[i] *= d[2 * i_JA + i_JB]
template <int R, int JA, int JB>
TCB_DEV void apply_pairdiag_ct(float2 (&a)[1 << R], float2 d00, float2 d01, float2 d10, float2 d11) {
  if constexpr (JB < R) {
#if defined(__CUDACC__)
    const float2 n00 = rot90(d00), n01 = rot90(d01), n10 = rot90(d10), n11 = rot90(d11);
    TCB_UNROLL
    for (int i = 0; i < (1 << R); ++i) {
      const int c = 2 * ((i >> JA) & 1) + ((i >> JB) & 1);
      a[i] = c == 0 ? packed_cmul(a[i], d00, n00)
                    : (c == 1 ? packed_cmul(a[i], d01, n01) : (c == 2 ? packed_cmul(a[i], d10, n10) : packed_cmul(a[i], d11, n11)));
    }
#else
    TCB_UNROLL
    for (int i = 0; i < (1 << R); ++i) {
      const int c = 2 * ((i >> JA) & 1) + ((i >> JB) & 1);
      a[i] = cmul(a[i], c == 0 ? d00 : (c == 1 ? d01 : (c == 2 ? d10 : d11)));
    }
#endif
  }
}
template <int R>
TCB_DEV void apply_pairdiag(float2 (&a)[1 << R], int j, int k, float2 d00, float2 d01, float2 d10,
                            float2 d11) {
  switch (j * 8 + k) {  // j < k (planner)
    case 0 * 8 + 1: apply_pairdiag_ct<R, 0, 1>(a, d00, d01, d10, d11); break;
    case 0 * 8 + 2: apply_pairdiag_ct<R, 0, 2>(a, d00, d01, d10, d11); break;
    case 0 * 8 + 3: apply_pairdiag_ct<R, 0, 3>(a, d00, d01, d10, d11); break;
    case 0 * 8 + 4: apply_pairdiag_ct<R, 0, 4>(a, d00, d01, d10, d11); break;
    case 1 * 8 + 2: apply_pairdiag_ct<R, 1, 2>(a, d00, d01, d10, d11); break;
    case 1 * 8 + 3: apply_pairdiag_ct<R, 1, 3>(a, d00, d01, d10, d11); break;
    case 1 * 8 + 4: apply_pairdiag_ct<R, 1, 4>(a, d00, d01, d10, d11); break;
    case 2 * 8 + 3: apply_pairdiag_ct<R, 2, 3>(a, d00, d01, d10, d11); break;
    case 2 * 8 + 4: apply_pairdiag_ct<R, 2, 4>(a, d00, d01, d10, d11); break;
    case 3 * 8 + 4: apply_pairdiag_ct<R, 3, 4>(a, d00, d01, d10, d11); break;
    default: break;
  }
}

TCB_DEV float2 lds2(const int32_t* p) { return *reinterpret_cast<const float2*>(p); }
TCB_DEV float4 lds4(const int32_t* p) { return *reinterpret_cast<const float4*>(p); }

// compile-time dispatch helpers (J is a literal after unrolling; out-of-range J never runs)
template <int R, int J, bool CTRL>
TCB_DEV void gate_on(float2 (&a)[1 << R], float2 m00, float2 m01, float2 m10, float2 m11, int cmask,
                     int cwant) {
  if constexpr (J < R) apply_1q<R, J, CTRL>(a, m00, m01, m10, m11, cmask, cwant);
}
template <int R, int J>
TCB_DEV void diag_on(float2 (&a)[1 << R], float2 f0, float2 f1) {
  if constexpr (J < R) apply_bitdiag<R, J>(a, f0, f1);
}

// one thread-constant control: true when the control bit has the wanted polarity
TCB_DEV bool ctrl_ok(int c, int tbase, uint64_t cta_bits) {
  const int p = c & 0x7f;
  const int bit = (c & 0x80) ? ((tbase >> p) & 1) : (int)((cta_bits >> p) & 1ull);
  return bit == ((c >> 8) & 1);
}

// register slot J (compile-time) of one round: its diagonal factor, then its gate.
//   tt : cursor into the round's RJ tables (section J follows section J-1)
template <int R, int J>
TCB_DEV void round_bit(float2 (&a)[1 << R], const int32_t* rd, int flags, int tbase, uint64_t cta_bits,
                       const int32_t*& tt) {
  if constexpr (J < R) {
    const bool gate = (flags >> J) & 1, fac = (flags >> (16 + J)) & 1;
    if (!gate && !fac) return;
    float2 f0 = c_one(), f1 = c_one();
    if (fac) {
      const float4 F = lds4(rd + RD_F + 4 * J);
      f0 = make_float2(F.x, F.y);
      f1 = make_float2(F.z, F.w);
      const int nj = rd[RD_NRJ + J];
      TCB_NOUNROLL
      for (int e = 0; e < nj; ++e, tt += TT_WORDS) {
        const int xb = (tbase >> tt[TT_B]) & 1;
        const float4 W = lds4(tt + TT_W + 4 * xb);
        f0 = cmul(f0, make_float2(W.x, W.y));
        f1 = cmul(f1, make_float2(W.z, W.w));
      }
    }
    if (!gate) {
      diag_on<R, J>(a, f0, f1);
      return;
    }
    const int cw0 = rd[RD_CTRL + 2 * J], cw1 = rd[RD_CTRL + 2 * J + 1];
    const int ntc = (cw0 >> 16) & 3;
    bool active = true;
    if (ntc > 0) active = ctrl_ok(cw1 & 0xffff, tbase, cta_bits);
    if (ntc > 1) active = active && ctrl_ok((cw1 >> 16) & 0xffff, tbase, cta_bits);
    float2 m00 = c_one(), m01 = make_float2(0.f, 0.f), m10 = m01, m11 = c_one();
    if (active) {
      const float4 A = lds4(rd + RD_M + 8 * J), B = lds4(rd + RD_M + 8 * J + 4);
      m00 = make_float2(A.x, A.y);
      m01 = make_float2(A.z, A.w);
      m10 = make_float2(B.x, B.y);
      m11 = make_float2(B.z, B.w);
    }
    if (!((flags >> (8 + J)) & 1)) {
      if (fac) {  // fold the diagonal factor into the matrix columns:  M' = M diag(f0, f1)
        m00 = cmul(m00, f0);
        m10 = cmul(m10, f0);
        m01 = cmul(m01, f1);
        m11 = cmul(m11, f1);
      } else if (!active) {
        return;
      }
      gate_on<R, J, false>(a, m00, m01, m10, m11, 0, 0);
    } else {
      // register-resident controls: the factor acts on every pair, the gate only on selected ones
      if (fac) diag_on<R, J>(a, f0, f1);
      if (active) gate_on<R, J, true>(a, m00, m01, m10, m11, cw0 & 0xff, (cw0 >> 8) & 0xff);
    }
  }
}

// fast path of round_bit: an uncontrolled gate with an optional factor (the whole of a QAOA / VQE
// layer).  Kept apart from the general path so that the hot loop is one compact stretch of code
// (the general path's controlled / diagonal-only variants are 2/3 of the instructions and would
// otherwise sit between the hot blocks: instruction-cache misses were 17% of all stall samples).
template <int R, int J>
TCB_DEV void round_bit_simple(float2 (&a)[1 << R], const int32_t* rd, int flags, int tbase, const int32_t*& tt) {
  if constexpr (J < R) {
    if (!((flags >> J) & 1)) return;
    const float4 A = lds4(rd + RD_M + 8 * J), B = lds4(rd + RD_M + 8 * J + 4);
    float2 m00 = make_float2(A.x, A.y), m01 = make_float2(A.z, A.w);
    float2 m10 = make_float2(B.x, B.y), m11 = make_float2(B.z, B.w);
    if ((flags >> (16 + J)) & 1) {
      const float4 F = lds4(rd + RD_F + 4 * J);
      float2 f0 = make_float2(F.x, F.y), f1 = make_float2(F.z, F.w);
      const int nj = rd[RD_NRJ + J];
      TCB_NOUNROLL
      for (int e = 0; e < nj; ++e, tt += TT_WORDS) {
        const int xb = (tbase >> tt[TT_B]) & 1;
        const float4 W = lds4(tt + TT_W + 4 * xb);
        f0 = cmul(f0, make_float2(W.x, W.y));
        f1 = cmul(f1, make_float2(W.z, W.w));
      }
      m00 = cmul(m00, f0);
      m10 = cmul(m10, f0);
      m01 = cmul(m01, f1);
      m11 = cmul(m11, f1);
    }
    gate_on<R, J, false>(a, m00, m01, m10, m11, 0, 0);
  }
}

// one thread's share of a register sub-pass.
//   tile : shared-memory tile (swizzled), sp : sub-pass header, tbase : the thread's tile index with
//   the register bits zero, cta_bits : CTA-constant flat-index bits (tile base | index_base)
template <int R>
TCB_DEV void run_reg_subpass(float2* tile, const int32_t* sp, int tbase, uint64_t cta_bits) {
  static_assert(R >= 2 && R <= 5, "register bits");
  TCB_SUBPROF_DECL
  // shared-memory address of amplitude i:  tbase and the register offsets have disjoint bits and
  // the swizzle is GF(2)-linear, so  swz(tbase | off_i) = swz(tbase) ^ XOR_j swz(1 << r_j).
  // Slot 0 is tile bit 0 (never swizzled): amplitudes 2p and 2p+1 are one 16-byte chunk.
  int rsw[R];
  TCB_UNROLL
  for (int j = 1; j < R; ++j) rsw[j] = swz(1 << sp[S_REGBITS + j]);
  const int sbase = swz(tbase);

  float2 a[1 << R];
  TCB_UNROLL
  for (int i = 0; i < (1 << R); i += 2) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 1; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    const float4 v = *reinterpret_cast<const float4*>(tile + ad);
    a[i] = make_float2(v.x, v.y);
    a[i + 1] = make_float2(v.z, v.w);
  }

  TCB_SUBPROF(8);
  const int nrounds = sp[S_NROUNDS];
  const int32_t* rd = sp + SUB_HDR_WORDS;
  TCB_NOUNROLL
  for (int r = 0; r < nrounds; ++r) {
    const int flags = rd[RD_FLAGS];
    const int32_t* tt = rd + RD_FIXED;
    // tables on two register slots act on the amplitudes directly; they belong to the diagonal
    // part, i.e. BEFORE this round's gates (their records sit after the per-slot sections)
    const int nrr = rd[RD_NRR];
    if (nrr > 0) {
      int nrj = 0;
      TCB_UNROLL
      for (int j = 0; j < R; ++j) nrj += rd[RD_NRJ + j];
      const int32_t* trr = tt + TT_WORDS * nrj;
      TCB_NOUNROLL
      for (int e = 0; e < nrr; ++e, trr += TT_WORDS)
        apply_pairdiag<R>(a, trr[TT_A], trr[TT_B], lds2(trr + TT_W), lds2(trr + TT_W + 2), lds2(trr + TT_W + 4),
                          lds2(trr + TT_W + 6));
    }
    if (!(flags & RD_GENERAL)) {
      round_bit_simple<R, 0>(a, rd, flags, tbase, tt);
      round_bit_simple<R, 1>(a, rd, flags, tbase, tt);
      round_bit_simple<R, 2>(a, rd, flags, tbase, tt);
      round_bit_simple<R, 3>(a, rd, flags, tbase, tt);
      round_bit_simple<R, 4>(a, rd, flags, tbase, tt);
      rd += rd[RD_WORDS];
      continue;
    }
    round_bit<R, 0>(a, rd, flags, tbase, cta_bits, tt);
    round_bit<R, 1>(a, rd, flags, tbase, cta_bits, tt);
    round_bit<R, 2>(a, rd, flags, tbase, cta_bits, tt);
    round_bit<R, 3>(a, rd, flags, tbase, cta_bits, tt);
    round_bit<R, 4>(a, rd, flags, tbase, cta_bits, tt);
    rd += rd[RD_WORDS];
  }
  TCB_SUBPROF(9);

  TCB_UNROLL
  for (int i = 0; i < (1 << R); i += 2) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 1; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    *reinterpret_cast<float4*>(tile + ad) = make_float4(a[i].x, a[i].y, a[i + 1].x, a[i + 1].y);
  }
  TCB_SUBPROF(10);
}

// ---- shared-memory dense sub-pass (k = 1..4 qubits, all inside the tile) ------------
template <int K>
TCB_DEV void run_smem_dense_group(float2* tile, const int32_t* op, const float2* __restrict__ m,
                                  int group) {
  int q[K];   // tile-bit index of gate qubit i (qubit 0 = matrix MSB)
  int qs[K];  // same, ascending
  const int src[4] = {O_A, O_B, O_AUX1, O_AUX2};
  TCB_UNROLL
  for (int i = 0; i < K; ++i) q[i] = op[src[i]];
  TCB_UNROLL
  for (int i = 0; i < K; ++i) qs[i] = q[i];
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

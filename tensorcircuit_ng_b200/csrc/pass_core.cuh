// pass_core.cuh — the fused statevector "tile pass": one HBM pass applies a whole
// program of gates.  Replaces the per-gate tensordot + transpose loop of the
// reference (tensorcircuit/cons.py:937-953) for circuit-shaped networks.
//
// Execution model (B200):
//   * the state (2^n complex64) is cut into tiles of 2^T amplitudes (T <= 13, 64 KB):
//     the T "tile bits" are the L lowest index bits (coalesced 128 B+ runs) plus T-L
//     arbitrary higher bits chosen by the host planner; all other index bits are
//     constant per CTA.
//   * one CTA = one tile: load -> shared memory -> a sequence of sub-passes -> store.
//   * a *register sub-pass* pulls 2^R amplitudes per thread (R = 5 "register bits" of the
//     tile) into registers and executes a list of ROUNDS.  A round is fully static code:
//         1. a diagonal part: every diagonal 1q/2q gate scheduled here, on ANY qubits of the
//            register.  The CTA prologue has already resolved each gate against the
//            CTA-constant (non-tile) bits and multiplied the results into one scalar C and one
//            2-entry factor F_b per tile bit b; gates on two tile bits stay as 4-entry tables.
//            A thread folds C and the F_b of its 8 thread-constant tile bits into one pending
//            scalar, and the F_b of its register bits into pending per-bit factors.
//         2. at most one fused 2x2 (optionally controlled) gate per register bit J = 0..4,
//            compile-time J.  Consecutive 1q gates on one qubit were multiplied together by the
//            prologue; pending diagonal factors are folded into the matrix columns
//            (M' = M diag(f0, f1)), so a whole QAOA cost layer costs a handful of multiplies
//            per THREAD, not per amplitude.
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
constexpr int PASS_MAGIC = 0x7CB20003;
constexpr int PASS_MAX_T = 13;
constexpr int PASS_R = 4;  // register bits per sub-pass of the compiled kernel (16 amplitudes / thread)
constexpr int PASS_MAX_WORDS = 6144;
constexpr int PASS_MAX_POOL = 1536;  // complex elements (12 KB)
// header
constexpr int H_MAGIC = 0, H_T = 1, H_L = 2, H_NSUB = 3, H_WORDS = 4, H_NNONTILE = 5, H_R = 6;
constexpr int H_NFILL = 7;        // number of fill records; their word offsets are the last H_NFILL words
// gate pool: the gate tensors a pass needs are copied once per CTA from the gate buffer into shared
// memory; fill sources address the pool.  Pool table = H_NPOOL triples (gatebuf offset, #elements,
// pool offset) stored right before the fill-offset table.
constexpr int H_NPOOL = 64, H_POOLSIZE = 65;
constexpr int H_NFILL_STATIC = 66;  // the first H_NFILL_STATIC fill records do not depend on CTA bits
constexpr int H_TILEPOS = 8;      // [16] flat-index bit position of tile bit t
constexpr int H_NONTILEPOS = 24;  // [40] flat-index bit positions of the non-tile bits, ascending
constexpr int HDR_WORDS = 80;
// sub-pass header
constexpr int S_NROUNDS = 0, S_KIND = 1, S_REGBITS = 2 /*[8]*/, S_GRPBITS = 10 /*[12]*/, S_WORDS = 22;
constexpr int SUB_HDR_WORDS = 24;
constexpr int SUB_REG = 0, SUB_SMEM_DENSE = 1;
// round record (register sub-pass)
constexpr int RD_FLAGS = 0;   // bit J: gate on register bit J; bit 8+J: that gate has register-resident
                              // controls; bit 16: diagonal part present
constexpr int RD_FMASK = 1;   // bit b: F_b is not the identity
constexpr int RD_NNN = 2;     // # two-tile-bit diagonal tables with both bits thread-constant
constexpr int RD_NRR = 3;     // # with both bits in registers
constexpr int RD_NRJ = 4;     // [5] # with register bit J first and a thread-constant second bit
constexpr int RD_WORDS = 9;   // total words of this round (fixed part + tables)
constexpr int RD_CTRL = 10;   // [5][2] control words of the gate on bit J
constexpr int RD_M = 20;      // [5][4] float2 fused matrices m00 m01 m10 m11      (device filled)
constexpr int RD_C = 60;      // float2 CTA scalar                                  (device filled)
constexpr int RD_F = 64;      // [13][2] float2 per-tile-bit diagonal factors      (device filled)
constexpr int RD_FIXED = 120;
// two-tile-bit table entry
constexpr int TT_A = 0, TT_B = 1, TT_W = 4, TT_WORDS = 12;  // W: d00 d01 d10 d11 (A is the first index)
// control words:  w0 = cmask | cwant << 8 | n_thread_ctrl << 16 ;  w1 = p0 | pol0 << 7 | p1 << 8 | pol1 << 15
// dense shared-memory op (SUB_SMEM_DENSE): one 16-word record after the sub-pass header
constexpr int OP_WORDS = 16;
constexpr int O_CODE = 0, O_A = 1, O_B = 2, O_MAT = 3, O_AUX0 = 4, O_AUX1 = 5, O_AUX2 = 6, O_AUX3 = 7;
constexpr int OP_DENSE = 5;  // AUX0 = k, tile-bit indices of the gate qubits in A,B,AUX1,AUX2; MAT -> 2^k x 2^k
// fill records (device prologue):  [dst word offset, kind, count, 0] + count x [mat_off, form | stride << 8, p, q]
constexpr int FK_SCALAR = 0, FK_PAIR = 1, FK_TABLE = 2, FK_MATRIX = 3;
enum FillForm : int {
  FF_D1 = 0,         // pair  (m[0], m[st])                          1q diagonal on the slot's bit
  FF_D2_FIRST = 1,   // pair  (m[x*st], m[(2+x)*st]),   x = bit p    slot bit is the gate's first qubit
  FF_D2_SECOND = 2,  // pair  (m[2x*st], m[(2x+1)*st]), x = bit p    slot bit is the gate's second qubit
  FF_S1 = 3,         // scalar m[x*st], x = bit p
  FF_S2 = 4,         // scalar m[(2x+y)*st], x = bit p, y = bit q
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

// ---- packed FP32x2 (Blackwell FFMA2): a complex number is one 64-bit register pair -------------
// ptxas folds the lane swap / scalar broadcast / per-lane negation into operand modifiers
// (R.F32x2.LO_HI, R.F32, -R.F32x2.HI_LO.NP), so  acc += m * x  is 2 packed instructions.
#if defined(__CUDACC__)
typedef unsigned long long c64;
__device__ __forceinline__ c64 c64_pack(float lo, float hi) {
  c64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float2 c64_unpack(c64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ c64 c64_fma(c64 a, c64 b, c64 c) {
  c64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ c64 c64_mul(c64 a, c64 b) {
  c64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// a 2x2 complex matrix prepared for packed application: per entry (re, re) and (-im, im)
struct PackedM {
  c64 r00, i00, r01, i01, r10, i10, r11, i11;
};
__device__ __forceinline__ PackedM pack_matrix(float2 m00, float2 m01, float2 m10, float2 m11) {
  PackedM p;
  p.r00 = c64_pack(m00.x, m00.x); p.i00 = c64_pack(-m00.y, m00.y);
  p.r01 = c64_pack(m01.x, m01.x); p.i01 = c64_pack(-m01.y, m01.y);
  p.r10 = c64_pack(m10.x, m10.x); p.i10 = c64_pack(-m10.y, m10.y);
  p.r11 = c64_pack(m11.x, m11.x); p.i11 = c64_pack(-m11.y, m11.y);
  return p;
}
// (b0, b1) = M (x, y)
__device__ __forceinline__ void packed_2x2(const PackedM& p, float2 x, float2 y, float2& b0, float2& b1) {
  const c64 X = c64_pack(x.x, x.y), Xs = c64_pack(x.y, x.x), Y = c64_pack(y.x, y.y), Ys = c64_pack(y.y, y.x);
  c64 r0 = c64_mul(p.r00, X);
  c64 r1 = c64_mul(p.r10, X);
  r0 = c64_fma(p.i00, Xs, r0);
  r1 = c64_fma(p.i10, Xs, r1);
  r0 = c64_fma(p.r01, Y, r0);
  r1 = c64_fma(p.r11, Y, r1);
  r0 = c64_fma(p.i01, Ys, r0);
  r1 = c64_fma(p.i11, Ys, r1);
  b0 = c64_unpack(r0);
  b1 = c64_unpack(r1);
}
#endif

// XOR-fold swizzle of the low 4 bits (8-byte words -> 16 bank pairs per half warp)
TCB_DEV int fold_hi(int t) { return ((t >> 4) ^ (t >> 8) ^ (t >> 12)) & 15; }
TCB_DEV int swz(int t) { return (t & ~15) | ((t ^ fold_hi(t)) & 15); }

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

// ---- CTA prologue: fill records -----------------------------------------------------
// Every source of a fill record is turned into a 2x2 complex matrix E (scalar s -> s*1,
// diagonal pair -> diag(d0, d1), dense 1q gate -> M, two-tile-bit table -> its 4 entries as is),
// and a record's value is the ordered product  E_{count-1} ... E_1 E_0  (later gates multiply from
// the left; diagonal kinds commute).  On the GPU a warp owns a record, lane i loads source i and
// the product is a shuffle tree (pass_kernel.cu); the emulator runs the same loader serially.
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
// load source `src` (4 words) of a record of kind `kind`
TCB_DEV Mat2 load_fill_source(const int32_t* src, int kind, const float2* __restrict__ gates,
                              uint64_t cta_bits) {
  const float2* m = gates + src[0];
  const int form = src[1] & 0xff, st = src[1] >> 8;
  const int x = (int)((cta_bits >> src[2]) & 1ull), y = (int)((cta_bits >> src[3]) & 1ull);
  Mat2 e = mat2_identity();
  if (kind == FK_SCALAR) {
    e.a = (form == FF_S1) ? m[x * st] : m[(2 * x + y) * st];
    e.d = e.a;
  } else if (kind == FK_PAIR) {
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
  if (kind == FK_SCALAR) {
    dst[0] = r.a;
  } else if (kind == FK_PAIR) {
    dst[0] = r.a;
    dst[1] = r.d;
  } else {
    dst[0] = r.a;
    dst[1] = r.b;
    dst[2] = r.c;
    dst[3] = r.d;
  }
}
// serial reference (emulator; also the GPU fallback for records longer than a warp)
TCB_DEV void run_fill_record(int32_t* prog, int rec_off, const float2* __restrict__ gates,
                             uint64_t cta_bits) {
  const int32_t* rec = prog + rec_off;
  const int kind = rec[1], count = rec[2];
  Mat2 acc = mat2_identity();
  for (int i = 0; i < count; ++i) acc = mat2_mul(load_fill_source(rec + 4 + 4 * i, kind, gates, cta_bits), acc);
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
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) a[i] = cmul(a[i], ((i >> J) & 1) ? u1 : u0);
}

// diagonal table on two register bits j, k (both runtime; rare)
template <int R>
TCB_DEV void apply_pairdiag(float2 (&a)[1 << R], int j, int k, float2 d00, float2 d01, float2 d10,
                            float2 d11) {
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    const bool xj = (i >> j) & 1, xk = (i >> k) & 1;
    a[i] = cmul(a[i], xj ? csel(xk, d11, d10) : csel(xk, d01, d00));
  }
}

TCB_DEV float2 lds2(const int32_t* p) { return *reinterpret_cast<const float2*>(p); }

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

template <int R>
struct SubState {
  float2 a[1 << R];
  float2 pc;    // pending thread-constant factor (folded into the next gate matrix)
  bool pc_set;
};

// register bit J (compile-time) of one round: its share of the diagonal part, then its gate.
//   tt : cursor into the round's two-tile-bit tables (section J follows the thread-constant ones)
template <int R, int J>
TCB_DEV void round_bit(SubState<R>& s, const int32_t* rd, int flags, int fmask, int rtbJ, int tbase,
                       uint64_t gidx, const int32_t*& tt) {
  if constexpr (J < R) {
    // ---- diagonal factor of this register bit: F_b of its tile bit x tables with a thread-constant partner
    float2 f0 = c_one(), f1 = c_one();
    bool have = false;
    if (flags & (1 << 16)) {
      if ((fmask >> rtbJ) & 1) {
        f0 = lds2(rd + RD_F + 4 * rtbJ);
        f1 = lds2(rd + RD_F + 4 * rtbJ + 2);
        have = true;
      }
      const int nj = rd[RD_NRJ + J];
      TCB_NOUNROLL
      for (int e = 0; e < nj; ++e, tt += TT_WORDS) {
        const int xb = (tbase >> tt[TT_B]) & 1;
        f0 = cmul(f0, lds2(tt + TT_W + 2 * xb));
        f1 = cmul(f1, lds2(tt + TT_W + 2 * (2 + xb)));
        have = true;
      }
    }
    const bool gate = (flags >> J) & 1;
    if (!gate && !have) return;
    if (s.pc_set) {  // any full-width operation on the amplitudes can carry the thread scalar
      f0 = cmul(f0, s.pc);
      f1 = cmul(f1, s.pc);
      s.pc_set = false;
      have = true;
    }
    if (!gate) {
      diag_on<R, J>(s.a, f0, f1);
      return;
    }
    const int cw0 = rd[RD_CTRL + 2 * J], cw1 = rd[RD_CTRL + 2 * J + 1];
    const int ntc = (cw0 >> 16) & 3;
    bool active = true;
    if (ntc > 0) active = (int)((gidx >> (cw1 & 0x7f)) & 1ull) == ((cw1 >> 7) & 1);
    if (ntc > 1) active = active && ((int)((gidx >> ((cw1 >> 8) & 0x7f)) & 1ull) == ((cw1 >> 15) & 1));
    float2 m00 = c_one(), m01 = make_float2(0.f, 0.f), m10 = m01, m11 = c_one();
    if (active) {
      m00 = lds2(rd + RD_M + 8 * J);
      m01 = lds2(rd + RD_M + 8 * J + 2);
      m10 = lds2(rd + RD_M + 8 * J + 4);
      m11 = lds2(rd + RD_M + 8 * J + 6);
    }
    if (!((flags >> (8 + J)) & 1)) {
      // fold the diagonal factor into the matrix columns:  M' = M diag(f0, f1)
      m00 = cmul(m00, f0);
      m10 = cmul(m10, f0);
      m01 = cmul(m01, f1);
      m11 = cmul(m11, f1);
      gate_on<R, J, false>(s.a, m00, m01, m10, m11, 0, 0);
    } else {
      // register-resident controls: the factor acts on every pair, the gate only on selected ones
      if (have) diag_on<R, J>(s.a, f0, f1);
      if (active) gate_on<R, J, true>(s.a, m00, m01, m10, m11, cw0 & 0xff, (cw0 >> 8) & 0xff);
    }
  }
}

// one thread's share of a register sub-pass.
//   tile : shared-memory tile (swizzled), sp : sub-pass header, group : which 2^R-amplitude group
//   this thread owns, cta_bits : CTA-constant flat-index bits (tile base | index_base),
//   hi_flat : optional table of tile_to_flat(h << L)
template <int R>
TCB_DEV void run_reg_subpass(float2* tile, const int32_t* hdr, const int32_t* sp, int group,
                             uint64_t cta_bits, const uint64_t* hi_flat = nullptr) {
  static_assert(R >= 1 && R <= 5, "register bits");
  const int T = hdr[H_T];
  int tbase = 0;
  for (int b = 0; b < T - R; ++b) tbase |= ((group >> b) & 1) << sp[S_GRPBITS + b];
  uint64_t gidx;
  if (hi_flat != nullptr) {
    const int L = hdr[H_L];
    gidx = cta_bits | hi_flat[tbase >> L] | (uint64_t)(tbase & ((1 << L) - 1));
  } else {
    gidx = cta_bits | tile_to_flat(tbase, hdr);
  }

  // shared-memory address of amplitude i:  tbase and the register offsets have disjoint bits and
  // the swizzle is GF(2)-linear, so  swz(tbase | off_i) = swz(tbase) ^ XOR_j swz(1 << r_j)
  int rsw[R];
  TCB_UNROLL
  for (int j = 0; j < R; ++j) rsw[j] = swz(1 << sp[S_REGBITS + j]);
  const int sbase = swz(tbase);

  SubState<R> s;
  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    s.a[i] = tile[ad];
  }
  s.pc = c_one();
  s.pc_set = false;

  const int nrounds = sp[S_NROUNDS];
  const int32_t* rd = sp + SUB_HDR_WORDS;
  TCB_NOUNROLL
  for (int r = 0; r < nrounds; ++r) {
    const int flags = rd[RD_FLAGS];
    const int fmask = rd[RD_FMASK];
    const int32_t* tt = rd + RD_FIXED;
    if (flags & (1 << 16)) {
      // ---- thread-constant part of the diagonal: C x F_b of the 8 thread bits x tables on two of them
      float2 c = lds2(rd + RD_C);
      TCB_NOUNROLL
      for (int b = 0; b < T - R; ++b) {
        const int tb = sp[S_GRPBITS + b];
        if ((fmask >> tb) & 1) c = cmul(c, lds2(rd + RD_F + 4 * tb + 2 * ((group >> b) & 1)));
      }
      const int nnn = rd[RD_NNN];
      TCB_NOUNROLL
      for (int e = 0; e < nnn; ++e, tt += TT_WORDS) {
        const int xa = (tbase >> tt[TT_A]) & 1, xb = (tbase >> tt[TT_B]) & 1;
        c = cmul(c, lds2(tt + TT_W + 2 * (2 * xa + xb)));
      }
      s.pc = s.pc_set ? cmul(s.pc, c) : c;
      s.pc_set = true;
      // tables on two register bits act on the amplitudes directly; they belong to the diagonal
      // part, i.e. BEFORE this round's gates (their records sit after the per-bit sections)
      int nrj = 0;
      TCB_UNROLL
      for (int j = 0; j < R; ++j) nrj += rd[RD_NRJ + j];
      const int32_t* trr = tt + TT_WORDS * nrj;
      const int nrr = rd[RD_NRR];
      TCB_NOUNROLL
      for (int e = 0; e < nrr; ++e, trr += TT_WORDS)
        apply_pairdiag<R>(s.a, trr[TT_A], trr[TT_B], lds2(trr + TT_W), lds2(trr + TT_W + 2), lds2(trr + TT_W + 4),
                          lds2(trr + TT_W + 6));
    }
    // ---- per register bit (compile-time J): its diagonal factor, then its fused 2x2 ----
    round_bit<R, 0>(s, rd, flags, fmask, sp[S_REGBITS + 0], tbase, gidx, tt);
    round_bit<R, 1>(s, rd, flags, fmask, sp[S_REGBITS + 1], tbase, gidx, tt);
    round_bit<R, 2>(s, rd, flags, fmask, sp[S_REGBITS + 2], tbase, gidx, tt);
    round_bit<R, 3>(s, rd, flags, fmask, sp[S_REGBITS + 3], tbase, gidx, tt);
    round_bit<R, 4>(s, rd, flags, fmask, sp[S_REGBITS + 4], tbase, gidx, tt);
    rd += rd[RD_WORDS];
  }
  if (s.pc_set) diag_on<R, 0>(s.a, s.pc, s.pc);

  TCB_UNROLL
  for (int i = 0; i < (1 << R); ++i) {
    int ad = sbase;
    TCB_UNROLL
    for (int j = 0; j < R; ++j)
      if ((i >> j) & 1) ad ^= rsw[j];
    tile[ad] = s.a[i];
  }
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

// sv_kernels.cu — unfused statevector kernels: per-gate dense / diagonal application
// (the literal R7/R9 order of tensorcircuit/cons.py:429-463), reductions for
// expectations (tensorcircuit/circuit.py:899-902 without a materialised bra), the
// adjoint-mode gate-gradient reduction and the pack/unpack halves of a qubit swap.
// All are HBM-bound streaming kernels: 128-bit accesses, grid = multiple of the SM count.
#include <vector>

#include "common.cuh"
#include "pass_core.cuh"  // cmul / cfma

namespace tcb {

static inline unsigned grid_for(uint64_t work_items, int threads, int per_sm = 8) {
  uint64_t blocks = (work_items + threads - 1) / threads;
  const uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// ---------------------------------------------------------------------------------
__global__ void init_zero_kernel(float4* state, uint64_t nvec_per_state, uint64_t nvec_total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec_total; i += stride) {
    const bool first = (i % nvec_per_state) == 0;
    state[i] = make_float4(first ? 1.f : 0.f, 0.f, 0.f, 0.f);
  }
}

int launch_init_zero(void* state, int nbits, int64_t batch, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_init_zero: nbits=%d out of range", nbits);
  const uint64_t nvec = 1ull << (nbits - 1);
  const uint64_t total = nvec * (uint64_t)batch;
  init_zero_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(state), nvec,
                                                             total);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// dense k-qubit gate, one thread per group of 2^K amplitudes
struct BitList {
  int pos[8];     // gate qubit i -> flat bit position (qubit 0 = matrix MSB)
  int sorted[8];  // ascending
};

template <int K>
__global__ void __launch_bounds__(256)
dense_kernel(float2* state, int nbits, uint64_t groups_per_state, BitList bl,
             const float2* __restrict__ mat, long long mat_bstride) {
  __shared__ float2 sm[(1 << K) * (1 << K)];
  const unsigned b = blockIdx.y;
  for (int e = threadIdx.x; e < (1 << K) * (1 << K); e += blockDim.x)
    sm[e] = mat[(size_t)b * mat_bstride + e];
  __syncthreads();
  float2* st = state + ((size_t)b << nbits);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < groups_per_state;
       gi += stride) {
    uint64_t base = gi;
#pragma unroll
    for (int i = 0; i < K; ++i) base = insert_zero(base, bl.sorted[i]);
    float2 v[1 << K];
    uint64_t off[1 << K];
#pragma unroll
    for (int c = 0; c < (1 << K); ++c) {
      uint64_t o = 0;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((c >> (K - 1 - i)) & 1) o |= 1ull << bl.pos[i];
      off[c] = o;
      v[c] = st[base | o];
    }
#pragma unroll
    for (int r = 0; r < (1 << K); ++r) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < (1 << K); ++c) acc = cfma(sm[r * (1 << K) + c], v[c], acc);
      st[base | off[r]] = acc;
    }
  }
}

int launch_dense(void* state, int nbits, int64_t batch, const int* bitpos, int k, const void* mat,
                 int64_t mat_bstride, cudaStream_t stream) {
  TCB_REQUIRE(k >= 1 && k <= 5, "tcb_sv_apply_dense: k=%d unsupported (1..5)", k);
  TCB_REQUIRE(nbits >= k && nbits <= 40, "tcb_sv_apply_dense: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_apply_dense: batch=%lld", (long long)batch);
  BitList bl;
  for (int i = 0; i < k; ++i) {
    TCB_REQUIRE(bitpos[i] >= 0 && bitpos[i] < nbits, "tcb_sv_apply_dense: bit %d out of range",
                bitpos[i]);
    for (int j = 0; j < i; ++j)
      TCB_REQUIRE(bitpos[i] != bitpos[j], "tcb_sv_apply_dense: duplicate bit %d", bitpos[i]);
    bl.pos[i] = bitpos[i];
    bl.sorted[i] = bitpos[i];
  }
  for (int i = 1; i < k; ++i)
    for (int j = i; j > 0 && bl.sorted[j - 1] > bl.sorted[j]; --j) {
      int t = bl.sorted[j];
      bl.sorted[j] = bl.sorted[j - 1];
      bl.sorted[j - 1] = t;
    }
  const uint64_t gps = 1ull << (nbits - k);
  dim3 grid(grid_for(gps, 256), (unsigned)batch);
  float2* st = reinterpret_cast<float2*>(state);
  const float2* m = reinterpret_cast<const float2*>(mat);
  switch (k) {
    case 1: dense_kernel<1><<<grid, 256, 0, stream>>>(st, nbits, gps, bl, m, mat_bstride); break;
    case 2: dense_kernel<2><<<grid, 256, 0, stream>>>(st, nbits, gps, bl, m, mat_bstride); break;
    case 3: dense_kernel<3><<<grid, 256, 0, stream>>>(st, nbits, gps, bl, m, mat_bstride); break;
    case 4: dense_kernel<4><<<grid, 256, 0, stream>>>(st, nbits, gps, bl, m, mat_bstride); break;
    case 5: dense_kernel<5><<<grid, 256, 0, stream>>>(st, nbits, gps, bl, m, mat_bstride); break;
  }
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// diagonal k-qubit gate: one thread per amplitude pair (float4)
struct DiagBits {
  int pos[8];
  int k;
};

__global__ void __launch_bounds__(256)
diag_kernel(float4* state, int nbits, uint64_t nvec_per_state, uint64_t nvec_total, DiagBits db,
            const float2* __restrict__ diag, long long dstride, long long mat_bstride,
            unsigned long long index_base) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec_total; i += stride) {
    const uint64_t b = i / nvec_per_state;
    const uint64_t x0 = ((i % nvec_per_state) << 1) | index_base;
    const float2* d = diag + (size_t)b * mat_bstride;
    int i0 = 0, i1 = 0;
    for (int q = 0; q < db.k; ++q) {
      const int p = db.pos[q];
      const int b0 = (int)((x0 >> p) & 1ull);
      const int b1 = (p == 0) ? 1 : b0;
      i0 = (i0 << 1) | b0;
      i1 = (i1 << 1) | b1;
    }
    const float2 d0 = d[(size_t)i0 * dstride], d1 = d[(size_t)i1 * dstride];
    float4 v = state[i];
    const float2 a = cmul(make_float2(v.x, v.y), d0), c = cmul(make_float2(v.z, v.w), d1);
    state[i] = make_float4(a.x, a.y, c.x, c.y);
  }
}

int launch_diag(void* state, int nbits, int64_t batch, const int* bitpos, int k, const void* diag,
                int64_t dstride, int64_t mat_bstride, uint64_t index_base, cudaStream_t stream) {
  TCB_REQUIRE(k >= 1 && k <= 8, "tcb_sv_apply_diag: k=%d unsupported (1..8)", k);
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_apply_diag: nbits=%d", nbits);
  DiagBits db;
  db.k = k;
  for (int i = 0; i < k; ++i) {
    TCB_REQUIRE(bitpos[i] >= 0 && bitpos[i] < 64, "tcb_sv_apply_diag: bit %d out of range", bitpos[i]);
    db.pos[i] = bitpos[i];
  }
  const uint64_t nvec = 1ull << (nbits - 1);
  const uint64_t total = nvec * (uint64_t)batch;
  diag_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      reinterpret_cast<float4*>(state), nbits, nvec, total, db, reinterpret_cast<const float2*>(diag),
      (long long)dstride, (long long)mat_bstride, (unsigned long long)index_base);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// product state: state[x] = prod_p v[p][bit p of (x | index_base)]   (one write pass, no read).
// The gates every qubit sees before its first multi-qubit gate act on |0> alone, so the host
// folds them into per-qubit 2-vectors (the reference's _merge_single_gates does the same to its
// input nodes, tensorcircuit/cons.py:298-374) and the first passes over the state disappear.
// Each thread builds 32 amplitudes by doubling over 5 index bits; for nbits >= 10 these are bit 0 and
// bits 6..9 while the lane supplies bits 1..5, so every store instruction of a warp writes 512
// contiguous bytes (the naive "32 consecutive amplitudes per thread" version ran at 1.7 TB/s).
__global__ void __launch_bounds__(256)
init_product_kernel(float4* __restrict__ state, const float2* __restrict__ vecs, int nbits, int total_bits,
                    unsigned long long index_base, uint64_t ngroups) {
  __shared__ float2 sv[64 * 2];
  for (int i = threadIdx.x; i < 2 * total_bits; i += blockDim.x) sv[i] = vecs[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  if (nbits >= 10) {
    const int own[5] = {0, 6, 7, 8, 9};
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
      // g = (block of 1024 amplitudes) * 32 + lane;  fixed bits: lane -> 1..5, block -> 10..
      const unsigned long long fixed = ((g >> 5) << 10) | ((g & 31ull) << 1) | index_base;
      float2 c = make_float2(1.f, 0.f);
      for (int p = 1; p < total_bits; ++p)
        if (p < 6 || p > 9) c = cmul(c, sv[2 * p + (int)((fixed >> p) & 1ull)]);
      float2 a[32];
      a[0] = c;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const float2 v0 = sv[2 * own[b]], v1 = sv[2 * own[b] + 1];
#pragma unroll
        for (int i = (1 << b) - 1; i >= 0; --i) {
          a[i | (1 << b)] = cmul(a[i], v1);
          a[i] = cmul(a[i], v0);
        }
      }
      float4* dst = state + ((fixed & ~index_base) >> 1);
#pragma unroll
      for (int j = 0; j < 16; ++j)  // local index bits 1..4 of `a` are tile bits 6..9: 64 amplitudes = 32 float4 apart
        dst[(size_t)j * 32] = make_float4(a[2 * j].x, a[2 * j].y, a[2 * j + 1].x, a[2 * j + 1].y);
    }
    return;
  }
  const int nlow = nbits < 5 ? nbits : 5;  // small states: 2^nlow consecutive amplitudes per thread
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const unsigned long long hi = (g << nlow) | index_base;
    float2 c = make_float2(1.f, 0.f);
    for (int p = nlow; p < total_bits; ++p) c = cmul(c, sv[2 * p + (int)((hi >> p) & 1ull)]);
    float2 a[32];
    a[0] = c;
    for (int p = 0; p < nlow; ++p) {
      const float2 v0 = sv[2 * p], v1 = sv[2 * p + 1];
      for (int i = (1 << p) - 1; i >= 0; --i) {
        a[i | (1 << p)] = cmul(a[i], v1);
        a[i] = cmul(a[i], v0);
      }
    }
    float2* dst = reinterpret_cast<float2*>(state) + (g << nlow);
    for (int i = 0; i < (1 << nlow); ++i) dst[i] = a[i];
  }
}

int launch_init_product(void* state, int nbits, const void* vecs, int total_bits, uint64_t index_base,
                        cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_init_product: nbits=%d", nbits);
  TCB_REQUIRE(total_bits >= nbits && total_bits <= 64, "tcb_sv_init_product: total_bits=%d", total_bits);
  const int nlow = nbits < 5 ? nbits : 5;
  const uint64_t ngroups = 1ull << (nbits - nlow);
  init_product_kernel<<<grid_for(ngroups, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(state),
                                                                  reinterpret_cast<const float2*>(vecs), nbits,
                                                                  total_bits, (unsigned long long)index_base, ngroups);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// Z-string expectations: one read of the state for all terms.
//
// <Z_S> = sum_x (-1)^{popc(x & m)} |psi_x|^2 is a Walsh-Hadamard coefficient of the probability
// vector.  Each thread holds 32 probabilities whose indices differ in 5 known bits (bit 0 and
// bits 9..12 of x); one in-register WHT (80 adds) turns them into the 32 possible signed sums,
// so every term costs ONE coefficient lookup + the sign of the remaining bits per 32 amplitudes
// instead of 32 sign evaluations.  Per-thread float partial sums per term live in shared memory
// ([term][thread], conflict-free); every EZ_FLUSH iterations they are warp-reduced into double
// accumulators, so the shuffle tree is off the streaming path (it was 80% of the old kernel).
constexpr int EZ_THREADS = 256;
constexpr int EZ_VEC = 16;  // float4 loads per thread per iteration (32 amplitudes)
constexpr int EZ_MAX_TERMS = 1024;
constexpr int EZ_TERMS_PER_LAUNCH = 64;  // (160 terms per read, one CTA per SM: measured slower end to end)
constexpr int EZ_FLUSH = 32;

__global__ void __launch_bounds__(EZ_THREADS)
expect_z_kernel(const float4* __restrict__ state, uint64_t nvec_per_state,
                const unsigned long long* __restrict__ zmasks, int nterms, int out_stride,
                unsigned long long index_base, double* out) {
  extern __shared__ __align__(16) unsigned char ez_smem[];
  float* coef = reinterpret_cast<float*>(ez_smem);               // [32][EZ_THREADS]
  float* part = coef + 32 * EZ_THREADS;                          // [nterms][EZ_THREADS]
  double* acc = reinterpret_cast<double*>(part + nterms * EZ_THREADS);  // [warps][nterms]
  unsigned long long* smask = reinterpret_cast<unsigned long long*>(acc + (EZ_THREADS / 32) * nterms);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = EZ_THREADS / 32;
  for (int e = tid; e < nwarps * nterms; e += EZ_THREADS) acc[e] = 0.0;
  for (int t = tid; t < nterms; t += EZ_THREADS) smask[t] = zmasks[t];
  for (int t = 0; t < nterms; ++t) part[t * EZ_THREADS + tid] = 0.f;
  __syncthreads();
  const unsigned b = blockIdx.y;
  const float4* st = state + (size_t)b * nvec_per_state;
  double* my = acc + warp * nterms;
  const uint64_t chunk = (uint64_t)EZ_THREADS * EZ_VEC;
  auto flush = [&]() {
    for (int t = 0; t < nterms; ++t) {
      float val = part[t * EZ_THREADS + tid];
      part[t * EZ_THREADS + tid] = 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      if (lane == 0) my[t] += (double)val;
    }
  };
  int it = 0;
  // group bits of x: bit 0 (pair inside a float4) and bits 9..12 (u); everything else is "rest"
  for (uint64_t v0 = (uint64_t)blockIdx.x * chunk; v0 < nvec_per_state;
       v0 += (uint64_t)gridDim.x * chunk) {
    float p[2 * EZ_VEC];
#pragma unroll
    for (int u = 0; u < EZ_VEC; ++u) {
      const uint64_t v = v0 + (uint64_t)u * EZ_THREADS + tid;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nvec_per_state) a = ldg_stream(st + v);
      p[2 * u] = a.x * a.x + a.y * a.y;
      p[2 * u + 1] = a.z * a.z + a.w * a.w;
    }
    // in-register Walsh-Hadamard transform over the 5 group bits (w = e | u << 1)
#pragma unroll
    for (int sft = 0; sft < 5; ++sft) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (!((i >> sft) & 1)) {
          const float x = p[i], y = p[i | (1 << sft)];
          p[i] = x + y;
          p[i | (1 << sft)] = x - y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) coef[i * EZ_THREADS + tid] = p[i];
    // (each thread only reads back its own column: no barrier needed)
    const unsigned long long xrest = ((v0 + (uint64_t)tid) << 1) | index_base;
#pragma unroll 4
    for (int t = 0; t < nterms; ++t) {
      const unsigned long long m = smask[t];
      const int c = (int)(m & 1ull) | (int)(((m >> 9) & 15ull) << 1);
      const float val = coef[c * EZ_THREADS + tid];
      // xrest has zeros at the 5 group bits, so they are not counted twice
      const float sgn = (__popcll(xrest & m) & 1) ? -val : val;
      part[t * EZ_THREADS + tid] += sgn;
    }
    if (++it == EZ_FLUSH) {
      flush();
      it = 0;
    }
  }
  flush();
  __syncthreads();
  for (int t = tid; t < nterms; t += EZ_THREADS) {
    double sum = 0.0;
    for (int w = 0; w < nwarps; ++w) sum += acc[w * nterms + t];
    atomicAdd(out + (size_t)b * out_stride + t, sum);
  }
}

int launch_expect_z(const void* state, int nbits, int64_t batch, const uint64_t* zmasks, int nterms,
                    uint64_t index_base, double* out, cudaStream_t stream) {
  TCB_REQUIRE(nterms >= 1 && nterms <= EZ_MAX_TERMS, "tcb_sv_expect_z: nterms=%d out of range", nterms);
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_expect_z: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_expect_z: batch=%lld", (long long)batch);
  TCB_REQUIRE((index_base & ((1ull << nbits) - 1ull)) == 0, "tcb_sv_expect_z: index_base overlaps local bits");
  const uint64_t nvec = 1ull << (nbits - 1);
  const uint64_t chunk = (uint64_t)EZ_THREADS * EZ_VEC;
  uint64_t gx = (nvec + chunk - 1) / chunk;
  const uint64_t cap = (uint64_t)sm_count() * 2;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)batch);
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(expect_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        110 * 1024));
    attr_set = true;
  }
  // the per-thread partial sums bound the terms of one launch; more terms = more reads of the state
  for (int t0 = 0; t0 < nterms; t0 += EZ_TERMS_PER_LAUNCH) {
    const int nt = nterms - t0 < EZ_TERMS_PER_LAUNCH ? nterms - t0 : EZ_TERMS_PER_LAUNCH;
    const size_t smem = 32 * EZ_THREADS * sizeof(float) + (size_t)nt * EZ_THREADS * sizeof(float) +
                        sizeof(double) * (EZ_THREADS / 32) * nt + sizeof(unsigned long long) * nt;
    expect_z_kernel<<<grid, EZ_THREADS, smem, stream>>>(
        reinterpret_cast<const float4*>(state), nvec,
        reinterpret_cast<const unsigned long long*>(zmasks) + t0, nt, nterms, (unsigned long long)index_base,
        out + t0);
    TCB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// block-level complex reduction into double atomics
__device__ __forceinline__ void block_reduce_add2(double re, double im, double* out) {
  __shared__ double sre[32], sim[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sre[warp] = re;
    sim[warp] = im;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    re = lane < nw ? sre[lane] : 0.0;
    im = lane < nw ? sim[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) {
      atomicAdd(out, re);
      atomicAdd(out + 1, im);
    }
  }
}

// general Pauli string: sum_x conj(psi[x ^ xm]) * i^ny * (-1)^popc(x & zm) * psi[x]
__global__ void __launch_bounds__(256)
expect_pauli_kernel(const float2* __restrict__ state, uint64_t n_per_state, unsigned long long xmask,
                    unsigned long long zmask, int ny, unsigned long long index_base, double* out) {
  const unsigned b = blockIdx.y;
  const float2* st = state + (size_t)b * n_per_state;
  float re = 0.f, im = 0.f;
  double dre = 0.0, dim_ = 0.0;
  int cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_per_state; x += stride) {
    const float2 a = st[x];
    const float2 c = st[x ^ xmask];
    const bool odd = __popcll((x | index_base) & zmask) & 1;
    // conj(c) * a
    float pr = c.x * a.x + c.y * a.y;
    float pi = c.x * a.y - c.y * a.x;
    if (odd) {
      pr = -pr;
      pi = -pi;
    }
    re += pr;
    im += pi;
    if (++cnt == 64) {
      dre += re;
      dim_ += im;
      re = im = 0.f;
      cnt = 0;
    }
  }
  dre += re;
  dim_ += im;
  // multiply by i^ny
  double rr = dre, ii = dim_;
  switch (ny & 3) {
    case 1: rr = -dim_; ii = dre; break;
    case 2: rr = -dre; ii = -dim_; break;
    case 3: rr = dim_; ii = -dre; break;
    default: break;
  }
  block_reduce_add2(rr, ii, out + 2 * (size_t)b);
}

int launch_expect_pauli(const void* state, int nbits, int64_t batch, uint64_t xmask, uint64_t zmask,
                        int ny, uint64_t index_base, double* out, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_expect_pauli: nbits=%d", nbits);
  TCB_REQUIRE(xmask < (1ull << nbits), "tcb_sv_expect_pauli: xmask touches non-local bits");
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_expect_pauli: batch=%lld", (long long)batch);
  const uint64_t n = 1ull << nbits;
  dim3 grid(grid_for(n, 256), (unsigned)batch);
  expect_pauli_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float2*>(state), n, xmask, zmask,
                                                ny, index_base, out);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// <a|b>
__global__ void __launch_bounds__(256)
inner_kernel(const float4* __restrict__ a, const float4* __restrict__ bb, uint64_t nvec_per_state,
             double* out) {
  const unsigned b = blockIdx.y;
  const float4* pa = a + (size_t)b * nvec_per_state;
  const float4* pb = bb + (size_t)b * nvec_per_state;
  float re = 0.f, im = 0.f;
  double dre = 0.0, dim_ = 0.0;
  int cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec_per_state; i += stride) {
    const float4 u = ldg_stream(pa + i), v = ldg_stream(pb + i);
    re += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
    im += u.x * v.y - u.y * v.x + u.z * v.w - u.w * v.z;
    if (++cnt == 32) {
      dre += re;
      dim_ += im;
      re = im = 0.f;
      cnt = 0;
    }
  }
  dre += re;
  dim_ += im;
  block_reduce_add2(dre, dim_, out + 2 * (size_t)b);
}

int launch_inner(const void* a, const void* b, int nbits, int64_t batch, double* out,
                 cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_inner: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_inner: batch=%lld", (long long)batch);
  const uint64_t nvec = 1ull << (nbits - 1);
  dim3 grid(grid_for(nvec, 256), (unsigned)batch);
  inner_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(a),
                                         reinterpret_cast<const float4*>(b), nvec, out);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// Matrix-free Pauli-sum operator  H = sum_t c_t P_t  (what the reference does with a
// COO matrix in templates/measurements.py:156-191 or term by term in quantum.py:2222-2358):
//   (H psi)[i] = sum_t c_t (-1)^popc((i ^ x_t) & z_t) psi[i ^ x_t],   c_t = w_t i^{ny_t}
// Terms arrive sorted by flip mask; every run of equal x_t is one load of psi[i ^ x] and its
// coefficients fold into one complex factor per amplitude before the multiply.  One launch writes
// H psi and / or accumulates <psi|H|psi>; each thread owns eight neighbouring amplitudes (64 contiguous
// bytes per access; flips of the three low bits permute registers) and the sign of a term is one popc for
// the high bits plus a per-term 8-bit table for the low ones, applied as +-1.0f in an FFMA.
constexpr int PS_MAX_TERMS = 1024;

// LB = log2(amplitudes per thread): 3 (64 contiguous bytes per access) or 1 for states of fewer than 8 amplitudes.
template <int LB, bool kWrite, bool kAccum, bool kValue>
__global__ void __launch_bounds__(256)
pauli_sum_kernel(const float4* __restrict__ state, uint64_t nblk_per_state,
                 const unsigned long long* __restrict__ xs, const unsigned long long* __restrict__ zs,
                 const float2* __restrict__ cs, int nterms, unsigned long long index_base,
                 float4* __restrict__ out_state, double* out_value) {
  constexpr int A = 1 << LB;   // amplitudes per thread
  constexpr int V = A / 2;     // float4 per thread
  __shared__ unsigned long long sx[PS_MAX_TERMS], sz[PS_MAX_TERMS];
  __shared__ float2 sc[PS_MAX_TERMS];
  __shared__ unsigned slm[PS_MAX_TERMS];  // bit j: parity of ((j ^ x) & z) over the low LB bits
  for (int t = threadIdx.x; t < nterms; t += blockDim.x) {
    const unsigned long long x = xs[t], z = zs[t];
    sx[t] = x;
    sz[t] = z;
    sc[t] = cs[t];
    unsigned lm = 0;
#pragma unroll
    for (int jj = 0; jj < A; ++jj) lm |= (unsigned)(__popc((jj ^ (unsigned)(x & (A - 1))) & (unsigned)(z & (A - 1))) & 1) << jj;
    slm[t] = lm;
  }
  __syncthreads();
  const unsigned b = blockIdx.y;
  const float4* st = state + (size_t)b * nblk_per_state * V;
  float4* os = kWrite ? out_state + (size_t)b * nblk_per_state * V : nullptr;
  double dre = 0.0, dim_ = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nblk_per_state; p += stride) {
    const unsigned long long ihi = ((p << LB) | index_base) >> LB;
    float2 acc[A];
#pragma unroll
    for (int jj = 0; jj < A; ++jj) acc[jj] = make_float2(0.f, 0.f);
    int t = 0;
    while (t < nterms) {
      const unsigned long long x = sx[t];
      const float4* src = st + ((p ^ (x >> LB)) * V);
      float2 v[A];
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 w = src[q];
        v[2 * q] = make_float2(w.x, w.y);
        v[2 * q + 1] = make_float2(w.z, w.w);
      }
      // v[j] <- v[j ^ (x & (A-1))]: uniform branches (every thread walks the same term list)
#pragma unroll
      for (int bit = 0; bit < LB; ++bit) {
        if ((x >> bit) & 1ull) {
#pragma unroll
          for (int jj = 0; jj < A; ++jj)
            if (!(jj & (1 << bit))) {
              const float2 tmp = v[jj];
              v[jj] = v[jj | (1 << bit)];
              v[jj | (1 << bit)] = tmp;
            }
        }
      }
      float2 d[A], du = make_float2(0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < A; ++jj) d[jj] = make_float2(0.f, 0.f);
      do {
        const unsigned long long z = sz[t];
        const float2 c = sc[t];
        const bool odd = __popcll((ihi ^ (x >> LB)) & (z >> LB)) & 1;
        if ((z & (unsigned long long)(A - 1)) == 0ull) {
          // (uniform branch) no Z on the thread's own bits: one sign for all its amplitudes — the term costs one
          // signed add per THREAD (most terms of a lattice Hamiltonian: every X term, every ZZ bond above bit 2)
          const float sg = odd ? -1.f : 1.f;
          du.x = fmaf(sg, c.x, du.x);
          du.y = fmaf(sg, c.y, du.y);
        } else {
          unsigned m = slm[t];
          if (odd) m = ~m;
#pragma unroll
          for (int jj = 0; jj < A; ++jj) {
            const float sg = __uint_as_float(0x3f800000u | ((m << (31 - jj)) & 0x80000000u));
            d[jj].x = fmaf(sg, c.x, d[jj].x);
            d[jj].y = fmaf(sg, c.y, d[jj].y);
          }
        }
        ++t;
      } while (t < nterms && sx[t] == x);
#pragma unroll
      for (int jj = 0; jj < A; ++jj) {
        const float2 dt = make_float2(d[jj].x + du.x, d[jj].y + du.y);
        acc[jj].x += dt.x * v[jj].x - dt.y * v[jj].y;
        acc[jj].y += dt.x * v[jj].y + dt.y * v[jj].x;
      }
    }
    if (kValue) {
      float re = 0.f, im = 0.f;
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const float4 u = st[p * V + q];  // conj(psi) . (H psi)
        re += u.x * acc[2 * q].x + u.y * acc[2 * q].y + u.z * acc[2 * q + 1].x + u.w * acc[2 * q + 1].y;
        im += u.x * acc[2 * q].y - u.y * acc[2 * q].x + u.z * acc[2 * q + 1].y - u.w * acc[2 * q + 1].x;
      }
      dre += (double)re;
      dim_ += (double)im;
    }
    if (kWrite) {
#pragma unroll
      for (int q = 0; q < V; ++q) {
        float4 r = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        if (kAccum) {
          const float4 o = os[p * V + q];
          r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
        }
        os[p * V + q] = r;
      }
    }
  }
  if (kValue) block_reduce_add2(dre, dim_, out_value + 2 * (size_t)b);
}

template <int LB>
static int pauli_sum_dispatch(const float4* st, int nbits, int64_t batch, const unsigned long long* xs,
                              const unsigned long long* zs, const float2* cs, int nterms, uint64_t index_base,
                              float4* os, bool acc, double* out_value, cudaStream_t stream) {
  const uint64_t nblk = 1ull << (nbits - LB);
  dim3 grid(grid_for(nblk, 256), (unsigned)batch);
  int first = 0;
  do {  // more terms than one shared-memory table: further launches accumulate
    const int cnt = nterms - first < PS_MAX_TERMS ? nterms - first : PS_MAX_TERMS;
#define TCB_PS_LAUNCH(W, AC, VAL) \
  pauli_sum_kernel<LB, W, AC, VAL><<<grid, 256, 0, stream>>>(st, nblk, xs + first, zs + first, cs + first, cnt, index_base, os, out_value)
    if (os && out_value) {
      if (acc) TCB_PS_LAUNCH(true, true, true); else TCB_PS_LAUNCH(true, false, true);
    } else if (os) {
      if (acc) TCB_PS_LAUNCH(true, true, false); else TCB_PS_LAUNCH(true, false, false);
    } else {
      TCB_PS_LAUNCH(false, false, true);
    }
#undef TCB_PS_LAUNCH
    TCB_CHECK_CUDA(cudaGetLastError());
    first += cnt;
    acc = true;
  } while (first < nterms);
  return 0;
}

int launch_pauli_sum(const void* state, int nbits, int64_t batch, const uint64_t* xmask,
                     const uint64_t* zmask, const void* coef, int nterms, uint64_t index_base,
                     void* out_state, int accumulate, double* out_value, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_pauli_sum: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_pauli_sum: batch=%lld", (long long)batch);
  TCB_REQUIRE(nterms >= 0, "tcb_sv_pauli_sum: nterms=%d", nterms);
  TCB_REQUIRE(out_state != nullptr || out_value != nullptr, "tcb_sv_pauli_sum: nothing to compute");
  TCB_REQUIRE(out_state != state, "tcb_sv_pauli_sum: out_state must not alias state");
  const float4* st = reinterpret_cast<const float4*>(state);
  float4* os = reinterpret_cast<float4*>(out_state);
  const unsigned long long* xs = reinterpret_cast<const unsigned long long*>(xmask);
  const unsigned long long* zs = reinterpret_cast<const unsigned long long*>(zmask);
  const float2* cs = reinterpret_cast<const float2*>(coef);
  if (nbits >= 3)
    return pauli_sum_dispatch<3>(st, nbits, batch, xs, zs, cs, nterms, index_base, os, accumulate != 0, out_value, stream);
  return pauli_sum_dispatch<1>(st, nbits, batch, xs, zs, cs, nterms, index_base, os, accumulate != 0, out_value, stream);
}

// ---------------------------------------------------------------------------------
// Sampling from the resident state.  The reference materialises p = |psi|^2 and its
// cumulative sum (2 x 2^n floats) and searches it (backends/abstract_backend.py:1828-1861); here neither
// exists: one read of the state leaves the mass of every 2^SB-amplitude segment (float64), a one-CTA scan
// turns that into the segment CDF, and every shot re-reads only the one segment it lands in.
constexpr int SAMPLE_PER = 16;  // amplitudes per thread of the resolving CTA

__global__ void __launch_bounds__(256)
segment_mass_kernel(const float2* __restrict__ state, int seg_bits, uint64_t nseg, double* __restrict__ mass) {
  // one warp per segment (segments of >= 32 amplitudes), grid-stride over segments
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t len = 1ull << seg_bits;
  for (uint64_t sgm = warp; sgm < nseg; sgm += nwarp) {
    const float2* st = state + (sgm << seg_bits);
    double acc = 0.0;
    for (uint64_t i = lane; i < len; i += 32 * 4) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t x = i + 32ull * j;
        if (x < len) {
          const float2 a = st[x];
          part += a.x * a.x + a.y * a.y;
        }
      }
      acc += (double)part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) mass[sgm] = acc;
  }
}

// inclusive scan of n float64 values in place, one CTA (n is at most a few hundred thousand)
__global__ void __launch_bounds__(1024)
scan_inplace_kernel(double* __restrict__ v, uint64_t n) {
  __shared__ double tot[1024];
  const uint64_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  double s = 0.0;
  for (uint64_t i = lo; i < hi; ++i) s += v[i];
  tot[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (unsigned t = 0; t < blockDim.x; ++t) {
      const double x = tot[t];
      tot[t] = run;
      run += x;
    }
  }
  __syncthreads();
  double run = tot[threadIdx.x];
  for (uint64_t i = lo; i < hi; ++i) {
    run += v[i];
    v[i] = run;
  }
}

// mode 0: CDF inversion with ONE uniform per shot (probability_sample): first index whose inclusive
//         cumulative mass reaches total * (1 - u).
// mode 1: conditional walk with nbits uniforms per shot, qubit 0 first (the order and the decision rule of
//         the reference's perfect sampling = measure_jit on every qubit, basecircuit.py:449-459,:516-531):
//         bit_j = 1 iff u_j - p(bit_j = 0 | bits before) + eps > 0.
constexpr double kMeasureEps = 0.31415926e-12;  // basecircuit.py:524
__global__ void __launch_bounds__(256)
sample_resolve_kernel(const float2* __restrict__ state, int nbits, int seg_bits, const double* __restrict__ cdf,
                      const double* __restrict__ status, int mode, long long* __restrict__ out_index,
                      double* __restrict__ out_prob) {
  __shared__ double pre[1 << 12];   // per-thread-chunk inclusive prefix of |psi|^2 in the segment
  __shared__ double tcum[257];      // exclusive prefix of the chunk totals
  __shared__ unsigned long long s_seg;
  __shared__ double s_target;
  const uint64_t shot = blockIdx.x;
  const uint64_t nseg = 1ull << (nbits - seg_bits);
  const uint64_t len = 1ull << seg_bits;
  const int per = len < (uint64_t)SAMPLE_PER ? (int)len : SAMPLE_PER;
  const int nthr = (int)(len / per);
  const double total = cdf[nseg - 1];
  auto seg_cdf = [&](uint64_t i) -> double { return i == 0 ? 0.0 : cdf[i - 1]; };  // exclusive
  if (threadIdx.x == 0) {
    uint64_t seg = 0;
    double target = 0.0;
    if (mode == 0) {
      const double r = total * (1.0 - status[shot]);
      uint64_t lo = 0, hi = nseg - 1;  // first segment with inclusive cdf >= r
      while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (cdf[mid] >= r) hi = mid; else lo = mid + 1;
      }
      seg = lo;
      target = r - seg_cdf(seg);
    } else {
      uint64_t lo = 0, hi = nseg;
      for (int j = 0; j < nbits - seg_bits; ++j) {
        const uint64_t mid = (lo + hi) >> 1;
        const double m0 = seg_cdf(mid) - seg_cdf(lo), m1 = seg_cdf(hi) - seg_cdf(mid);
        const double p0 = m0 + m1 > 0.0 ? m0 / (m0 + m1) : 1.0;
        if (status[shot * nbits + j] - p0 + kMeasureEps > 0.0) lo = mid; else hi = mid;
      }
      seg = lo;
    }
    s_seg = seg;
    s_target = target;
  }
  __syncthreads();
  const float2* st = state + ((uint64_t)s_seg << seg_bits);
  if ((int)threadIdx.x < nthr) {
    double run = 0.0;
    for (int e = 0; e < per; ++e) {
      const float2 a = st[(uint64_t)threadIdx.x * per + e];
      run += (double)(a.x * a.x + a.y * a.y);
      pre[threadIdx.x * per + e] = run;
    }
    tcum[threadIdx.x + 1] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    tcum[0] = 0.0;
    for (int t = 1; t <= nthr; ++t) {
      run += tcum[t];
      tcum[t] = run;  // tcum[t] = mass of chunks [0, t)
    }
    uint64_t idx;
    if (mode == 0) {
      const double tgt = s_target;
      int lo = 0, hi = nthr - 1;  // first chunk whose inclusive mass reaches the target
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tcum[mid + 1] >= tgt) hi = mid; else lo = mid + 1;
      }
      const double rem = tgt - tcum[lo];
      int e = 0;
      while (e < per - 1 && pre[lo * per + e] < rem) ++e;
      idx = (uint64_t)lo * per + e;
    } else {
      // mass of [a, b) inside the segment from the two-level prefix
      auto cum = [&](uint64_t x) -> double {  // mass of [0, x)
        if (x == 0) return 0.0;
        const uint64_t c = (x - 1) / per;
        return tcum[c] + pre[x - 1];
      };
      uint64_t lo = 0, hi = len;
      for (int j = nbits - seg_bits; j < nbits; ++j) {
        const uint64_t mid = (lo + hi) >> 1;
        const double m0 = cum(mid) - cum(lo), m1 = cum(hi) - cum(mid);
        const double p0 = m0 + m1 > 0.0 ? m0 / (m0 + m1) : 1.0;
        if (status[shot * nbits + j] - p0 + kMeasureEps > 0.0) lo = mid; else hi = mid;
      }
      idx = lo;
    }
    const float2 a = st[idx];
    out_index[shot] = (long long)(((uint64_t)s_seg << seg_bits) | idx);
    if (out_prob) out_prob[shot] = (double)(a.x * a.x + a.y * a.y) / total;
  }
}

int launch_sample_prepare(const void* state, int nbits, int seg_bits, double* cdf, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_sample_prepare: nbits=%d", nbits);
  TCB_REQUIRE(seg_bits >= 0 && seg_bits <= 12 && seg_bits <= nbits, "tcb_sv_sample_prepare: seg_bits=%d", seg_bits);
  TCB_REQUIRE(nbits - seg_bits <= 28, "tcb_sv_sample_prepare: too many segments (nbits - seg_bits = %d)", nbits - seg_bits);
  const uint64_t nseg = 1ull << (nbits - seg_bits);
  segment_mass_kernel<<<grid_for(nseg * 32, 256), 256, 0, stream>>>(reinterpret_cast<const float2*>(state), seg_bits,
                                                                   nseg, cdf);
  TCB_CHECK_CUDA(cudaGetLastError());
  scan_inplace_kernel<<<1, 1024, 0, stream>>>(cdf, nseg);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_sample(const void* state, int nbits, int seg_bits, const double* cdf, const double* status,
                  int64_t shots, int mode, long long* out_index, double* out_prob, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_sample: nbits=%d", nbits);
  TCB_REQUIRE(seg_bits >= 0 && seg_bits <= 12 && seg_bits <= nbits, "tcb_sv_sample: seg_bits=%d", seg_bits);
  TCB_REQUIRE(mode == 0 || mode == 1, "tcb_sv_sample: mode=%d (0 = cdf, 1 = conditional walk)", mode);
  TCB_REQUIRE(shots >= 0 && shots <= (1ll << 31) - 1, "tcb_sv_sample: shots=%lld", (long long)shots);
  if (shots == 0) return 0;
  sample_resolve_kernel<<<(unsigned)shots, 256, 0, stream>>>(reinterpret_cast<const float2*>(state), nbits, seg_bits,
                                                            cdf, status, mode, out_index, out_prob);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// adjoint-mode gate gradient: G[r][c] = sum_rest lam[rest,r] * conj(psi[rest,c])
template <int K>
__global__ void __launch_bounds__(256)
gate_grad_kernel(const float2* __restrict__ lam, const float2* __restrict__ psi, int nbits,
                 uint64_t groups_per_state, BitList bl, double* grad, long long grad_bstride) {
  constexpr int D = 1 << K;
  const unsigned b = blockIdx.y;
  const float2* pl = lam + ((size_t)b << nbits);
  const float2* pp = psi + ((size_t)b << nbits);
  float2 acc[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) acc[e] = make_float2(0.f, 0.f);
  double2 dacc[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) dacc[e] = make_double2(0.0, 0.0);
  int cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups_per_state; g += stride) {
    uint64_t base = g;
#pragma unroll
    for (int i = 0; i < K; ++i) base = insert_zero(base, bl.sorted[i]);
    float2 l[D], p[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      uint64_t o = 0;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((c >> (K - 1 - i)) & 1) o |= 1ull << bl.pos[i];
      l[c] = pl[base | o];
      p[c] = pp[base | o];
    }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        // l[r] * conj(p[c])
        acc[r * D + c].x += l[r].x * p[c].x + l[r].y * p[c].y;
        acc[r * D + c].y += l[r].y * p[c].x - l[r].x * p[c].y;
      }
    if (++cnt == 32) {
#pragma unroll
      for (int e = 0; e < D * D; ++e) {
        dacc[e].x += acc[e].x;
        dacc[e].y += acc[e].y;
        acc[e] = make_float2(0.f, 0.f);
      }
      cnt = 0;
    }
  }
#pragma unroll
  for (int e = 0; e < D * D; ++e) {
    dacc[e].x += acc[e].x;
    dacc[e].y += acc[e].y;
  }
  double* gout = grad + (size_t)b * grad_bstride * 2;
#pragma unroll
  for (int e = 0; e < D * D; ++e) {
    block_reduce_add2(dacc[e].x, dacc[e].y, gout + 2 * e);
    __syncthreads();
  }
}

int launch_gate_grad(const void* lam, const void* psi, int nbits, int64_t batch, const int* bitpos,
                     int k, void* grad, int64_t grad_bstride, cudaStream_t stream) {
  TCB_REQUIRE(k >= 1 && k <= 2, "tcb_sv_gate_grad: k=%d unsupported (1..2)", k);
  TCB_REQUIRE(nbits >= k && nbits <= 40, "tcb_sv_gate_grad: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_gate_grad: batch=%lld", (long long)batch);
  BitList bl;
  for (int i = 0; i < k; ++i) {
    TCB_REQUIRE(bitpos[i] >= 0 && bitpos[i] < nbits, "tcb_sv_gate_grad: bit %d out of range", bitpos[i]);
    bl.pos[i] = bitpos[i];
    bl.sorted[i] = bitpos[i];
  }
  if (k == 2 && bl.sorted[0] > bl.sorted[1]) {
    int t = bl.sorted[0];
    bl.sorted[0] = bl.sorted[1];
    bl.sorted[1] = t;
  }
  const uint64_t gps = 1ull << (nbits - k);
  dim3 grid(grid_for(gps, 256, 4), (unsigned)batch);
  const float2* l = reinterpret_cast<const float2*>(lam);
  const float2* p = reinterpret_cast<const float2*>(psi);
  double* g = reinterpret_cast<double*>(grad);
  if (k == 1)
    gate_grad_kernel<1><<<grid, 256, 0, stream>>>(l, p, nbits, gps, bl, g, grad_bstride);
  else
    gate_grad_kernel<2><<<grid, 256, 0, stream>>>(l, p, nbits, gps, bl, g, grad_bstride);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// One gate of the adjoint walk in ONE pass over the two states (instead of apply / reduce / apply):
//   psi <- U^dagger psi ;  G[r][c] += sum_rest lam[rest,r] conj(psi_in[rest,c]) ;  lam <- U^dagger lam
// 2 reads + 2 writes of a state per gate instead of 4 reads + 2 writes and three launches.
template <int K>
__global__ void __launch_bounds__(256, K == 2 ? 2 : 4)
adjoint_step_kernel(float2* __restrict__ lam, float2* __restrict__ psi, int nbits, uint64_t groups_per_state,
                    BitList bl, const float2* __restrict__ udag, long long udag_bstride, double* grad,
                    long long grad_bstride) {
  constexpr int D = 1 << K;
  const unsigned b = blockIdx.y;
  float2* pl = lam + ((size_t)b << nbits);
  float2* pp = psi + ((size_t)b << nbits);
  __shared__ float2 u[D * D];  // broadcast reads: keeps 2 D^2 registers free for loads in flight
  if (threadIdx.x < D * D) u[threadIdx.x] = udag[(size_t)b * udag_bstride + threadIdx.x];
  __syncthreads();
  // two-level float accumulation per thread (runs of 32), double across threads and blocks
  float2 acc[D * D], acc2[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) acc[e] = acc2[e] = make_float2(0.f, 0.f);
  int cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups_per_state; g += stride) {
    uint64_t base = g;
#pragma unroll
    for (int i = 0; i < K; ++i) base = insert_zero(base, bl.sorted[i]);
    float2 l[D], p[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      uint64_t o = 0;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((c >> (K - 1 - i)) & 1) o |= 1ull << bl.pos[i];
      l[c] = pl[base | o];
      p[c] = pp[base | o];
    }
    float2 pin[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      float2 sp = make_float2(0.f, 0.f), sl = sp;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const float2 m = u[r * D + c];
        sp.x += m.x * p[c].x - m.y * p[c].y;
        sp.y += m.x * p[c].y + m.y * p[c].x;
        sl.x += m.x * l[c].x - m.y * l[c].y;
        sl.y += m.x * l[c].y + m.y * l[c].x;
      }
      pin[r] = sp;
      uint64_t o = 0;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if ((r >> (K - 1 - i)) & 1) o |= 1ull << bl.pos[i];
      pp[base | o] = sp;
      pl[base | o] = sl;
    }
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        acc[r * D + c].x += l[r].x * pin[c].x + l[r].y * pin[c].y;
        acc[r * D + c].y += l[r].y * pin[c].x - l[r].x * pin[c].y;
      }
    if ((++cnt & 31) == 0) {
#pragma unroll
      for (int e = 0; e < D * D; ++e) {
        acc2[e].x += acc[e].x;
        acc2[e].y += acc[e].y;
        acc[e] = make_float2(0.f, 0.f);
      }
    }
  }
  double* gout = grad + (size_t)b * grad_bstride * 2;
#pragma unroll
  for (int e = 0; e < D * D; ++e) {
    const double re = (double)acc2[e].x + (double)acc[e].x;
    const double im = (double)acc2[e].y + (double)acc[e].y;
    block_reduce_add2(re, im, gout + 2 * e);
    __syncthreads();
  }
}

int launch_adjoint_step(void* lam, void* psi, int nbits, int64_t batch, const int* bitpos, int k,
                        const void* udag, int64_t udag_bstride, void* grad, int64_t grad_bstride,
                        cudaStream_t stream) {
  TCB_REQUIRE(k >= 1 && k <= 2, "tcb_sv_adjoint_step: k=%d unsupported (1..2)", k);
  TCB_REQUIRE(nbits >= k && nbits <= 40, "tcb_sv_adjoint_step: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_adjoint_step: batch=%lld", (long long)batch);
  TCB_REQUIRE(lam != psi, "tcb_sv_adjoint_step: lam and psi must be different buffers");
  BitList bl;
  for (int i = 0; i < k; ++i) {
    TCB_REQUIRE(bitpos[i] >= 0 && bitpos[i] < nbits, "tcb_sv_adjoint_step: bit %d out of range", bitpos[i]);
    bl.pos[i] = bitpos[i];
    bl.sorted[i] = bitpos[i];
  }
  if (k == 2) {
    TCB_REQUIRE(bl.pos[0] != bl.pos[1], "tcb_sv_adjoint_step: repeated bit %d", bl.pos[0]);
    if (bl.sorted[0] > bl.sorted[1]) {
      int t = bl.sorted[0];
      bl.sorted[0] = bl.sorted[1];
      bl.sorted[1] = t;
    }
  }
  const uint64_t gps = 1ull << (nbits - k);
  dim3 grid(grid_for(gps, 256, 8), (unsigned)batch);
  float2* l = reinterpret_cast<float2*>(lam);
  float2* p = reinterpret_cast<float2*>(psi);
  const float2* u = reinterpret_cast<const float2*>(udag);
  double* g = reinterpret_cast<double*>(grad);
  if (k == 1)
    adjoint_step_kernel<1><<<grid, 256, 0, stream>>>(l, p, nbits, gps, bl, u, udag_bstride, g, grad_bstride);
  else
    adjoint_step_kernel<2><<<grid, 256, 0, stream>>>(l, p, nbits, gps, bl, u, udag_bstride, g, grad_bstride);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// Gradients of a RUN of consecutive diagonal gates in one read of the two states.  Diagonal gates commute,
// so with psi_L / lam_L the states AFTER the run,  dL/dd_j[c] = d_j[c] * sum_{i in c} lam_L[i] conj(psi_L[i])
// for every gate j of the run (unit-modulus d): the only state-sized work is the marginal of the
// elementwise product over the gate's one or two bits.  Up to CM_G gates per launch (moment sums,
// predicated adds; two-level float accumulation, double across threads).
// A gate's four bins are combinations of the +-1 MOMENTS  M_S = sum_i (-1)^popc(i & S) q_i,  q = lam conj(psi),
// over S in {0, a, b, ab}: the launcher collects the DISTINCT masks of the whole run (a TFIM layer of 23 rzz and
// 24 rz gates has 47 of them, not 141: every single-bit moment is shared by three gates), CM_M masks per read.
constexpr int CM_M = 24;

struct CmMasks {
  unsigned long long m[CM_M];
};

__device__ __forceinline__ float2 cm_flip(float2 v, unsigned sign_bit) {  // sign_bit: 0 or 0x80000000
  return make_float2(__uint_as_float(__float_as_uint(v.x) ^ sign_bit), __uint_as_float(__float_as_uint(v.y) ^ sign_bit));
}

// mom[0] += M_0 (first launch of a series only), mom[1 + j] += M_{g.m[j]}.  A thread takes the 8 amplitudes of the
// three lowest address bits of both states per iteration (64 contiguous bytes each).  LOW = false: no mask touches
// those bits, so the sign of a mask is one value for all 8 products — one parity, one sign flip and one complex add
// per mask per EIGHT amplitudes (the round-1 form paid that per pair and was ALU-bound at 1.9 TB/s).  LOW = true:
// the 3-bit Walsh-Hadamard transform of the 8 products first, every mask then picks the entry of its low part.
template <bool LOW>
__global__ void __launch_bounds__(256, 2)
cross_moments_kernel(const float4* __restrict__ lam, const float4* __restrict__ psi, uint64_t noct, CmMasks g,
                     int nmasks, int write_m0, double* mom, long long mom_bstride) {
  lam += (size_t)blockIdx.y * noct * 4;
  psi += (size_t)blockIdx.y * noct * 4;
  mom += (size_t)blockIdx.y * mom_bstride;
  // first-level sums in registers, second level per thread in shared memory ([slot][thread], conflict-free)
  extern __shared__ float cm_acc2[];
  float2 acc[CM_M], m0 = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < CM_M; ++j) acc[j] = make_float2(0.f, 0.f);
  for (int e = 0; e < CM_M * 2 + 2; ++e) cm_acc2[e * 256 + threadIdx.x] = 0.f;
  int cnt = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < noct; p += stride) {
    float4 l[4], s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // eight independent 16-byte loads in flight
      l[k] = ldg_stream(lam + 4 * p + k);
      s[k] = ldg_stream(psi + 4 * p + k);
    }
    float2 w[8];  // q_i = lam_i conj(psi_i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      w[2 * k] = make_float2(l[k].x * s[k].x + l[k].y * s[k].y, l[k].y * s[k].x - l[k].x * s[k].y);
      w[2 * k + 1] = make_float2(l[k].z * s[k].z + l[k].w * s[k].w, l[k].w * s[k].z - l[k].z * s[k].w);
    }
    if (LOW) {  // in-place Walsh-Hadamard over the three low bits: w[S] = sum_i (-1)^popc(i & S) q_i
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (!((i >> b) & 1)) {
            const float2 x = w[i], y = w[i | (1 << b)];
            w[i] = make_float2(x.x + y.x, x.y + y.y);
            w[i | (1 << b)] = make_float2(x.x - y.x, x.y - y.y);
          }
    } else {
#pragma unroll
      for (int i = 1; i < 8; ++i) {
        w[0].x += w[i].x;
        w[0].y += w[i].y;
      }
    }
    m0.x += w[0].x;
    m0.y += w[0].y;
#pragma unroll
    for (int j = 0; j < CM_M; ++j) {
      const unsigned long long m = g.m[j];
      const unsigned sg = (unsigned)(__popcll(p & (m >> 3)) & 1) << 31;
      float2 base = w[0];
      if (LOW) {
        switch ((int)(m & 7ull)) {  // (uniform: the masks are kernel parameters)
          case 1: base = w[1]; break;
          case 2: base = w[2]; break;
          case 3: base = w[3]; break;
          case 4: base = w[4]; break;
          case 5: base = w[5]; break;
          case 6: base = w[6]; break;
          case 7: base = w[7]; break;
          default: break;
        }
      }
      const float2 u = cm_flip(base, sg);
      acc[j].x += u.x;
      acc[j].y += u.y;
    }
    if ((++cnt & 15) == 0) {
      cm_acc2[(CM_M * 2) * 256 + threadIdx.x] += m0.x;
      cm_acc2[(CM_M * 2 + 1) * 256 + threadIdx.x] += m0.y;
      m0 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < CM_M; ++j) {
        cm_acc2[(j * 2) * 256 + threadIdx.x] += acc[j].x;
        cm_acc2[(j * 2 + 1) * 256 + threadIdx.x] += acc[j].y;
        acc[j] = make_float2(0.f, 0.f);
      }
    }
  }
  if (write_m0) {
    block_reduce_add2((double)cm_acc2[(CM_M * 2) * 256 + threadIdx.x] + (double)m0.x,
                      (double)cm_acc2[(CM_M * 2 + 1) * 256 + threadIdx.x] + (double)m0.y, mom);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < CM_M; ++j) {
    if (j < nmasks) {  // (uniform)
      block_reduce_add2((double)cm_acc2[(j * 2) * 256 + threadIdx.x] + (double)acc[j].x,
                        (double)cm_acc2[(j * 2 + 1) * 256 + threadIdx.x] + (double)acc[j].y, mom + 2 * (1 + j));
      __syncthreads();
    }
  }
}

// states of fewer than 8 amplitudes: one thread, plain loops (every mask may touch every bit)
__global__ void cross_moments_tiny_kernel(const float2* __restrict__ lam, const float2* __restrict__ psi, int nbits,
                                          const CmMasks g, int nmasks, int write_m0, double* mom, long long mom_bstride) {
  if (threadIdx.x != 0) return;
  lam += (size_t)blockIdx.y << nbits;
  psi += (size_t)blockIdx.y << nbits;
  mom += (size_t)blockIdx.y * mom_bstride;
  for (int j = -1; j < nmasks; ++j) {
    if (j < 0 && !write_m0) continue;
    double re = 0.0, im = 0.0;
    for (unsigned i = 0; i < (1u << nbits); ++i) {
      const float2 l = lam[i], s = psi[i];
      const double sg = (j >= 0 && (__popcll((unsigned long long)i & g.m[j]) & 1)) ? -1.0 : 1.0;
      re += sg * ((double)l.x * s.x + (double)l.y * s.y);
      im += sg * ((double)l.y * s.x - (double)l.x * s.y);
    }
    mom[2 * (1 + j)] += re;
    mom[2 * (1 + j) + 1] += im;
  }
}

// moments -> bins: out[j][(ca << 1) | cb] = 1/4 sum_{sa, sb} (-1)^(ca sa + cb sb) M[(sa << 1) | sb] with
// M = {M_0, M_b, M_a, M_ab}; a one-qubit gate (no bit a) has M_a = M_0, M_ab = M_b.  `mom` may BE `out` (the
// moments of a run fit its output: 1 + #masks <= 4 #gates): one block per batch row reads every gate's moments
// before any bin is written.
constexpr int CM_BINS_G = 256;
struct CmGateIdx {
  short ia[CM_BINS_G], ib[CM_BINS_G], iab[CM_BINS_G];  // slots in the moments buffer; ia < 0: one-qubit gate
};
__global__ void __launch_bounds__(CM_BINS_G)
cm_moments_to_bins_kernel(const double* mom, long long mom_bstride, CmGateIdx gi, int ngates, double* out,
                          long long out_bstride) {
  const int j = threadIdx.x;
  const double* mb = mom + (size_t)blockIdx.y * mom_bstride;
  double* o = out + (size_t)blockIdx.y * out_bstride + 8 * (size_t)j;
  double re[4], im[4];
  if (j < ngates) {
    const int ia = gi.ia[j], ib = gi.ib[j], iab = gi.iab[j];
    const int slot[4] = {0, ib, ia < 0 ? 0 : ia, ia < 0 ? ib : iab};
    for (int s = 0; s < 4; ++s) {
      re[s] = mb[2 * slot[s]];
      im[s] = mb[2 * slot[s] + 1];
    }
  }
  __syncthreads();
  if (j >= ngates) return;
  for (int c = 0; c < 4; ++c) {
    double r = 0.0, i = 0.0;
    for (int s = 0; s < 4; ++s) {
      const double sg = (__popc(c & s) & 1) ? -0.25 : 0.25;
      r += sg * re[s];
      i += sg * im[s];
    }
    o[2 * c] = r;
    o[2 * c + 1] = i;
  }
}

int launch_cross_marginals(const void* lam, const void* psi, int nbits, int64_t batch, int ngates,
                           const int* gate_bits, double* out, int64_t out_bstride, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_cross_marginals: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_cross_marginals: batch=%lld", (long long)batch);
  TCB_REQUIRE(ngates >= 0, "tcb_sv_cross_marginals: ngates=%d", ngates);
  TCB_REQUIRE(batch == 1 || out_bstride >= (int64_t)4 * ngates,
              "tcb_sv_cross_marginals: out_batch_stride must be >= 4 * ngates");
  if (ngates == 0) return 0;
  // distinct masks of the run, and for every gate the moment slots of {a, b, ab} (slot 0 = M_0); the masks that
  // touch the three lowest bits come first (they take the Walsh-Hadamard variant of the kernel)
  std::vector<unsigned long long> masks;
  auto add_mask = [&](unsigned long long m) {
    for (size_t k = 0; k < masks.size(); ++k)
      if (masks[k] == m) return;
    masks.push_back(m);
  };
  for (int pass = 0; pass < 2; ++pass)  // pass 0: low-touching masks, pass 1: the rest
    for (int j = 0; j < ngates; ++j) {
      const int a = gate_bits[2 * j], b = gate_bits[2 * j + 1];
      TCB_REQUIRE(a >= 0 && a < nbits && b >= -1 && b < nbits && a != b,
                  "tcb_sv_cross_marginals: gate %d has bits (%d, %d)", j, a, b);
      unsigned long long cand[3] = {1ull << a, b < 0 ? 0ull : 1ull << b, b < 0 ? 0ull : (1ull << a) | (1ull << b)};
      for (int c = 0; c < 3; ++c)
        if (cand[c] != 0ull && (((cand[c] & 7ull) != 0ull) == (pass == 0))) add_mask(cand[c]);
    }
  int nlow = 0;
  while (nlow < (int)masks.size() && (masks[nlow] & 7ull) != 0ull) ++nlow;
  auto slot_of = [&](unsigned long long m) {
    for (size_t k = 0; k < masks.size(); ++k)
      if (masks[k] == m) return (int)k + 1;
    return 0;
  };
  std::vector<int> ia(ngates), ib(ngates), iab(ngates);
  for (int j = 0; j < ngates; ++j) {
    const int a = gate_bits[2 * j], b = gate_bits[2 * j + 1];
    if (b < 0) {  // one-qubit gate: its bit plays the LSB of the bin index
      ia[j] = -1;
      ib[j] = slot_of(1ull << a);
      iab[j] = ib[j];
    } else {
      ia[j] = slot_of(1ull << a);
      ib[j] = slot_of(1ull << b);
      iab[j] = slot_of((1ull << a) | (1ull << b));
    }
  }
  const int nmom = (int)masks.size();
  TCB_REQUIRE(nmom < 32000, "tcb_sv_cross_marginals: %d distinct masks in one run", nmom);
  // the moments live in `out` itself (1 + nmom <= 4 ngates slots) when one block can turn them into bins;
  // longer runs take a stream-ordered scratch buffer
  const bool in_place = ngates <= CM_BINS_G;
  long long mom_bstride = 2 * out_bstride;  // doubles per batch element
  double* mom = out;
  if (in_place) {
    if (batch == 1)
      TCB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 8 * (size_t)ngates, stream));
    else
      TCB_CHECK_CUDA(cudaMemset2DAsync(out, sizeof(double) * 2 * (size_t)out_bstride, 0, sizeof(double) * 8 * (size_t)ngates,
                                       (size_t)batch, stream));
  } else {
    mom_bstride = 2 * (long long)(1 + nmom);
    const size_t mom_bytes = sizeof(double) * (size_t)mom_bstride * (size_t)batch;
    TCB_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&mom), mom_bytes, stream));
    TCB_CHECK_CUDA(cudaMemsetAsync(mom, 0, mom_bytes, stream));
  }
  constexpr size_t cm_smem = sizeof(float) * 256 * (CM_M * 2 + 2);
  static bool attr_set = false;
  if (!attr_set) {
    TCB_CHECK_CUDA(cudaFuncSetAttribute(cross_moments_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cm_smem));
    TCB_CHECK_CUDA(cudaFuncSetAttribute(cross_moments_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cm_smem));
    attr_set = true;
  }
  bool m0_done = false;
  auto run_range = [&](int lo, int hi, bool low) -> int {
    for (int first = lo; first < hi; first += CM_M) {
      const int cnt = hi - first < CM_M ? hi - first : CM_M;
      CmMasks g;
      for (int j = 0; j < CM_M; ++j) g.m[j] = j < cnt ? masks[first + j] : 0ull;
      double* dst = mom + 2 * first;
      const int wm0 = m0_done ? 0 : 1;
      if (nbits < 3) {
        cross_moments_tiny_kernel<<<dim3(1, (unsigned)batch), 32, 0, stream>>>(
            reinterpret_cast<const float2*>(lam), reinterpret_cast<const float2*>(psi), nbits, g, cnt, wm0,
            first == 0 ? mom : dst, mom_bstride);
      } else {
        const uint64_t noct = 1ull << (nbits - 3);
        dim3 grid(grid_for(noct, 256, 2), (unsigned)batch);
        if (low)
          cross_moments_kernel<true><<<grid, 256, cm_smem, stream>>>(reinterpret_cast<const float4*>(lam),
                                                                     reinterpret_cast<const float4*>(psi), noct, g, cnt,
                                                                     wm0, dst, mom_bstride);
        else
          cross_moments_kernel<false><<<grid, 256, cm_smem, stream>>>(reinterpret_cast<const float4*>(lam),
                                                                      reinterpret_cast<const float4*>(psi), noct, g, cnt,
                                                                      wm0, dst, mom_bstride);
      }
      TCB_CHECK_CUDA(cudaGetLastError());
      m0_done = true;
    }
    return 0;
  };
  if (int rc = run_range(0, nbits < 3 ? nmom : nlow, true)) return rc;
  if (nbits >= 3)
    if (int rc = run_range(nlow, nmom, false)) return rc;
  for (int first = 0; first < ngates; first += CM_BINS_G) {
    const int cnt = ngates - first < CM_BINS_G ? ngates - first : CM_BINS_G;
    CmGateIdx gi;
    for (int j = 0; j < CM_BINS_G; ++j) {
      const int src = first + (j < cnt ? j : 0);
      gi.ia[j] = (short)ia[src];
      gi.ib[j] = (short)ib[src];
      gi.iab[j] = (short)iab[src];
    }
    dim3 grid(1, (unsigned)batch);
    cm_moments_to_bins_kernel<<<grid, CM_BINS_G, 0, stream>>>(mom, mom_bstride, gi, cnt, out + 8 * first, 2 * out_bstride);
    TCB_CHECK_CUDA(cudaGetLastError());
  }
  if (!in_place) TCB_CHECK_CUDA(cudaFreeAsync(mom, stream));
  return 0;
}

// ---------------------------------------------------------------------------------
// Cross reduced density matrices of single qubits between two states, many qubits per read:
//   C_q[r][c] = sum_rest lam[rest, q = r] conj(psi[rest, q = c])
// for every qubit of a tile of up to 10 bits (the three lowest address bits + up to seven chosen ones; a tile
// of 1024 amplitudes of each state sits in shared memory).  With lam_0 / psi_0 the states BEFORE a layer of
// one-qubit gates on distinct qubits, dL/dU_q = U_q C_q for every gate of the layer at once (the other
// gates of the layer are unitaries on other qubits and drop out of the partial trace).
constexpr int CR_MAXB = 10;

struct CrBits {
  int nlow, nsel;
  int sel[7];
};

__global__ void __launch_bounds__(256)
cross_rdm_kernel(const float2* __restrict__ lam, const float2* __restrict__ psi, int nbits, CrBits cb, int t_begin,
                 double* out, long long out_bstride) {
  lam += (size_t)blockIdx.y << nbits;
  psi += (size_t)blockIdx.y << nbits;
  out += (size_t)blockIdx.y * out_bstride;
  __shared__ __align__(16) float2 sl[1 << CR_MAXB], sp[1 << CR_MAXB];
  __shared__ double sacc[CR_MAXB * 3 * 2 + 2];  // per bit [0][0], [0][1], [1][0]; then the total sum lam conj(psi)
  const int tb = cb.nlow + cb.nsel;          // tile bits
  const int tsize = 1 << tb;
  const int tid = threadIdx.x;
  for (int e = tid; e < CR_MAXB * 6 + 2; e += blockDim.x) sacc[e] = 0.0;
  // C[1][1] = (sum over ALL amplitudes of lam conj(psi)) - C[0][0] for every bit: three products per pair, not four
  float2 acc[CR_MAXB][3], tot = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < CR_MAXB; ++t)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[t][c] = make_float2(0.f, 0.f);
  auto warp_add = [&](float re, float im, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&sacc[2 * slot], (double)re);
      atomicAdd(&sacc[2 * slot + 1], (double)im);
    }
  };
  auto flush = [&]() {
#pragma unroll
    for (int t = 0; t < CR_MAXB; ++t)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (t >= t_begin && t < tb) warp_add(acc[t][c].x, acc[t][c].y, t * 3 + c);  // (uniform)
        acc[t][c] = make_float2(0.f, 0.f);
      }
    warp_add(tot.x, tot.y, CR_MAXB * 3);
    tot = make_float2(0.f, 0.f);
  };
  const uint64_t ntiles = 1ull << (nbits - tb);
  int it = 0;
  __syncthreads();
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // tile base: spread the tile counter over the bits that are NOT in the tile
    uint64_t base = tile << cb.nlow;
    for (int j = 0; j < cb.nsel; ++j) base = insert_zero(base, cb.sel[j]);  // sel ascending
    if (tb == CR_MAXB) {
      // full tile: 2 x 2 independent 16-byte loads per thread in flight before anything is stored (a simple
      // element loop keeps one 8-byte load per state in flight and is latency-bound at a tenth of HBM speed)
      float4 rl[2], rp[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int e = (tid + k * 256) << 1;  // even amplitude index inside the tile; low three bits = address bits
        uint64_t a = base | (uint64_t)(e & 7);
#pragma unroll
        for (int j = 0; j < 7; ++j) a |= (uint64_t)((e >> (3 + j)) & 1) << cb.sel[j];
        rl[k] = ldg_stream(reinterpret_cast<const float4*>(lam + a));
        rp[k] = ldg_stream(reinterpret_cast<const float4*>(psi + a));
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int e = (tid + k * 256) << 1;
        *reinterpret_cast<float4*>(&sl[e]) = rl[k];
        *reinterpret_cast<float4*>(&sp[e]) = rp[k];
        tot.x += rl[k].x * rp[k].x + rl[k].y * rp[k].y + rl[k].z * rp[k].z + rl[k].w * rp[k].w;
        tot.y += rl[k].y * rp[k].x - rl[k].x * rp[k].y + rl[k].w * rp[k].z - rl[k].z * rp[k].w;
      }
    } else {
      for (int e = tid; e < tsize; e += blockDim.x) {
        uint64_t a = base | (uint64_t)(e & ((1 << cb.nlow) - 1));
        for (int j = 0; j < cb.nsel; ++j) a |= (uint64_t)((e >> (cb.nlow + j)) & 1) << cb.sel[j];
        const float2 l = lam[a], q = psi[a];
        sl[e] = l;
        sp[e] = q;
        tot.x += l.x * q.x + l.y * q.y;
        tot.y += l.y * q.x - l.x * q.y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CR_MAXB; ++t) {
      if (t >= t_begin && t < tb) {
        for (int j = tid; j < (tsize >> 1); j += blockDim.x) {
          const int i0 = (int)insert_zero((uint64_t)j, t), i1 = i0 | (1 << t);
          const float2 l0 = sl[i0], l1 = sl[i1], p0 = sp[i0], p1 = sp[i1];
          acc[t][0].x += l0.x * p0.x + l0.y * p0.y;  acc[t][0].y += l0.y * p0.x - l0.x * p0.y;  // [0][0]
          acc[t][1].x += l0.x * p1.x + l0.y * p1.y;  acc[t][1].y += l0.y * p1.x - l0.x * p1.y;  // [0][1]
          acc[t][2].x += l1.x * p0.x + l1.y * p0.y;  acc[t][2].y += l1.y * p0.x - l1.x * p0.y;  // [1][0]
        }
      }
    }
    __syncthreads();
    if (++it == 32) {
      flush();
      it = 0;
    }
  }
  flush();
  __syncthreads();
  // out[t][r][c]: [0][0], [0][1], [1][0] as accumulated, [1][1] = total - [0][0]
  for (int e = tid; e < tb * 8; e += blockDim.x) {
    const int t = e >> 3, c = (e >> 1) & 3, ri = e & 1;
    if (t < t_begin) continue;
    const double v = c < 3 ? sacc[(t * 3 + c) * 2 + ri] : sacc[CR_MAXB * 6 + ri] - sacc[(t * 3) * 2 + ri];
    atomicAdd(out + e, v);
  }
}

// Register / shuffle form (states of >= 2^(3 + nsel + 1) amplitudes): a thread holds the 8 amplitudes of the three
// lowest address bits of BOTH states in registers (64 contiguous bytes each: 8 independent 16-byte loads in flight),
// thread-index bit j < nsel stands for the selected bit sel[j] and the higher thread-index bits pick one of the
// 2^(8 - nsel) units a CTA works on at a time.  Per bit the partial traces need
//   low bits      : both partners are registers of the same thread;
//   lane bits     : the partner's psi comes through 16 shuffles, the thread with bit = 0 accumulates
//                   lam[x] conj(psi[x | t]) (-> C[0][1]), its partner lam[x | t] conj(psi[x]) (-> C[1][0]);
//   warp bits     : the same through one shared-memory exchange of psi per iteration;
//   C[0][0]       : needs no partner — the thread's running sum of lam conj(psi), counted at the end by the
//                   threads whose bit is 0 (C[1][1] = total - C[0][0]).
// No tile in shared memory and 2 instead of ~20 shared-memory bytes per byte of HBM traffic: the tile form above
// ran at 1.06 TB/s, bound by its shared-memory reads.
__global__ void __launch_bounds__(256, 2)
cross_rdm_reg_kernel(const float2* __restrict__ lam, const float2* __restrict__ psi, int nbits, CrBits cb, int t_begin,
                     double* out, long long out_bstride) {
  lam += (size_t)blockIdx.y << nbits;
  psi += (size_t)blockIdx.y << nbits;
  out += (size_t)blockIdx.y * out_bstride;
  __shared__ __align__(16) float4 xch[256 * 4];        // psi of every thread (warp-bit partners)
  __shared__ double sacc[CR_MAXB * 3 * 2 + 2];
  const int tid = threadIdx.x;
  const int nsel = cb.nsel;
  for (int e = tid; e < CR_MAXB * 6 + 2; e += blockDim.x) sacc[e] = 0.0;
  uint64_t toff = 0;
  for (int j = 0; j < nsel; ++j) toff |= (uint64_t)((tid >> j) & 1) << cb.sel[j];
  const unsigned unit_in_cta = (unsigned)tid >> nsel, units_per_cta = 256u >> nsel;
  const uint64_t nunits = 1ull << (nbits - 3 - nsel);
  const uint64_t niter = (nunits + (uint64_t)units_per_cta * gridDim.x - 1) / ((uint64_t)units_per_cta * gridDim.x);
  const bool do_low = t_begin == 0;

  float2 s = make_float2(0.f, 0.f);       // sum of lam conj(psi)
  float2 lo[3][3];                         // low bit t: [0][0], [0][1], [1][0]
  float2 hi[7];                            // selected bit j: lam_own conj(psi_partner)
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int c = 0; c < 3; ++c) lo[t][c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 7; ++j) hi[j] = make_float2(0.f, 0.f);

  auto cmac = [](float2& acc, float2 l, float2 q) {  // acc += l conj(q)
    acc.x = fmaf(l.x, q.x, fmaf(l.y, q.y, acc.x));
    acc.y = fmaf(l.y, q.x, fmaf(-l.x, q.y, acc.y));
  };
  auto warp_add = [&](float re, float im, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&sacc[2 * slot], (double)re);
      atomicAdd(&sacc[2 * slot + 1], (double)im);
    }
  };
  auto flush = [&]() {
    if (do_low) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) warp_add(lo[t][c].x, lo[t][c].y, t * 3 + c);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < nsel) {  // (uniform)
        const bool one = (tid >> j) & 1;
        warp_add(one ? 0.f : s.x, one ? 0.f : s.y, (3 + j) * 3 + 0);
        warp_add(one ? 0.f : hi[j].x, one ? 0.f : hi[j].y, (3 + j) * 3 + 1);
        warp_add(one ? hi[j].x : 0.f, one ? hi[j].y : 0.f, (3 + j) * 3 + 2);
      }
      hi[j] = make_float2(0.f, 0.f);
    }
    warp_add(s.x, s.y, CR_MAXB * 3);
    s = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
      for (int c = 0; c < 3; ++c) lo[t][c] = make_float2(0.f, 0.f);
  };

  __syncthreads();
  int since_flush = 0;
  // software pipeline: the loads of iteration it + 1 are in flight while iteration it is computed (a CTA's
  // threads move in lockstep through the shared-memory exchange, so nothing else hides the DRAM latency)
  float4 rl[4], rq[4];
  auto issue = [&](uint64_t it) -> bool {
    const uint64_t unit = (it * gridDim.x + blockIdx.x) * units_per_cta + unit_in_cta;
    const bool live = it < niter && unit < nunits;  // (a CTA's last iteration may have idle units; they still join the barriers)
    if (live) {
      uint64_t base = unit << 3;
      for (int j = 0; j < nsel; ++j) base = insert_zero(base, cb.sel[j]);  // sel ascending
      const float4* pl = reinterpret_cast<const float4*>(lam + (base | toff));
      const float4* pq = reinterpret_cast<const float4*>(psi + (base | toff));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rl[k] = ldg_stream(pl + k);
        rq[k] = ldg_stream(pq + k);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) rl[k] = rq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return live;
  };
  issue(0);
  for (uint64_t it = 0; it < niter; ++it) {
    float2 l[8], q[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      l[2 * k] = make_float2(rl[k].x, rl[k].y);
      l[2 * k + 1] = make_float2(rl[k].z, rl[k].w);
      q[2 * k] = make_float2(rq[k].x, rq[k].y);
      q[2 * k + 1] = make_float2(rq[k].z, rq[k].w);
    }
    issue(it + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) cmac(s, l[i], q[i]);
    if (do_low) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (!((i >> t) & 1)) {
            cmac(lo[t][0], l[i], q[i]);
            cmac(lo[t][1], l[i], q[i | (1 << t)]);
            cmac(lo[t][2], l[i | (1 << t)], q[i]);
          }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (j < nsel) {  // lane bits (uniform branch)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 pq2;
          pq2.x = __shfl_xor_sync(0xffffffffu, q[i].x, 1 << j);
          pq2.y = __shfl_xor_sync(0xffffffffu, q[i].y, 1 << j);
          cmac(hi[j], l[i], pq2);
        }
      }
    }
    if (nsel > 5) {  // warp bits: psi goes through shared memory once
#pragma unroll
      for (int k = 0; k < 4; ++k) xch[k * 256 + tid] = make_float4(q[2 * k].x, q[2 * k].y, q[2 * k + 1].x, q[2 * k + 1].y);
      __syncthreads();
#pragma unroll
      for (int j = 5; j < 7; ++j) {
        if (j < nsel) {
          const int partner = tid ^ (1 << j);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 w = xch[k * 256 + partner];
            cmac(hi[j], l[2 * k], make_float2(w.x, w.y));
            cmac(hi[j], l[2 * k + 1], make_float2(w.z, w.w));
          }
        }
      }
      __syncthreads();
    }
    if (++since_flush == 64) {
      flush();
      since_flush = 0;
    }
  }
  flush();
  __syncthreads();
  // out[t][r][c]: [0][0], [0][1], [1][0] as accumulated, [1][1] = total - [0][0]
  const int tb = 3 + nsel;
  for (int e = tid; e < tb * 8; e += blockDim.x) {
    const int t = e >> 3, c = (e >> 1) & 3, ri = e & 1;
    if (t < t_begin) continue;
    const double v = c < 3 ? sacc[(t * 3 + c) * 2 + ri] : sacc[CR_MAXB * 6 + ri] - sacc[(t * 3) * 2 + ri];
    atomicAdd(out + e, v);
  }
}

int launch_cross_rdm(const void* lam, const void* psi, int nbits, int64_t batch, int nsel, const int* sel_bits,
                     int skip_low, double* out, int64_t out_bstride, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_cross_rdm: nbits=%d", nbits);
  TCB_REQUIRE(batch >= 1 && batch <= 65535, "tcb_sv_cross_rdm: batch=%lld", (long long)batch);
  CrBits cb;
  cb.nlow = nbits < 3 ? nbits : 3;
  cb.nsel = nsel;
  TCB_REQUIRE(nsel >= 0 && nsel <= 7 && cb.nlow + nsel <= nbits, "tcb_sv_cross_rdm: nsel=%d", nsel);
  for (int j = 0; j < 7; ++j) cb.sel[j] = 0;
  for (int j = 0; j < nsel; ++j) {
    cb.sel[j] = sel_bits[j];
    TCB_REQUIRE(sel_bits[j] >= cb.nlow && sel_bits[j] < nbits && (j == 0 || sel_bits[j] > sel_bits[j - 1]),
                "tcb_sv_cross_rdm: selected bits must be ascending, >= %d and < nbits (bit %d)", cb.nlow, sel_bits[j]);
  }
  if (cb.nlow == 3 && nbits >= 3 + nsel + 1) {
    const uint64_t nunits = 1ull << (nbits - 3 - nsel);
    const uint64_t per_cta = 256u >> nsel;
    uint64_t grid = (nunits + per_cta - 1) / per_cta;
    uint64_t cap = (uint64_t)sm_count() * 2;
    if (batch > 1 && cap > (uint64_t)sm_count()) cap = batch >= 4 ? (uint64_t)sm_count() / 2 : sm_count();  // (batch rows share the SMs)
    if (grid > cap) grid = cap;
    dim3 g2((unsigned)grid, (unsigned)batch);
    cross_rdm_reg_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const float2*>(lam), reinterpret_cast<const float2*>(psi),
                                                 nbits, cb, skip_low ? 3 : 0, out, 2 * out_bstride);
    TCB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const uint64_t ntiles = 1ull << (nbits - cb.nlow - nsel);
  uint64_t grid = (uint64_t)sm_count() * 4;
  if (grid > ntiles) grid = ntiles;
  if (batch > 1 && grid > (uint64_t)sm_count()) grid = sm_count();  // (batch rows share the SMs)
  dim3 g2((unsigned)grid, (unsigned)batch);
  cross_rdm_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const float2*>(lam), reinterpret_cast<const float2*>(psi),
                                           nbits, cb, skip_low ? cb.nlow : 0, out, 2 * out_bstride);
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// pack / unpack the half of the state with local bit == want (global<->local qubit swap)
__global__ void __launch_bounds__(256)
pack_half_kernel(const float2* __restrict__ state, float2* __restrict__ buf, uint64_t nhalf, int bit,
                 int want, int unpack, float2* state_w) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nhalf; i += stride) {
    const uint64_t x = insert_zero(i, bit) | ((uint64_t)want << bit);
    if (unpack)
      state_w[x] = buf[i];
    else
      buf[i] = state[x];
  }
}

int launch_pack_half(const void* state, void* buf, int nbits, int bit, int want, int unpack,
                     cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40 && bit >= 0 && bit < nbits, "tcb_sv_pack_half: bad bit %d", bit);
  const uint64_t nhalf = 1ull << (nbits - 1);
  pack_half_kernel<<<grid_for(nhalf, 256), 256, 0, stream>>>(
      reinterpret_cast<const float2*>(state), reinterpret_cast<float2*>(buf), nhalf, bit, want & 1,
      unpack, reinterpret_cast<float2*>(const_cast<void*>(state)));
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// pack / unpack the sub-block of the state whose `nsel` selected local bits equal `pattern`
// (m-qubit global<->local swap: one such block goes to each of the 2^m - 1 partner ranks).
// Elements [first, first + count) of the block, in block order; V = float4 moves amplitude pairs
// (legal when bit 0 is not selected and first, count are even).
struct SelBits {
  int n;
  int pos[8];  // ascending
};
template <typename V>
__global__ void __launch_bounds__(256)
pack_bits_kernel(V* __restrict__ state, V* __restrict__ buf, uint64_t first, uint64_t count, SelBits sb,
                 uint64_t patmask, int unpack) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // four independent elements per thread and iteration: when `buf` is a peer GPU's staging buffer (sharded.py) the
  // stores are NVLink writes, and the wire wants many of them in flight
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < count; i0 += 4 * stride) {
    uint64_t xs[4];
    V vals[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t x = first + i0 + u * stride;
#pragma unroll 1
      for (int k = 0; k < sb.n; ++k) x = insert_zero(x, sb.pos[k]);
      xs[u] = x | patmask;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < count) vals[u] = unpack ? buf[i] : state[xs[u]];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < count) {
        if (unpack)
          state[xs[u]] = vals[u];
        else
          buf[i] = vals[u];
      }
    }
  }
}

int launch_pack_bits(void* state, void* buf, int nbits, int nsel, const int* sel_bits, uint64_t pattern,
                     uint64_t first, uint64_t count, int unpack, cudaStream_t stream) {
  TCB_REQUIRE(nbits >= 1 && nbits <= 40, "tcb_sv_pack_bits: nbits=%d", nbits);
  TCB_REQUIRE(nsel >= 1 && nsel <= 8 && nsel <= nbits, "tcb_sv_pack_bits: nsel=%d out of range", nsel);
  SelBits sb;
  sb.n = nsel;
  uint64_t patmask = 0;
  for (int k = 0; k < nsel; ++k) {
    TCB_REQUIRE(sel_bits[k] >= 0 && sel_bits[k] < nbits, "tcb_sv_pack_bits: bit %d out of range", sel_bits[k]);
    TCB_REQUIRE(k == 0 || sel_bits[k] > sel_bits[k - 1], "tcb_sv_pack_bits: selected bits must be ascending");
    sb.pos[k] = sel_bits[k];
    patmask |= ((pattern >> k) & 1ull) << sel_bits[k];
  }
  const uint64_t block = 1ull << (nbits - nsel);
  TCB_REQUIRE(first + count <= block, "tcb_sv_pack_bits: range exceeds the block");
  if (count == 0) return 0;
  if (sel_bits[0] >= 1 && (first & 1ull) == 0 && (count & 1ull) == 0) {
    for (int k = 0; k < nsel; ++k) sb.pos[k] -= 1;
    pack_bits_kernel<float4><<<grid_for(count / 2, 256), 256, 0, stream>>>(
        reinterpret_cast<float4*>(state), reinterpret_cast<float4*>(buf), first / 2, count / 2, sb, patmask >> 1,
        unpack);
  } else {
    pack_bits_kernel<float2><<<grid_for(count, 256), 256, 0, stream>>>(
        reinterpret_cast<float2*>(state), reinterpret_cast<float2*>(buf), first, count, sb, patmask, unpack);
  }
  TCB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tcb

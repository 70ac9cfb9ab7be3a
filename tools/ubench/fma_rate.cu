// Developer micro-benchmark: issue rate of the packed FP32 instructions the pass kernel is made of.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(float2* out, float2 m0, float2 m1, int n) {
  float2 a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, 1.f - i * 1e-2f);
  float2 n0 = make_float2(-m0.y, m0.x), n1 = make_float2(-m1.y, m1.x);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) {  // plain packed FMA, pair operands only
        a[i] = __ffma2_rn(a[i], m0, n0);
        a[i + 1] = __ffma2_rn(a[i + 1], m1, n1);
      } else if (MODE == 1) {  // the 2x2 complex matvec of the pass kernel (8 packed per pair)
        float2 x = a[i], y = a[i + 1];
        float2 r0 = __fmul2_rn(make_float2(x.x, x.x), m0);
        float2 r1 = __fmul2_rn(make_float2(x.x, x.x), m1);
        r0 = __ffma2_rn(make_float2(x.y, x.y), n0, r0);
        r1 = __ffma2_rn(make_float2(x.y, x.y), n1, r1);
        r0 = __ffma2_rn(make_float2(y.x, y.x), m1, r0);
        r1 = __ffma2_rn(make_float2(y.x, y.x), m0, r1);
        r0 = __ffma2_rn(make_float2(y.y, y.y), n1, r0);
        r1 = __ffma2_rn(make_float2(y.y, y.y), n0, r1);
        a[i] = r0;
        a[i + 1] = r1;
      } else if (MODE == 2) {  // scalar FFMA
        a[i].x = fmaf(a[i].x, m0.x, m0.y);
        a[i].y = fmaf(a[i].y, m0.x, m0.y);
        a[i + 1].x = fmaf(a[i + 1].x, m1.x, m1.y);
        a[i + 1].y = fmaf(a[i + 1].y, m1.x, m1.y);
      } else {  // N-form: 4 packed per pair
        float2 x = a[i], y = a[i + 1];
        float2 t0 = __ffma2_rn(make_float2(y.y, y.y), n0, x);
        float2 t1 = __ffma2_rn(make_float2(x.y, x.y), n1, y);
        a[i] = __ffma2_rn(make_float2(y.x, y.x), m0, t0);
        a[i + 1] = __ffma2_rn(make_float2(x.x, x.x), m1, t1);
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double packed_per_iter, double flops_per_iter) {
  float2* out; cudaMalloc(&out, 148 * 4 * 128 * sizeof(float2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 128>>>(out, make_float2(0.999f, 0.01f), make_float2(0.01f, 0.999f), 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 128>>>(out, make_float2(0.999f, 0.01f), make_float2(0.01f, 0.999f), ITERS);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double thr = 148.0 * 4 * 128;
  printf("%-32s %8.3f ms  %7.2f TFLOP/s  %6.3f instr/clk/SMSP (at 1.965 GHz)\n", name, ms,
         thr * ITERS * flops_per_iter / ms / 1e9, thr / 32 * ITERS * packed_per_iter / (ms * 1e-3 * 1.965e9) / (148 * 4));
  cudaFree(out);
}
int main() {
  run<0>("FFMA2 pair operands", 16, 16 * 4);
  run<1>("2x2 complex matvec (8/pair)", 64, 64 * 4);
  run<2>("scalar FFMA", 32, 32 * 2);
  run<3>("N-form (4/pair)", 32, 32 * 4);
  return 0;
}

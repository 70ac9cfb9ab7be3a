// Developer micro-benchmark: throughput of cvt.rna.tf32.f32 vs an integer round-and-mask (tn_gemm_tc.cu split).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, float s, int n) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i + s;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        unsigned r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a[i]));
        a[i] = __uint_as_float(r) + s;
      } else if (MODE == 1) {
        unsigned r = (__float_as_uint(a[i]) + 0x1000u) & 0xffffe000u;
        a[i] = __uint_as_float(r) + s;
      } else {
        a[i] = a[i] * 1.0001f + s;
      }
    }
  }
  float t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE>
void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 2 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 256>>>(out, 1e-7f, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 256>>>(out, 1e-7f, ITERS);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 2 * 256 * 16.0 * ITERS;
  printf("%-28s %8.3f ms  %6.1f elem/clk/SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3 * 1.965e9) / 148);
}
int main() { run<0>("cvt.rna.tf32 + fadd"); run<1>("iadd + lop3 + fadd"); run<2>("ffma only"); return 0; }

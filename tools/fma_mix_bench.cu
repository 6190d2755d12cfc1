// Microbenchmark: can plain FFMA (fmalite) issue beside FFMA2 (fmaheavy) on sm_100a?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/fma_mix tools/fma_mix_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int N2, int N1>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s) {
  float2 a2[N2 > 0 ? N2 : 1];
  float a1[N1 > 0 ? N1 : 1];
  for (int i = 0; i < N2; ++i) a2[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int i = 0; i < N1; ++i) a1[i] = threadIdx.x * 1e-3f - i;
  float2 m2 = make_float2(s, s * 0.999f);
  float2 c2 = make_float2(1e-3f, 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < (N2 > N1 ? N2 : N1); ++i) {
        if (i < N2) a2[i] = __ffma2_rn(a2[i], m2, c2);
        if (i < N1) a1[i] = fmaf(a1[i], s, 1e-3f);
      }
    }
  }
  float acc = 0;
  for (int i = 0; i < N2; ++i) acc += a2[i].x + a2[i].y;
  for (int i = 0; i < N1; ++i) acc += a1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int N2, int N1>
void run(const char* name, float* d) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<N2, N1><<<148, 512>>>(d, 16, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<N2, N1><<<148, 512>>>(d, iters, 0.999f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)148 * 512 * iters * 4 * (2.0 * N2 + N1);
  printf("%-28s N2=%2d N1=%2d  %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM @1.965GHz\n", name, N2, N1, ms,
         fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.965);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 512 * 4);
  run<16, 0>("FFMA2 only", d);
  run<0, 16>("FFMA only", d);
  run<0, 32>("FFMA only", d);
  run<16, 16>("FFMA2:FFMA 1:1", d);
  run<16, 8>("FFMA2:FFMA 2:1", d);
  run<8, 16>("FFMA2:FFMA 1:2", d);
  run<24, 8>("FFMA2:FFMA 3:1", d);
  cudaError_t e = cudaGetLastError();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}

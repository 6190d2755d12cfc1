// Microbenchmark: throughput of FFMA2 operand forms on sm_100a:
//   0: all-vector  acc = a2 * m2 + acc       1: scalar-broadcast  acc = (s,s) * m2 + acc  (R.F32 form)
//   2: scalar-broadcast with the vector operand from the constant bank (UR.F32x2 form)
#include <cuda_runtime.h>
#include <stdio.h>
__constant__ float2 cw[64];
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s, const float* in) {
  float2 acc[20];
  for (int i = 0; i < 20; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float sc[4];
  for (int i = 0; i < 4; ++i) sc[i] = in[threadIdx.x + 32 * i];
  float2 m2[5];
  for (int i = 0; i < 5; ++i) m2[i] = make_float2(s + i, s * 0.999f - i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          if (MODE == 0) acc[p * 5 + c] = __ffma2_rn(make_float2(sc[p], sc[(p + 1) & 3]), m2[c], acc[p * 5 + c]);
          if (MODE == 1) acc[p * 5 + c] = __ffma2_rn(make_float2(sc[p], sc[p]), m2[c], acc[p * 5 + c]);
          if (MODE == 2) acc[p * 5 + c] = __ffma2_rn(make_float2(sc[p], sc[p]), cw[(it & 3) * 16 + r * 4 + c], acc[p * 5 + c]);
        }
    }
  }
  float a = 0;
  for (int i = 0; i < 20; ++i) a += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
template <int MODE>
void run(const char* name, float* d, const float* in) {
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 512>>>(d, 16, 0.999f, in);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148, 512>>>(d, iters, 0.999f, in);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)148 * 512 * iters * 4 * 40.0;
  printf("%-28s %.3f ms  %.1f FMA/clk/SM @1.965GHz\n", name, ms, fma / ms / 1e6 / 148 / 1.965);
}
int main() {
  float *d, *in; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  run<0>("vector x vector", d, in);
  run<1>("scalar(R.F32) x vector", d, in);
  run<2>("scalar(R.F32) x const(UR)", d, in);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

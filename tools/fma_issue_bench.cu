// Microbenchmark: does FFMA2 hold the SMSP issue port for 2 cycles?  Interleave NL shared-memory loads
// (LSU pipe; results unused) per 16 FFMA2 (or per 32 FFMA) and watch the FMA rate.
#include <cuda_runtime.h>
#include <stdio.h>
template <int PACKED, int NL>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s) {
  __shared__ float sm[1024];
  sm[threadIdx.x] = s; sm[threadIdx.x + 512] = s;
  __syncthreads();
  const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]);
  float2 a2[16];
  for (int i = 0; i < 16; ++i) a2[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 m2 = make_float2(s, s * 0.999f);
  float2 c2 = make_float2(1e-3f, 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (PACKED) a2[i] = __ffma2_rn(a2[i], m2, c2);
        else { a2[i].x = fmaf(a2[i].x, m2.x, c2.x); a2[i].y = fmaf(a2[i].y, m2.y, c2.y); }
        if ((i * NL) / 16 != ((i + 1) * NL) / 16) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(addr + 4 * (i & 7)));
        }
      }
    }
  }
  float acc = 0;
  for (int i = 0; i < 16; ++i) acc += a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int PACKED, int NL>
void run(const char* name, float* d) {
  const int iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PACKED, NL><<<148, 512>>>(d, 16, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<PACKED, NL><<<148, 512>>>(d, iters, 0.999f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)148 * 512 * iters * 4 * 32.0;
  printf("%-6s LDS per 32 FMA = %2d : %.3f ms  %.1f FMA/clk/SM @1.965GHz\n", name, NL, ms,
         fma / ms / 1e6 / 148 / 1.965);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 512 * 4);
  run<1, 0>("FFMA2", d); run<1, 2>("FFMA2", d); run<1, 4>("FFMA2", d); run<1, 8>("FFMA2", d); run<1, 16>("FFMA2", d);
  run<0, 0>("FFMA", d);  run<0, 2>("FFMA", d);  run<0, 4>("FFMA", d);  run<0, 8>("FFMA", d);  run<0, 16>("FFMA", d);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

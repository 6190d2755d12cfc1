// Hardware experiment for a tensor-core joint with single-write operands: does tcgen05.mma kind::f16
// accept MN-major SWIZZLE_64B operands whose MN atoms OVERLAP -- atom stride (LBO) of one pixel (64 bytes) -- so that the
// N rows of B are (pixel shift s, slot) for a pixel-major tile [pixel][32 fp16 slots] written ONCE?
//   D[(a, xs), (s, ys)] = sum_{p < 16} X[a][p0 + p][xs] * Y[p0 + p + s][ys]      a < 4 rows, s < 3 shifts, 32 slots each
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mi-based-regularized-semi-supervised-segmentation_b200/csrc -o tools/_bin/mn_major_micro tools/mn_major_micro.cu
//   tools/_bin/mn_major_micro [swizzle_mode_bits=4] [shift_stride_bytes=64]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tma.cuh"
using namespace iic;

constexpr int NPX = 32;          // pixels staged per row buffer
constexpr int ROWB = NPX * 64;   // bytes of one row buffer [pixel][64 B]

__device__ __forceinline__ uint32_t swz64(uint32_t byte_addr) { return byte_addr ^ (((byte_addr >> 7) & 3u) << 4); }

__global__ void k(const __half* X, const __half* Y, float* D, int swz_bits, int shift_stride, int p0) {
  extern __shared__ __align__(1024) unsigned char sm_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  unsigned char* xa = sm;                 // 4 row buffers
  unsigned char* yb = sm + 4 * ROWB;      // 1 row buffer
  const uint32_t sbase = smem_u32(sm);
  // write operands: element (row a, pixel p, slot t) at byte a*ROWB + p*64 + t*2, XOR-swizzled on the ABSOLUTE address
  for (int e = threadIdx.x; e < 4 * NPX * 32; e += blockDim.x) {
    const int t = e % 32, p = (e / 32) % NPX, a = e / (32 * NPX);
    uint32_t off = a * ROWB + p * 64 + t * 2;
    uint32_t addr = sbase + off;
    if (swz_bits) addr = swz64(addr);
    *reinterpret_cast<__half*>(sm + (addr - sbase)) = X[e];
  }
  for (int e = threadIdx.x; e < NPX * 32; e += blockDim.x) {
    const int t = e % 32, p = e / 32;
    uint32_t addr = sbase + 4 * ROWB + p * 64 + t * 2;
    if (swz_bits) addr = swz64(addr);
    *reinterpret_cast<__half*>(sm + (addr - sbase)) = Y[e];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t saddr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((saddr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)swz_bits << 61;
      return d;
    };
    // MN-major, SWIZZLE_64B: LBO = stride between MN atoms (32 slots), SBO = stride between groups of 8 K rows (pixels)
    const uint64_t da = desc(smem_u32(xa) + p0 * 64, ROWB, 512);
    const uint64_t db = desc(smem_u32(yb) + p0 * 64, shift_stride, 512);
    // kind::f16: D f32 (bit 4), A/B fp16 (0), a_major (bit 15) = b_major (bit 16) = 1 (MN-major), N = 96, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((96u >> 3) << 17) | (8u << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0, 1);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wid < 4) {
    for (int c = 0; c < 96; c += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)(wid * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int q = 0; q < 8; ++q) D[(wid * 32 + lane) * 96 + c + q] = __uint_as_float(v[q]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main(int argc, char** argv) {
  const int swz = argc > 1 ? atoi(argv[1]) : 4, shift_stride = argc > 2 ? atoi(argv[2]) : 64, p0 = argc > 3 ? atoi(argv[3]) : 0;
  std::vector<__half> hx(4 * NPX * 32), hy(NPX * 32);
  std::vector<float> fx(hx.size()), fy(hy.size());
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 17 - 8) / 8.f; hx[i] = __float2half(fx[i]); }
  for (size_t i = 0; i < hy.size(); ++i) { fy[i] = (float)(rand() % 13 - 6) / 4.f; hy[i] = __float2half(fy[i]); }
  __half *dx, *dy; float* dd;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dy, hy.size() * 2); cudaMalloc(&dd, 128 * 96 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, 128 * 96 * 4);
  const int smem = 5 * ROWB + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(dx, dy, dd, swz, shift_stride, p0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("swz=%d shift_stride=%d p0=%d: %s\n", swz, shift_stride, p0, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> d(128 * 96);
  cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
  const int shift_px = shift_stride / 64;
  int bad = 0;
  double maxerr = 0;
  for (int a = 0; a < 4; ++a) for (int xs = 0; xs < 32; ++xs) for (int s = 0; s < 3; ++s) for (int ys = 0; ys < 32; ++ys) {
    double ref = 0;
    for (int p = 0; p < 16; ++p) ref += (double)fx[(a * NPX + p0 + p) * 32 + xs] * fy[(p0 + p + s * shift_px) * 32 + ys];
    const double got = d[(a * 32 + xs) * 96 + s * 32 + ys];
    const double err = fabs(got - ref);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3 && bad < 8) { printf("  mismatch a=%d xs=%d s=%d ys=%d: got %g ref %g\n", a, xs, s, ys, got, ref); }
    bad += err > 1e-3;
  }
  printf("%s: %d of %d entries wrong, max err %g\n", bad ? "FAIL" : "PASS", bad, 128 * 96, maxerr);
  return 0;
}

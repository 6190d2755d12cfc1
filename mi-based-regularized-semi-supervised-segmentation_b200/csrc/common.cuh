// Shared helpers for libiic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/iic_b200.h"

namespace iic {

// ---- error plumbing: nothing throws across the C ABI ------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define IIC_CHECK_CUDA(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::iic::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                       cudaGetErrorString(_e));                                            \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define IIC_REQUIRE(cond, ...)                                                             \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::iic::set_error(__VA_ARGS__);                                                       \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

int sm_count_cached(int device);   // <0 on error
int current_device();

// ---- patch windows (contrastyou/losses/iic_loss.py:152-160) -----------------------------------------
struct PatchGrid {
  int H, W;        // full map
  int ph, pw;      // effective patch extent = min(patch, map)
  int sh, sw;      // steps
  int nh, nw;      // windows per axis
};

__host__ __device__ inline int patch_axis_count(int extent, int patch, int step) {
  // np.arange(0, extent - patch, step) has ceil((extent-patch)/step) entries when extent > patch,
  // then one more window flush with the border is appended.
  if (extent <= patch) return 1;
  return (extent - patch + step - 1) / step + 1;
}
__host__ __device__ inline int patch_axis_origin(int idx, int count, int extent, int patch, int step) {
  if (idx < count - 1) return idx * step;
  int o = extent - patch;
  return o > 0 ? o : 0;
}
__host__ inline bool make_patch_grid(int H, int W, int patch_h, int patch_w, int step_h, int step_w,
                                     PatchGrid* g) {
  if (H <= 0 || W <= 0 || patch_h <= 0 || patch_w <= 0) return false;
  if ((H > patch_h && step_h <= 0) || (W > patch_w && step_w <= 0)) return false;
  g->H = H; g->W = W;
  g->ph = patch_h < H ? patch_h : H;
  g->pw = patch_w < W ? patch_w : W;
  g->sh = step_h > 0 ? step_h : 1;
  g->sw = step_w > 0 ? step_w : 1;
  g->nh = patch_axis_count(H, patch_h, g->sh);
  g->nw = patch_axis_count(W, patch_w, g->sw);
  return true;
}

struct View4 {          // (B, K, H, W) float32 view, W-stride 1
  const float* p;
  long long sn, sc, sh;
};

// ---- warp / block reductions (fixed order -> deterministic) ------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of doubles; `scratch` has >= 33 doubles of shared memory. All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
// NaN-propagating block-wide min (fmin would drop NaNs; a NaN joint must poison the loss as in torch.min)
__device__ __forceinline__ double block_min_nan(double v, bool has_nan, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  unsigned any_nan = __ballot_sync(0xffffffffu, has_nan);
  __syncthreads();
  if (lane == 0) scratch[wid] = any_nan ? __longlong_as_double(0x7ff8000000000000LL) : v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? scratch[lane] : __longlong_as_double(0x7ff0000000000000LL);
    bool n2 = (t != t);
    unsigned an = __ballot_sync(0xffffffffu, n2);
    t = warp_min(n2 ? __longlong_as_double(0x7ff0000000000000LL) : t);
    if (lane == 0) scratch[32] = an ? __longlong_as_double(0x7ff8000000000000LL) : t;
  }
  __syncthreads();
  return scratch[32];
}

}  // namespace iic

// Per-pixel register helpers shared by the streaming kernels that keep all channels of V consecutive pixels in
// registers (csrc/sup.cu, csrc/flip.cu): channel-row loads (16-byte when V = 4), int64 label loads, channel softmax.
#pragma once
#include "common.cuh"

namespace iic {

constexpr int SUP_CMAX = 8;      // channels kept in registers by the generic (run-time C) instantiations

template <int V, int NC>
__device__ __forceinline__ void sup_load(const float* __restrict__ src, long long inner, int C,
                                         float (&v)[NC][V]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < C) {
      if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src + (long long)c * inner));
        v[c][0] = t.x; v[c][1 % V] = t.y; v[c][2 % V] = t.z; v[c][3 % V] = t.w;
      } else {
        v[c][0] = __ldg(src + (long long)c * inner);
      }
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) v[c][e] = -INFINITY;
    }
  }
}

template <int V>
__device__ __forceinline__ void sup_load_labels(const long long* __restrict__ src, long long (&l)[V]) {
  if (V == 4) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(src));
    const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(src) + 1);
    l[0] = a.x; l[1 % V] = a.y; l[2 % V] = b.x; l[3 % V] = b.y;
  } else {
    l[0] = __ldg(src);
  }
}

// softmax over the channels of pixel e, in place; returns the index of the first maximal logit
template <int V, int NC>
__device__ __forceinline__ int sup_softmax(float (&v)[NC][V], int e) {
  float mx = v[0][e];
  int arg = 0;
#pragma unroll
  for (int c = 1; c < NC; ++c) {
    if (v[c][e] > mx) { mx = v[c][e]; arg = c; }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) { v[c][e] = __expf(v[c][e] - mx); s += v[c][e]; }   // exp(-inf) = 0 pads
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c][e] *= inv;
  return arg;
}

// CTAs per sample (grid = (gx, outer)) for one resident wave of `ctas_per_sm` 256-thread CTAs per SM, with every
// thread walking the same number of pixel groups: k = iterations per thread, then just enough CTAs for k.
// gx * outer <= max(sms * ctas_per_sm, outer).
inline int pixel_ctas_per_sample(long long outer, long long groups, int ctas_per_sm) {
  int sms = sm_count_cached(current_device());
  if (sms <= 0) sms = 148;
  long long cap = ((long long)sms * ctas_per_sm) / outer;
  if (cap < 1) cap = 1;
  const long long k = (groups + 256 * cap - 1) / (256 * cap);
  long long gx = (groups + 256 * k - 1) / (256 * k);
  if (gx < 1) gx = 1;
  return (int)gx;
}

}  // namespace iic

// Supervised branch of the udaiic iteration: one streaming pass each way (HBM-bound).
//   loss : sup_loss = KL_div()(label_logits.softmax(1), class2one_hot(labeled_target.squeeze(1), C))
//          semi_seg/epocher.py:165-166; dc2:deepclustering2/loss/kl_losses.py:107-126 (reduction "mean");
//          dc2:deepclustering2/utils/assertion.py:101-116 class2one_hot (a `long` one-hot; its sset assert -> BAD_LABEL flag)
//   dice : the tensors UniversalDice.add appends for (label_logits.max(1)[1], labeled_target.squeeze(1))
//          semi_seg/epocher.py:183-184; dc2:deepclustering2/meters2/individual_meters/general_dice_meter.py:41-95
//          (_intersaction = sum(pred*target), _union = sum(pred+target) over the pixels of one sample)
// With a one-hot target only the labelled class contributes: -1 * log((p_l + eps) / (1 + eps)) * w_l, the other
// classes give -0 * log(finite) = 0 (eps > 0).  The softmax, the one-hot and the argmax live in registers; the
// (B,C,H,W) int64 one-hot (8*C bytes per pixel) and the probability map are never materialised.
// A thread owns V consecutive pixels of one sample (V = 4 with 16-byte loads when the rows allow it) and keeps NC
// channels in registers: NC = C for the usual 2..4 classes (no padded work), NC = 8 with run-time guards otherwise.
#include "common.cuh"
#include "pixel_common.cuh"

namespace iic {

struct SupWorkspace {
  unsigned int ticket;   // self-resetting arrival counter
  unsigned int pad_;
  // followed by double partial[gridDim.x * gridDim.y]
};

// grid (gx, outer): blockIdx.y = sample, so the Dice counters of a CTA belong to one sample
template <int V, int NC>
__global__ void __launch_bounds__(256, NC <= 4 ? 3 : 2) sup_fwd_kernel(const float* __restrict__ logits,
                                                      const long long* __restrict__ labels, int C_rt,
                                                      long long inner, float eps, const float* __restrict__ weight,
                                                      double denom, float* __restrict__ loss_out,
                                                      long long* __restrict__ dice_out, int* __restrict__ flags,
                                                      SupWorkspace* ws) {
  __shared__ double scratch[40];
  __shared__ unsigned int s_cnt[2][SUP_CMAX];
  __shared__ bool is_last;
  double* partial = reinterpret_cast<double*>(ws + 1);
  const int C = NC < SUP_CMAX ? NC : C_rt;      // the host picks NC == C for C <= 4
  const long long o = blockIdx.y, outer = gridDim.y;
  const float* lg = logits + o * C * inner;
  const long long* lb = labels + o * inner;
  if (threadIdx.x < 2 * SUP_CMAX) (&s_cnt[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  float w[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) w[c] = (weight && c < C) ? __ldg(weight + c) : 1.f;
  unsigned int ci[NC], cu[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { ci[c] = 0u; cu[c] = 0u; }
  float local = 0.f;
  bool bad = false;
  const long long ngroups = inner / V;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ngroups;
       q += (long long)gridDim.x * blockDim.x) {
    float p[NC][V];
    long long l[V];
    sup_load<V, NC>(lg + q * V, inner, C, p);
    sup_load_labels<V>(lb + q * V, l);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int arg = sup_softmax<V, NC>(p, e);
      const bool ok = (l[e] >= 0 && l[e] < C);
      bad |= !ok;
      const int li = ok ? (int)l[e] : -1;
      float pl = 1.f, wl = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c == li) { pl = p[c][e]; wl = w[c]; }
        ci[c] += (unsigned)((c == arg) & (c == li));
        cu[c] += (unsigned)(c == arg) + (unsigned)(c == li);
      }
      // -target * log((prob + eps) / (target + eps)) * weight with target = 1 (kl_losses.py:114-118)
      local += -logf((pl + eps) / (1.f + eps)) * wl;
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, IIC_FLAG_BAD_LABEL);
  if (dice_out != nullptr) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < C) {
        const unsigned a = __reduce_add_sync(0xffffffffu, ci[c]);
        const unsigned b = __reduce_add_sync(0xffffffffu, cu[c]);
        if ((threadIdx.x & 31) == 0) {
          if (a) atomicAdd(&s_cnt[0][c], a);
          if (b) atomicAdd(&s_cnt[1][c], b);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * SUP_CMAX) {
      const int which = threadIdx.x / SUP_CMAX, c = threadIdx.x % SUP_CMAX;
      const unsigned v = s_cnt[which][c];
      // dice_out[which][sample][class]; integer atomics: the result does not depend on the order
      if (c < C && v)
        atomicAdd(reinterpret_cast<unsigned long long*>(dice_out) + ((long long)which * outer + o) * C + c,
                  (unsigned long long)v);
    }
  }
  const double bsum = block_sum((double)local, scratch);
  const unsigned int ncta = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    partial[blockIdx.y * gridDim.x + blockIdx.x] = bsum;
    __threadfence();
    is_last = (atomicAdd(&ws->ticket, 1u) == ncta - 1);
  }
  __syncthreads();
  if (is_last) {
    // the last CTA adds the per-CTA sums: thread t takes partial[t], partial[t+256], ... in order, then a
    // fixed-shape block reduction -> deterministic
    __threadfence();
    const volatile double* pp = partial;
    double tsum = 0.0;
    for (unsigned int b = threadIdx.x; b < ncta; b += blockDim.x) tsum += pp[b];
    double tot = block_sum(tsum, scratch);
    if (threadIdx.x == 0) {
      tot /= denom;
      loss_out[0] = (float)tot;
      if (tot != tot) atomicOr(flags, IIC_FLAG_NAN_LOSS);
      ws->ticket = 0;
    }
  }
}

// d loss / d logits: gp_l = -w_l / (p_l + eps) * g / (outer*inner) on the labelled class only, then the softmax
// adjoint p_c * (gp_c - sum_k gp_k p_k).  A pixel with an out-of-range label gets a zero gradient.
template <int V, int NC>
__global__ void __launch_bounds__(256, NC <= 4 ? 3 : 2) sup_bwd_kernel(const float* __restrict__ logits,
                                                      const long long* __restrict__ labels, int C_rt,
                                                      long long inner, float eps, const float* __restrict__ weight,
                                                      float inv_denom, const float* __restrict__ grad_loss,
                                                      float* __restrict__ grad_logits) {
  const int C = NC < SUP_CMAX ? NC : C_rt;
  const long long o = blockIdx.y;
  const float* lg = logits + o * C * inner;
  const long long* lb = labels + o * inner;
  float* gl = grad_logits + o * C * inner;
  const float g = (grad_loss ? __ldg(grad_loss) : 1.f) * inv_denom;
  float w[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) w[c] = (weight && c < C) ? __ldg(weight + c) : 1.f;
  const long long ngroups = inner / V;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ngroups;
       q += (long long)gridDim.x * blockDim.x) {
    float p[NC][V];
    long long l[V];
    sup_load<V, NC>(lg + q * V, inner, C, p);
    sup_load_labels<V>(lb + q * V, l);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      sup_softmax<V, NC>(p, e);
      const int li = (l[e] >= 0 && l[e] < C) ? (int)l[e] : -1;
      float pl = 1.f, wl = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c == li) { pl = p[c][e]; wl = w[c]; }
      }
      const float gpl = -wl / (pl + eps) * g;       // 0 when the label is out of range (wl = 0)
      const float dot = gpl * pl;
#pragma unroll
      for (int c = 0; c < NC; ++c) p[c][e] = p[c][e] * ((c == li ? gpl : 0.f) - dot);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < C) {
        if (V == 4)
          *reinterpret_cast<float4*>(gl + (long long)c * inner + q * V) =
              make_float4(p[c][0], p[c][1 % V], p[c][2 % V], p[c][3 % V]);
        else
          gl[(long long)c * inner + q] = p[c][0];
      }
    }
  }
}

static bool sup_vec_ok(const void* a, const void* b, const void* c, long long inner) {
  if (inner % 4 != 0) return false;
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

}  // namespace iic

using namespace iic;

extern "C" size_t iic_sup_workspace_bytes(int device, long long outer) {
  int sms = sm_count_cached(device);
  if (sms <= 0) sms = 148;
  long long n = (long long)sms * 8;
  if (outer > n) n = outer;
  return sizeof(SupWorkspace) + (size_t)n * sizeof(double);
}

extern "C" int iic_sup_forward(const float* logits, const long long* labels, long long outer, int C,
                               long long inner, double eps, const float* weight, float* loss_out,
                               long long* dice_out, int* flags, void* workspace, void* stream) {
  IIC_REQUIRE(logits && labels && loss_out && flags && workspace, "iic_sup_forward: null pointer");
  IIC_REQUIRE(outer > 0 && C > 0 && inner > 0, "iic_sup_forward: empty tensor");
  IIC_REQUIRE(outer <= 65535, "iic_sup_forward: outer=%lld > 65535 samples unsupported", outer);
  IIC_REQUIRE(C <= SUP_CMAX, "iic_sup_forward: C=%d > %d classes unsupported", C, SUP_CMAX);
  IIC_REQUIRE(eps > 0.0, "iic_sup_forward: eps must be > 0 (the reference's loss is NaN at eps = 0)");
  cudaStream_t st = (cudaStream_t)stream;
  if (dice_out) IIC_CHECK_CUDA(cudaMemsetAsync(dice_out, 0, (size_t)2 * outer * C * sizeof(long long), st));
  const double denom = (double)outer * (double)inner;
  const bool vec = sup_vec_ok(logits, labels, logits, inner);
  const dim3 grid(pixel_ctas_per_sample(outer, vec ? inner / 4 : inner, C <= 4 ? 3 : 2), (unsigned)outer);
#define IIC_SUP_FWD(VV, NN)                                                                                      \
  sup_fwd_kernel<VV, NN><<<grid, 256, 0, st>>>(logits, labels, C, inner, (float)eps, weight, denom, loss_out,    \
                                               dice_out, flags, (SupWorkspace*)workspace)
  if (vec) { if (C == 2) IIC_SUP_FWD(4, 2); else if (C == 3) IIC_SUP_FWD(4, 3); else if (C == 4) IIC_SUP_FWD(4, 4); else IIC_SUP_FWD(4, SUP_CMAX); }
  else     { if (C == 2) IIC_SUP_FWD(1, 2); else if (C == 3) IIC_SUP_FWD(1, 3); else if (C == 4) IIC_SUP_FWD(1, 4); else IIC_SUP_FWD(1, SUP_CMAX); }
#undef IIC_SUP_FWD
  IIC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int iic_sup_backward(const float* logits, const long long* labels, long long outer, int C,
                                long long inner, double eps, const float* weight, const float* grad_loss,
                                float* grad_logits, void* stream) {
  IIC_REQUIRE(logits && labels && grad_logits, "iic_sup_backward: null pointer");
  IIC_REQUIRE(outer > 0 && C > 0 && inner > 0, "iic_sup_backward: empty tensor");
  IIC_REQUIRE(outer <= 65535, "iic_sup_backward: outer=%lld > 65535 samples unsupported", outer);
  IIC_REQUIRE(C <= SUP_CMAX, "iic_sup_backward: C=%d > %d classes unsupported", C, SUP_CMAX);
  IIC_REQUIRE(eps > 0.0, "iic_sup_backward: eps must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_denom = (float)(1.0 / ((double)outer * (double)inner));
  const bool vec = sup_vec_ok(logits, labels, grad_logits, inner);
  const dim3 grid(pixel_ctas_per_sample(outer, vec ? inner / 4 : inner, C <= 4 ? 3 : 2), (unsigned)outer);
#define IIC_SUP_BWD(VV, NN)                                                                                      \
  sup_bwd_kernel<VV, NN><<<grid, 256, 0, st>>>(logits, labels, C, inner, (float)eps, weight, inv_denom,          \
                                               grad_loss, grad_logits)
  if (vec) { if (C == 2) IIC_SUP_BWD(4, 2); else if (C == 3) IIC_SUP_BWD(4, 3); else if (C == 4) IIC_SUP_BWD(4, 4); else IIC_SUP_BWD(4, SUP_CMAX); }
  else     { if (C == 2) IIC_SUP_BWD(1, 2); else if (C == 3) IIC_SUP_BWD(1, 3); else if (C == 4) IIC_SUP_BWD(1, 4); else IIC_SUP_BWD(1, SUP_CMAX); }
#undef IIC_SUP_BWD
  IIC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

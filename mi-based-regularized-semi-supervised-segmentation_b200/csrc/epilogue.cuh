// Block-level bodies of the small-matrix epilogues (float64), shared by the single-term kernels of epilogue.cu and the
// batched kernel of finish.cu:
//   local   min-shift, per-displacement normalise, symmetrise, marginals, entropy and the analytic dL/dJ
//           (contrastyou/losses/iic_loss.py:124-146,186), one CTA per (patch, displacement)
//   global  P = sym(J)/S and the two entropy expressions (iic_loss.py:56-69,91-92), one CTA per term
#pragma once
#include "common.cuh"

namespace iic {

// One (patch, displacement) of a local term.  Jp = the patch's joints [T*T][K][K]; writes this displacement's slices of
// the two backward coefficient tensors (patch base pointers Wxp / Wyp) and returns the displacement's loss (all threads).
// sm: 40 + 3*K doubles of shared memory.
__device__ __forceinline__ double local_epilogue_block(const double* __restrict__ Jp, int K, int T, int d, double lamda,
                                                       double scale, float* __restrict__ Wxp, float* __restrict__ Wyp,
                                                       double* __restrict__ GA_out_d, double* sm) {
  double* scratch = sm;            // 33
  double* marg = sm + 40;          // K : marginal (row sum == column sum of the symmetric P)
  double* lm = marg + K;           // K : log(marg + eps)
  double* gm = lm + K;             // K : log(marg + eps) + marg / (marg + eps)
  const int T2 = T * T;
  const int dy = d / T, dx = d % T;
  const size_t KK = (size_t)K * K;
  const double* Jd = Jp + (size_t)d * KK;
  const double eps = 1e-16;
  const int Kp = (K + 3) & ~3;
  const int tid = threadIdx.x, nt = blockDim.x;

  // 1. m = min over every displacement and both cluster axes of this patch (iic_loss.py:124)
  double mn = __longlong_as_double(0x7ff0000000000000LL);
  bool has_nan = false;
  for (size_t e = tid; e < (size_t)T2 * KK; e += nt) {
    const double v = Jp[e];
    has_nan |= (v != v);
    mn = fmin(mn, v);
  }
  const double m = block_min_nan(mn, has_nan, scratch);

  // 2. A = J_d - m + 1e-16 ; s = sum A ; marginals of P = (A + A^T) / (2 s)
  for (int k = tid; k < K; k += nt) {
    double rs = 0.0, cs = 0.0;
    for (int q = 0; q < K; ++q) {
      rs += Jd[(size_t)k * K + q] - m + 1e-16;
      cs += Jd[(size_t)q * K + k] - m + 1e-16;
    }
    marg[k] = rs + cs;        // scaled by 1/(2s) below
    lm[k] = rs;               // stash the row sum for the total
  }
  __syncthreads();
  double part = 0.0;
  for (int k = tid; k < K; k += nt) part += lm[k];
  const double s = block_sum(part, scratch);
  for (int k = tid; k < K; k += nt) {
    const double mk = marg[k] / (2.0 * s);
    marg[k] = mk;
    lm[k] = log(mk + eps);
    gm[k] = lm[k] + mk / (mk + eps);
  }
  __syncthreads();

  // 3. loss_d and tot = sum_ij GQ_ij Q_ij.  P is symmetric and its row and column marginals agree, so
  //    GP (d loss / d P) is symmetric too and GQ = (GP + GP^T)/2 = GP.
  double l_part = 0.0, t_part = 0.0;
  for (size_t e = tid; e < KK; e += nt) {
    const int i = (int)(e / K), j = (int)(e % K);
    const double a = Jd[e] - m + 1e-16, at = Jd[(size_t)j * K + i] - m + 1e-16;
    const double q = a / s;
    const double p = (a + at) / (2.0 * s);
    const double lp = log(p + eps);
    l_part += -p * (lp - lamda * lm[j] - lamda * lm[i]);
    const double gq = -lp - p / (p + eps) + lamda * (gm[j] + gm[i]);
    t_part += gq * q;
  }
  const double loss_d = block_sum(l_part, scratch);
  const double tot = block_sum(t_part, scratch);

  // 4. GA = dL/dJ_d = (GQ - tot) / s, scaled by 1/(T^2 n_patches); write the two sweep layouts
  //    Wy[cin=i][dy*T+dx][j]            (gy[j] += Wy * x_i shifted by (dy-pad, dx-pad))
  //    Wx[cin=j][(T-1-dy)*T+(T-1-dx)][i] (gx[i] += Wx * y_j shifted by (pad-dy, pad-dx))
  const int dflip = (T - 1 - dy) * T + (T - 1 - dx);
  for (size_t e = tid; e < (size_t)K * Kp; e += nt) {
    const int a_ = (int)(e / Kp), b_ = (int)(e % Kp);   // a_ = cin, b_ = cout (padded)
    float w = 0.f;
    if (b_ < K) {
      // Wy: cin = i = a_, cout = j = b_ ; Wx: cin = j = a_, cout = i = b_.  GA is symmetric in (i,j)
      // only through GQ; (GQ - tot)/s is symmetric as well, so one evaluation serves both.
      const int i = a_, j = b_;
      const double a = Jd[(size_t)i * K + j] - m + 1e-16, at = Jd[(size_t)j * K + i] - m + 1e-16;
      const double p = (a + at) / (2.0 * s);
      const double lp = log(p + eps);
      const double gq = -lp - p / (p + eps) + lamda * (gm[j] + gm[i]);
      const double ga = (gq - tot) / s * scale;
      w = (float)ga;
      if (GA_out_d) GA_out_d[(size_t)i * K + j] = ga;
    }
    Wyp[((size_t)a_ * T2 + d) * Kp + b_] = w;
    Wxp[((size_t)a_ * T2 + dflip) * Kp + b_] = w;
  }
  return loss_d;
}

// One global term: P_out (nullable), losses_out[2] (nullable) = loss(lamb), loss_no_lamb.  sm: 40 + 2*K doubles.
__device__ __forceinline__ void global_epilogue_block(const double* __restrict__ J, int K, double lamb, int symmetric,
                                                      float* __restrict__ losses_out, float* __restrict__ P_out,
                                                      int* __restrict__ flags, double* sm) {
  double* scratch = sm;      // 33
  double* pi = sm + 40;      // K  row marginals    p_i = sum_j P[i][j]
  double* pj = pi + K;       // K  column marginals p_j = sum_i P[i][j]
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t KK = (size_t)K * K;
  const double eps = 1e-10;
  auto Jsym = [&](int i, int j) {
    return symmetric ? (J[(size_t)i * K + j] + J[(size_t)j * K + i]) / 2.0 : J[(size_t)i * K + j];
  };
  double part = 0.0;
  for (size_t e = tid; e < KK; e += nt) part += Jsym((int)(e / K), (int)(e % K));
  const double S = block_sum(part, scratch);
  for (int k = tid; k < K; k += nt) {
    double r = 0.0, c = 0.0;
    for (int q = 0; q < K; ++q) {
      r += Jsym(k, q);
      c += Jsym(q, k);
    }
    pi[k] = r / S;
    pj[k] = c / S;
  }
  __syncthreads();
  double l1 = 0.0, l2 = 0.0;
  for (size_t e = tid; e < KK; e += nt) {
    const int i = (int)(e / K), j = (int)(e % K);
    const double p = Jsym(i, j) / S;
    if (P_out) P_out[e] = (float)p;
    if (losses_out) {
      const double lp = log(p + eps), lj = log(pj[j] + eps), li = log(pi[i] + eps);
      l1 += -p * (lp - lamb * lj - lamb * li);
      l2 += -p * (lp - lj - li);
    }
  }
  if (losses_out) {
    const double L1 = block_sum(l1, scratch);
    const double L2 = block_sum(l2, scratch);
    if (tid == 0) {
      losses_out[0] = (float)L1;
      losses_out[1] = (float)L2;
      if (L1 != L1 || L2 != L2) atomicOr(flags, IIC_FLAG_NAN_LOSS);
    }
  }
}

}  // namespace iic

// Stand-alone harness for the row-block tcgen05 backward sweeps at K = 9 or 10, padding 1 (csrc/local_bwd_tcrb10h.cu, the fp16-split form): random source
// maps and random coefficient tensors, a sample of output pixels checked against an fp64 CPU loop.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/tc_bwdrb10h tools/tc_bwdrb10h_harness.cu   [-DIIC_TC_TRACE]
//   tools/_bin/tc_bwdrb [B H W K pad]
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mi-based-regularized-semi-supervised-segmentation_b200/csrc/common.cuh"
namespace iic {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
const char* get_error() { return g_err; }
int current_device() { return 0; }
int sm_count_cached(int) { return 148; }
static Options g_opt;
const Options& options() { return g_opt; }
int ensure_dyn_smem(const void* f, int bytes) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : 1; }
}
#include "../mi-based-regularized-semi-supervised-segmentation_b200/csrc/local_bwd_tcrb10h.cu"

int main(int argc, char** argv) {
  int B = 1, H = 5, W = 32, KC = 10, pad = 1;
  if (argc >= 4) { B = atoi(argv[1]); H = atoi(argv[2]); W = atoi(argv[3]); }
  if (argc >= 6) { KC = atoi(argv[4]); pad = atoi(argv[5]); }
  const int T = 2 * pad + 1, T2 = T * T, Kp4 = (KC + 3) & ~3;
  const size_t n = (size_t)B * KC * H * W, nw = (size_t)KC * T2 * Kp4;
  std::vector<float> hx(n), hy(n), hwx(nw), hwy(nw);
  srand(2);
  for (size_t i = 0; i < n; ++i) {
    const float a = (float)rand() / RAND_MAX, b = (float)rand() / RAND_MAX;
    hx[i] = a * a * a * 0.05f;
    hy[i] = b * b * b * 0.05f;
  }
  for (size_t i = 0; i < nw; ++i) { hwx[i] = (float)rand() / RAND_MAX - 0.4f; hwy[i] = (float)rand() / RAND_MAX - 0.6f; }
  float *dx_, *dy_, *dwx, *dwy, *dgx, *dgy, *dg;
  cudaMalloc(&dx_, n * 4); cudaMalloc(&dy_, n * 4); cudaMalloc(&dgx, n * 4); cudaMalloc(&dgy, n * 4);
  cudaMalloc(&dwx, nw * 4); cudaMalloc(&dwy, nw * 4); cudaMalloc(&dg, 4);
  cudaMemcpy(dx_, hx.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dy_, hy.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dwx, hwx.data(), nw * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dwy, hwy.data(), nw * 4, cudaMemcpyHostToDevice);
  const float g = 0.75f;
  cudaMemcpy(dg, &g, 4, cudaMemcpyHostToDevice);
  cudaMemset(dgx, 0xff, n * 4); cudaMemset(dgy, 0xff, n * 4);
  const long long sc = (long long)H * W, sn = sc * KC;
  int rc = iic::local_bwd_tcrb10h_try(dx_, sn, sc, W, dy_, sn, sc, W, B, KC, H, W, pad, dwx, dwy, dg, dgx, dgy, sn, sn, 0, 1.f, 0);
  cudaError_t err = cudaDeviceSynchronize();
  printf("try rc=%d (%s); first launch: %s\n", rc, iic::get_error(), cudaGetErrorString(err));
  if (rc != 0 || err != cudaSuccess) return 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 3;
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) iic::local_bwd_tcrb10h_try(dx_, sn, sc, W, dy_, sn, sc, W, B, KC, H, W, pad, dwx, dwy, dg, dgx, dgy, sn, sn, 0, 1.f, 0);
  cudaEventRecord(e1);
  err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("timed launches: %s\n", cudaGetErrorString(err)); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
#ifdef IIC_TC_TRACE
  {
    long long tr[4][64][6];
    cudaMemcpyFromSymbol(tr, iic::bwdrb10h::g_trace, sizeof(tr));
    const long long t0 = tr[0][0][0];
    printf("stage | producer: enter waited | issuer0: enter a_full_ok issued | issuer1: enter a_full_ok issued | transform: enter a_empty_ok raw_ok done\n");
    for (int k = 0; k < 28; ++k)
      printf("%3d | %7lld %7lld | %7lld %7lld %7lld | %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", k + 16, tr[0][k][0] - t0, tr[0][k][1] - t0,
             tr[1][k][0] - t0, tr[1][k][1] - t0, tr[1][k][2] - t0, tr[2][k][0] - t0, tr[2][k][1] - t0, tr[2][k][2] - t0,
             tr[3][k][0] - t0, tr[3][k][1] - t0, tr[3][k][2] - t0, tr[3][k][3] - t0);
    long long ct[32][8];
    cudaMemcpyFromSymbol(ct, iic::bwdrb10h::g_ctrace, sizeof(ct));
    printf("chunk | issuer: tmem_ready wait enter, ok | drain: accum_full wait enter, ok, drained, zeroed   (clk, relative to the first trace stamp)\n");
    for (int k = 0; k < 12; ++k)
      printf("%3d | %8lld %8lld | %8lld %8lld %8lld %8lld | in stores %lld clk, in TMEM waits %lld clk\n", k, ct[k][0] - t0, ct[k][1] - t0, ct[k][2] - t0, ct[k][3] - t0, ct[k][4] - t0, ct[k][5] - t0, ct[k][6], ct[k][7]);
  }
#endif
  std::vector<float> gx(n), gy(n);
  cudaMemcpy(gx.data(), dgx, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(gy.data(), dgy, n * 4, cudaMemcpyDeviceToHost);
  double ex = 0, ey = 0, mx = 0, my = 0;
  long long nan = 0;
  for (size_t i = 0; i < n; ++i) if (gx[i] != gx[i] || gy[i] != gy[i]) ++nan;
  // sample: all channels at a set of pixels including borders
  const int rows[] = {0, 1, H / 2, H - 2 > 0 ? H - 2 : 0, H - 1};
  const int cols[] = {0, 1, 2, 3, 4, 31 < W ? 31 : W - 1, W / 2, 127 < W ? 127 : W - 1, 128 < W ? 128 : W - 1, W - 2, W - 1};
  for (int b = 0; b < B; b += (B > 2 ? B - 1 : 1))
    for (int u : rows) for (int v : cols) for (int o = 0; o < KC; o += 3) {
      double rx = 0, ry = 0;
      for (int c = 0; c < KC; ++c) for (int ty = 0; ty < T; ++ty) for (int tx = 0; tx < T; ++tx) {
        const int uu = u + ty - pad, vv = v + tx - pad;
        if (uu < 0 || uu >= H || vv < 0 || vv >= W) continue;
        rx += (double)hwx[((size_t)c * T2 + ty * T + tx) * Kp4 + o] * hy[(((size_t)b * KC + c) * H + uu) * W + vv];
        ry += (double)hwy[((size_t)c * T2 + ty * T + tx) * Kp4 + o] * hx[(((size_t)b * KC + c) * H + uu) * W + vv];
      }
      rx *= g; ry *= g;
      const size_t idx = (((size_t)b * KC + o) * H + u) * W + v;
      ex = fmax(ex, fabs(gx[idx] - rx)); ey = fmax(ey, fabs(gy[idx] - ry));
      mx = fmax(mx, fabs(rx)); my = fmax(my, fabs(ry));
    }
  const double flop = 2.0 * 2 * T2 * KC * KC * (double)B * H * W;
  printf("K=%d pad=%d B=%d H=%d W=%d  %.3f ms (both gradients)  %.1f TFLOP/s (fp32-equivalent)  NaNs %lld\n", KC, pad, B, H, W, ms, flop / ms / 1e9, nan);
  printf("max-norm rel err gx %.3e gy %.3e (max ref %.3e %.3e)\n", ex / mx, ey / my, mx, my);
  printf((ex / mx < 2e-6 && ey / my < 2e-6 && nan == 0) ? "PASS\n" : "FAIL\n");
  return 0;
}

// Stand-alone harness for the packed tcgen05 local joint at K = 9 or 10, padding 1 (csrc/local_fwd_tcp10.cu):
// runs the kernel on random simplex-like maps, adds the per-CTA slots in fp64 and compares a sample of entries with
// an fp64 CPU loop; prints the time per launch and the fp32-equivalent TFLOP/s.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/tc_jointp10 tools/tc_jointp10_harness.cu
//   tools/_bin/tc_jointp [B H W K pad]
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../mi-based-regularized-semi-supervised-segmentation_b200/csrc/common.cuh"
namespace iic {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
const char* get_error() { return g_err; }
int current_device() { return 0; }
int sm_count_cached(int) { return 148; }
}
#include "local_fwd_tcp10.cu"

int main(int argc, char** argv) {
  int B = 1, H = 6, W = 32, KC = 10, pad = 1;
  if (argc >= 4) { B = atoi(argv[1]); H = atoi(argv[2]); W = atoi(argv[3]); }
  if (argc >= 6) { KC = atoi(argv[4]); pad = atoi(argv[5]); }
  const int T = 2 * pad + 1;
  const size_t n = (size_t)B * KC * H * W;
  std::vector<float> hx(n), hy(n);
  srand(1);
  for (size_t i = 0; i < n; ++i) {
    const float a = (float)rand() / RAND_MAX, b = (float)rand() / RAND_MAX;
    hx[i] = a * a * a * 0.3f;
    hy[i] = b * b * b * 0.3f;
  }
  float *dx_, *dy_, *dws;
  double* dJ;
  cudaMalloc(&dx_, n * 4); cudaMalloc(&dy_, n * 4);
  cudaMemcpy(dx_, hx.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dy_, hy.data(), n * 4, cudaMemcpyHostToDevice);
  const size_t E = (size_t)T * T * KC * KC;
  const size_t wsf = iic::local_joint_tcp10_slot_floats(KC, pad) * 148;
  cudaMalloc(&dws, wsf * 4);
  cudaMemset(dws, 0xff, wsf * 4);
  cudaMalloc(&dJ, E * 8);
  const long long sc = (long long)H * W, sn = sc * KC;
  int rc = iic::local_joint_tcp10_try(dx_, sn, sc, W, dy_, sn, sc, W, B, KC, H, W, pad, dws, wsf, dJ, 0);
  cudaError_t err = cudaDeviceSynchronize();
  printf("try rc=%d (%s); first launch: %s\n", rc, iic::get_error(), cudaGetErrorString(err));
  if (rc != 0 || err != cudaSuccess) return 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 5;
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) iic::local_joint_tcp10_try(dx_, sn, sc, W, dy_, sn, sc, W, B, KC, H, W, pad, dws, wsf, dJ, 0);
  cudaEventRecord(e1);
  err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("timed launches: %s\n", cudaGetErrorString(err)); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  std::vector<double> hJ(E);
  cudaMemcpy(hJ.data(), dJ, E * 8, cudaMemcpyDeviceToHost);
  double max_rel = 0.0, sum_rel = 0.0, sum_rel2 = 0.0; int cnt = 0;
  for (int ddy = 0; ddy < T; ++ddy) for (int ddx = 0; ddx < T; ++ddx)
    for (int i = 0; i < KC; i += 1)
      for (int j = 0; j < KC; j += 3) {
        double ref = 0.0;
        for (int b = 0; b < B; ++b)
          for (int u = 0; u < H; ++u) {
            const int xr = u + ddy - pad;
            if (xr < 0 || xr >= H) continue;
            const float* xrow = &hx[((size_t)(b * KC + i) * H + xr) * W];
            const float* yrow = &hy[((size_t)(b * KC + j) * H + u) * W];
            for (int v = 0; v < W; ++v) {
              const int xc = v + ddx - pad;
              if (xc < 0 || xc >= W) continue;
              ref += (double)xrow[xc] * (double)yrow[v];
            }
          }
        const double got = hJ[(((size_t)ddy * T + ddx) * KC + i) * KC + j];
        const double re = (got - ref) / ref;
        sum_rel += re; sum_rel2 += re * re; ++cnt;
        if (fabs(re) > max_rel) max_rel = fabs(re);
      }
  const double flop = 2.0 * T * T * KC * KC * (double)B * H * W;
  const double mean = sum_rel / cnt;
  printf("K=%d pad=%d B=%d H=%d W=%d  %.3f ms (joint + slot reduce)  %.1f TFLOP/s (fp32-equivalent)  max rel err %.3e\n", KC, pad, B, H, W, ms,
         flop / ms / 1e9, max_rel);
  printf("signed rel err: mean %.3e  std %.3e over %d entries\n", mean, sqrt(fmax(sum_rel2 / cnt - mean * mean, 0.0)), cnt);
  printf(max_rel < 2e-5 ? "PASS\n" : "FAIL\n");
  return 0;
}

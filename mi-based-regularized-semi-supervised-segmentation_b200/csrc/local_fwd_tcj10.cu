// Joint of the local IIC term on the tensor cores for 10 clusters, 3 x 3 window (BASELINE config 2):
//   J[dy][dx][i][j] = sum_{n,u,v} x[n,i,u+dy-1,v+dx-1] * y[n,j,u,v]                       (iic_loss.py:120-123)
// The FFMA2 kernel (local_fwd_fast3.inc) needs 900 FMAs per pixel and sits at 57 % of the FP32 pipe; the round-1
// tensor-core attempt lost because every x element was rewritten six times (three column shifts x fp32 + bf16 copies).
// Here every element is written ONCE:
//   * operands are fp16 hi/lo pairs (a = a1 + a2, 22 bits; all four part products are computed, fp32 accumulation),
//     staged PIXEL-major: one 64-byte group of 32 slots per pixel, [a1(0..9) a2(0..9) 0 x 12], 64-byte swizzled;
//   * the MMA consumes them MN-major (reduction = pixels, 16 per instruction): the M atoms of A are four consecutive x
//     rows (dy - 1 .. dy + 2, the fourth unused), the N atoms of B are the SAME y row at three pixel shifts -- an atom
//     stride of one pixel (64 bytes), i.e. overlapping atoms: tools/mn_major_micro.cu shows the hardware takes them and
//     applies the swizzle to the absolute address.  One MMA (M = 128, N = 96, K = 16 pixels) therefore adds 16 pixels of
//     one y row into all nine displacements: D[(dy, x slot), (shift, y slot)].
// Pipeline per CTA (one per SM, its share of the B*H image rows): TMA row loads (x and y interleaved) -> eight transform
// warps (thread = pixel: split, pack, four swizzled 16-byte stores; the simplex assertion of iic_loss.py:113 rides along)
// -> one issuing lane -> five TMEM accumulator sets taken round robin by y row (shortens every accumulation run: the
// tensor core accumulates in fp32 with truncation, local_fwd_tc.cu) -> at the end four warps drain the sets, add the parts
// and write this CTA's slot in the standard [d][i][j] layout that iic_finish reduces.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace iic {
namespace fwdtcj10 {
using namespace tc;

constexpr int PAD = 1;
constexpr int NPXB = 256;                // pixels per row buffer
constexpr int ROWB = NPXB * 64;          // 16 KB: [pixel][32 fp16 slots]
constexpr int NX = 5, NXM = 2;           // x row ring + mirror slots (slots 0, 1 are written twice so that four consecutive
                                         // slots starting anywhere in the ring are contiguous)
constexpr int NY = 3;                    // y row ring; NX = NY + 2 lets one wait cover both rings (see the transform loop)
constexpr int NRAW = 4;
constexpr int RAW_SLOT = 10 * 256 * 4;   // 10240: fp32 [K][W + 8]
constexpr int NSET = 5;                  // TMEM accumulator sets of 96 columns
constexpr int NCOL = 96;
constexpr int MAXW = 236;                // (W + 2) pixels of an x row in at most 15 k-steps of 16
constexpr int NTHREADS = 14 * 32;        // warps: 0 TMA, 1 issuer, 2-9 transform, 10-13 drain (13 also allocates TMEM)
constexpr int X_BYTES = (NX + NXM) * ROWB;
constexpr int SMEM_BYTES = X_BYTES + NY * ROWB + NRAW * RAW_SLOT + 1024;
constexpr float SCALE = 256.f;           // both maps are scaled by 2^8 before the split; 2^-16 in the drain

struct Params {
  int B, H, W, K;
  float* partial;      // [gridDim.x][9][K][K]
  int* flags;          // nullable: simplex assertion on x
#ifdef IIC_TCJ_DEBUG
  int dbg;             // harness only: 1 = no MMAs, 2 = no transform work, 4 = no TMA loads
#endif
};
#ifdef IIC_TCJ_DEBUG
#define TCJ_DBG(bit) (P.dbg & (bit))
__device__ long long g_tcj_trace[5][256];
#define TCJ_T(role, idx) do { if (blockIdx.x == 0 && lane == 0 && (idx) < 256) g_tcj_trace[role][idx] = clock64(); } while (0)
#else
#define TCJ_DBG(bit) 0
#define TCJ_T(role, idx) do { } while (0)
#endif

// MN-major, SWIZZLE_64B: LBO = byte stride between the 32-slot atoms along M / N, SBO = stride between groups of 8 pixels
__device__ __forceinline__ uint64_t make_desc_mn_sw64(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// the chunk of image rows that starts at global row r (rows of all images, B*H) inside the CTA share [r, R1)
struct Chunk { int n, h0, nr; };
__device__ __forceinline__ Chunk next_chunk(long long r, long long R1, int H) {
  Chunk c;
  c.n = (int)(r / H);
  c.h0 = (int)(r - (long long)c.n * H);
  c.nr = H - c.h0;
  if (c.nr > R1 - r) c.nr = (int)(R1 - r);
  return c;
}

__global__ void __launch_bounds__(NTHREADS, 1)
local_joint_tcj10_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapy, const Params P) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t raw_full[NRAW], raw_empty[NRAW], y_full[NY], y_done[NY], all_done;
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* x_ring = smem;
  unsigned char* y_ring = smem + X_BYTES;
  unsigned char* raw_ring = y_ring + NY * ROWB;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rows_total = (long long)P.B * P.H;
  const long long R0 = (long long)blockIdx.x * rows_total / gridDim.x;
  const long long R1 = (long long)(blockIdx.x + 1) * rows_total / gridDim.x;
  const int SW = P.W + 8;
  const int raw_bytes = P.K * SW * 4;
  const int ksteps = (P.W + 2 + 15) / 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NRAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 8); }
    for (int s = 0; s < NY; ++s) { mbar_init(&y_full[s], 8); mbar_init(&y_done[s], 1); }
    mbar_init(&all_done, 1);
    mbar_fence_init();
  }
  if (wid == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the row buffers start as zeros: the columns outside the map (and slots 20..31) are never written again
  for (int e = threadIdx.x; e < (X_BYTES + NY * ROWB) / 16; e += NTHREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (wid == 0) {
    // ===== TMA producer: raw rows in job order  X(h0-1) X(h0) X(h0+1) Y(h0) X(h0+2) Y(h0+1) ... X(h0+nr) Y(h0+nr-1) =====
    if (lane == 0) {
      tma_prefetch_desc(&mapx);
      tma_prefetch_desc(&mapy);
      int t = 0, s = 0;
      unsigned sph = 0;
      for (long long r = R0; r < R1;) {
        const Chunk c = next_chunk(r, R1, P.H);
        const int njobs = 2 * c.nr + 2;
        for (int j = 0; j < njobs; ++j, ++t) {
          // job j: j < 3 -> X local row j; else alternating Y(ly), X(lr): j = 3 + 2*ly -> Y(ly); j = 4 + 2*ly -> X(ly + 3)
          const bool is_y = j >= 3 && ((j - 3) & 1) == 0;
          const int lrow = j < 3 ? j : (is_y ? (j - 3) / 2 : (j - 4) / 2 + 3);
          if (t >= NRAW) mbar_wait(&raw_empty[s], sph ^ 1u, 1);
          TCJ_T(0, t);
          if (TCJ_DBG(4)) { mbar_arrive(&raw_full[s]); if (++s == NRAW) { s = 0; sph ^= 1u; } continue; }
          mbar_arrive_expect_tx(&raw_full[s], raw_bytes);
          if (is_y) tma_load_4d(raw_ring + s * RAW_SLOT, &mapy, &raw_full[s], -4, c.h0 + lrow, 0, c.n);
          else tma_load_4d(raw_ring + s * RAW_SLOT, &mapx, &raw_full[s], -4, c.h0 - PAD + lrow, 0, c.n);
          if (++s == NRAW) { s = 0; sph ^= 1u; }
        }
        r += c.nr;
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer: per y row `ksteps` MMAs (M = 128: four x rows; N = 96: the y row at three shifts; K = 16 pixels) =====
    // kind::f16, fp16 operands, fp32 accumulator, A and B MN-major (bits 15, 16), N = 96, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NCOL >> 3) << 17) | (8u << 24);
    int yg = 0, xbase = 0;                 // running counts of y rows and of x rows at the chunk start
    for (long long r = R0; r < R1;) {
      const Chunk c = next_chunk(r, R1, P.H);
      for (int ly = 0; ly < c.nr; ++ly, ++yg) {
        const int ys = yg % NY;
        mbar_wait(&y_full[ys], (unsigned)(yg / NY) & 1u, 5);       // y row ly and x rows ly .. ly + 2 are in place
        asm volatile("tcgen05.fence::after_thread_sync;");
        TCJ_T(1, yg);
        if (lane == 0) {
          const int xs = (xbase + ly) % NX;                         // slot of x local row ly (= image row u - 1)
          const uint64_t a0 = make_desc_mn_sw64(smem_u32(x_ring + xs * ROWB), ROWB);
          const uint64_t b0 = make_desc_mn_sw64(smem_u32(y_ring + ys * ROWB), 64);
          const uint32_t d_tmem = tmem_base + (uint32_t)((yg % NSET) * NCOL);
          for (int k = 0; k < (TCJ_DBG(1) ? 0 : ksteps); ++k)
            umma_bf16(d_tmem, a0 + (uint64_t)(k * 64), b0 + (uint64_t)(k * 64), idesc, (yg >= NSET || k > 0) ? 1u : 0u);   // 16 pixels = 1024 bytes = 64 address units
          umma_commit(&y_done[ys]);
        }
        __syncwarp();
      }
      xbase += c.nr + 2;
      r += c.nr;
    }
    if (lane == 0) umma_commit(&all_done);
    __syncwarp();
  } else if (wid >= 2 && wid < 10) {
    // ===== transform: all eight warps take every job together, thread = pixel column =====
    const int tid = threadIdx.x - 64;           // 0..255
    int t = 0, s = 0, yg = 0, xg = 0;           // job counter, raw slot, running y / x row counts
    unsigned sph = 0;
    bool bad = false;
    for (long long r = R0; r < R1;) {
      const Chunk c = next_chunk(r, R1, P.H);
      // a new chunk reuses x slots whose rows the previous chunk's last y rows may still be reading: wait for all of them
      if (yg > 0)
        for (int q = yg > NY ? yg - NY : 0; q < yg; ++q) mbar_wait(&y_done[q % NY], (unsigned)(q / NY) & 1u, 6);
      const int njobs = 2 * c.nr + 2;
      const int ybase = yg;
      for (int j = 0; j < njobs; ++j, ++t) {
        const bool is_y = j >= 3 && ((j - 3) & 1) == 0;
        const int lrow = j < 3 ? j : (is_y ? (j - 3) / 2 : (j - 4) / 2 + 3);
        unsigned char* dst;
        unsigned char* dst2 = nullptr;
        int boff;
        if (is_y) {
          // slot reuse: y row (yg - NY) is done -- waited for by the X job just before (below), or at the chunk start
          dst = y_ring + (yg % NY) * ROWB;
          boff = 2;                            // buffer pixel = column + 2
        } else {
          // x local row lrow takes the slot of the x row NX before it, last read by y local row lrow - NX = (lrow - 2) - NY:
          // the same y row whose slot the NEXT job, Y(lrow - 2), reuses (NX = NY + 2): one wait serves both
          if (lrow >= NX) {
            const int q = ybase + lrow - NX;
            mbar_wait(&y_done[q % NY], (unsigned)(q / NY) & 1u, 8);
          }
          const int xs = xg % NX;
          dst = x_ring + xs * ROWB;
          if (xs < NXM) dst2 = x_ring + (xs + NX) * ROWB;
          boff = 1;                            // buffer pixel = column + 1
        }
        if (wid == 2) TCJ_T(2, t);
        mbar_wait(&raw_full[s], sph, 4);
        if (wid == 2) TCJ_T(3, t);
        const float* raw = reinterpret_cast<const float*>(raw_ring + s * RAW_SLOT);
        const int col = tid;
        if (col < P.W && !TCJ_DBG(2)) {
          float v[10];
#pragma unroll
          for (int ch = 0; ch < 10; ++ch) v[ch] = ch < P.K ? raw[ch * SW + col + 4] : 0.f;
          if (!is_y && P.flags) {
            // simplex(x_out) of iic_loss.py:113 on the rows of the map (the halo rows outside it are zero fill)
            const int img_row = c.h0 - PAD + lrow;
            if (img_row >= 0 && img_row < P.H) {
              float sum = 0.f;
#pragma unroll
              for (int ch = 0; ch < 10; ++ch) sum += v[ch];
              if (!(fabsf(sum - 1.f) <= 2e-4f)) bad = true;
            }
          }
          uint32_t hp[5], lp[5];
#pragma unroll
          for (int c2 = 0; c2 < 5; ++c2) {
            const float s0 = v[2 * c2] * SCALE, s1 = v[2 * c2 + 1] * SCALE;
            const __half2 hh = __floats2half2_rn(s0, s1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
            hp[c2] = *reinterpret_cast<const uint32_t*>(&hh);
            lp[c2] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          // slots [h0..h9 | l0..l9 | 0 x 12] = chunks (h0-7) (h8 h9 l0-5) (l6-9 0 0 0 0) (0); chunk index XOR (pixel >> 1) & 3
          const int b = col + boff;
          const int sw = (b >> 1) & 3;
          const uint4 c0 = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          const uint4 c1 = make_uint4(hp[4], lp[0], lp[1], lp[2]);
          const uint4 c2v = make_uint4(lp[3], lp[4], 0u, 0u);
          unsigned char* px = dst + b * 64;
          *reinterpret_cast<uint4*>(px + ((0 ^ sw) << 4)) = c0;
          *reinterpret_cast<uint4*>(px + ((1 ^ sw) << 4)) = c1;
          *reinterpret_cast<uint4*>(px + ((2 ^ sw) << 4)) = c2v;
          if (dst2) {
            unsigned char* px2 = dst2 + b * 64;
            *reinterpret_cast<uint4*>(px2 + ((0 ^ sw) << 4)) = c0;
            *reinterpret_cast<uint4*>(px2 + ((1 ^ sw) << 4)) = c1;
            *reinterpret_cast<uint4*>(px2 + ((2 ^ sw) << 4)) = c2v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[s]);
        if (is_y) {
          // every x row this y row pairs with was written by earlier jobs of the same eight warps: publish the row
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&y_full[yg % NY]);
          ++yg;
        } else {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          ++xg;
        }
        if (wid == 2) TCJ_T(4, t);
        if (++s == NRAW) { s = 0; sph ^= 1u; }
      }
      r += c.nr;
    }
    if (P.flags && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(P.flags, IIC_FLAG_NOT_SIMPLEX);
  } else if (wid >= 10) {
    // ===== drain: after the last MMA add the sets and the (part x part) products, write this CTA's slot =====
    const int q4 = wid & 3;                      // TMEM lane quarter = x row atom: dy = q4 (the fourth row is not used)
    mbar_wait(&all_done, 0, 9);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (q4 < 3) {
      const long long nrows = R1 - R0;
      const int nsets = nrows < NSET ? (int)nrows : NSET;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
      float acc[NCOL];
#pragma unroll
      for (int c = 0; c < NCOL; ++c) acc[c] = 0.f;
      for (int st = 0; st < nsets; ++st) {
#pragma unroll
        for (int c = 0; c < NCOL; c += 16) {
          uint32_t v[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(lane_base + (uint32_t)(st * NCOL + c)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 16; ++q) acc[c + q] += __uint_as_float(v[q]);
        }
      }
      // lane = x slot (i: hi part, 10 + i: lo part); column = 32 * shift + y slot (j: hi, 10 + j: lo); dx = 2 - shift
      float* slot = P.partial + (size_t)blockIdx.x * (9 * P.K * P.K);
#pragma unroll
      for (int sft = 0; sft < 3; ++sft)
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const float v = acc[sft * 32 + j] + acc[sft * 32 + 10 + j];
          const float tot = v + __shfl_down_sync(0xffffffffu, v, 10);
          if (lane < P.K && j < P.K)
            slot[((size_t)(q4 * 3 + (2 - sft)) * P.K + lane) * P.K + j] = tot * (1.f / (SCALE * SCALE));
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (wid == 13) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace fwdtcj10

// 0 = launched, 1 = error, -1 = not eligible.  *ncta = number of partial slots written ([9][K][K] floats each).
// *checked = 1 when `flags` was given and the simplex assertion on x ran inside the kernel.
int local_joint_tcj10_try(const float* x, long long x_sn, long long x_sc, long long x_sh, const float* y, long long y_sn,
                          long long y_sc, long long y_sh, int B, int K, int H, int W, int pad, float* partial, int max_ctas,
                          int* ncta, int* flags, int* checked, cudaStream_t st) {
  using namespace fwdtcj10;
  *checked = 0;
  if (pad != PAD || K < 2 || K > 10 || W % 4 != 0 || W < 8 || W > MAXW) return -1;
  const long long rows = (long long)B * H;
  if (rows < 4LL * max_ctas) return -1;            // small maps: the pipeline fill dominates, the FFMA2 kernel is faster
  CUtensorMap mx, my;
  if (!make_map_4d(&mx, x, B, K, H, W, x_sn, x_sc, x_sh, W + 8, 1, K)) return -1;
  if (!make_map_4d(&my, y, B, K, H, W, y_sn, y_sc, y_sh, W + 8, 1, K)) return -1;
  Params P;
  P.B = B; P.H = H; P.W = W; P.K = K;
  P.partial = partial;
  P.flags = flags;
  *checked = flags != nullptr;
  *ncta = max_ctas;
  IIC_CHECK_RC(ensure_dyn_smem((const void*)local_joint_tcj10_kernel, SMEM_BYTES));
  local_joint_tcj10_kernel<<<max_ctas, NTHREADS, SMEM_BYTES, st>>>(mx, my, P);
  IIC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace iic

// Local IIC joint on the tensor cores for 10 clusters, 3 x 3 window (BASELINE config 2).  Packed scheme of
// local_fwd_tcp.cu with one more level of packing, so that a 128-row MMA is full at K = 10: a k-block is a 16-pixel
// column segment of FOUR consecutive x rows q0 .. q0+3,
//   A rows  = (x row xr, channel i, column shift dx)     4 x 10 x 3 = 120 rows
//   B rows  = (channel j, y row yr), the SIX y rows q0-1 .. q0+4                  N = 64 (60 used)
// and D[(xr,i,dx), (j,yr)] holds every (x row, y row) pair of the block; the pairs with yr - xr in {0,1,2} are the
// three row displacements dy = xr - yr + 2 (the other half of D is unused: an MMA costs the same for any N <= 128).
// 4 MMAs (2 products) per 64 pixels.  Reference arithmetic: contrastyou/losses/iic_loss.py:120-123.
//
// What the clock64() traces of the other tensor-core kernels showed is built in: every mbarrier poll costs a warp
// ~300 clk and a lane issues an MMA only every ~88 clk although the pipe takes one every ~62 clk, so there are two
// transform groups (6 warps each) and two issuing warps that take alternate k-blocks, the accumulator is zeroed by
// the drain (every MMA accumulates: issue order is free) and double buffered (the drain of one accumulation run
// overlaps the MMAs of the next).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../mi-based-regularized-semi-supervised-segmentation_b200/csrc/common.cuh"
#include "../../mi-based-regularized-semi-supervised-segmentation_b200/csrc/tma.cuh"

namespace iic {
namespace fwdtcp10 {

constexpr int T = 3, PAD = 1;
constexpr int XR = 4, YR = XR + 2;             // x rows per k-block, y rows that pair with them
constexpr int PXB = 16, XRW = 28;              // pixels per k-block; staged x columns c0-4 .. c0+23
constexpr int NB = 64;                         // B rows (K * 6 <= 60 used)
constexpr int ATILE = 128 * PXB * 4;           // 8192
constexpr int BTILE = NB * PXB * 4;            // 4096
constexpr int OP_BYTES = 2 * ATILE + 2 * BTILE;   // 24576: A fp32, A bf16, B fp32, B bf16
constexpr int NOP = 4;
constexpr int XRAW_MAX = 4608;                 // K * XR * 112 bytes <= 4480, padded so that the y box is 512-byte aligned
constexpr int RAW_BYTES = 9216;                // + K * YR * 64 <= 3840
constexpr int NRAW = 6;
constexpr int NTHREADS = 512;                  // warps: 0 3 TMA (3 also TMEM), 1 2 issuers, 4-15 transform (two groups), 4-7 also drain
constexpr int SEG_KB = 32;                     // k-blocks per accumulation run (4 rows x 512 pixels)
constexpr int SLOT_FLOATS = 128 * NB;
constexpr int SMEM_BYTES = NOP * OP_BYTES + NRAW * RAW_BYTES + 1024;

__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&p);
}
template <bool LO>
__device__ __forceinline__ uint4 pack8(const float* v) {
  float t[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) t[q] = LO ? tf32_lo(v[q]) : v[q];
  return make_uint4(pack_bf16(t[0], t[1]), pack_bf16(t[2], t[3]), pack_bf16(t[4], t[5]), pack_bf16(t[6], t[7]));
}


struct Params {
  int B, H, W, K, segs_w, groups;   // segs_w = W / 16, groups = ceil(H / 4)
  float* partial;                   // [gridDim.x][SLOT_FLOATS]
};

// slot element of accumulator row m, column c: chunks of 8 columns, float4-interleaved over rows (coalesced drains)
__host__ __device__ inline size_t slot_index(int m, int c) {
  return ((((size_t)(c / 8)) * 2 + (c % 8) / 4) * 128 + m) * 4 + (c % 4);
}

__global__ void __launch_bounds__(NTHREADS, 1)
local_joint_tcp10_kernel(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapy, const Params P) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t raw_full[NRAW], raw_empty[NRAW], op_full[NOP], op_empty[NOP], accum_bar[2], zeroed_bar[2];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* raw_ring = smem + NOP * OP_BYTES;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_total = P.B * P.groups * P.segs_w;
  const int kb0 = (int)((long long)blockIdx.x * nkb_total / gridDim.x);
  const int kb1 = (int)((long long)(blockIdx.x + 1) * nkb_total / gridDim.x);
  const int nkb = kb1 - kb0;
  const int nrun = (nkb + SEG_KB - 1) / SEG_KB;
  const int K = P.K;
  const int xraw_bytes = K * XR * XRW * 4, yraw_bytes = K * YR * 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NRAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 6); }
    for (int s = 0; s < NOP; ++s) { mbar_init(&op_full[s], 6); mbar_init(&op_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&accum_bar[s], 2); mbar_init(&zeroed_bar[s], 4); }
    mbar_fence_init();
  }
  if (wid == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (wid == 0 || wid == 3) {
    // ===== TMA producers: two lanes (warp 0, warp 3) take alternate k-blocks -- one lane's iteration (barrier poll,
    // expect_tx, two tensor copies) is as long as everything else in a k-block =====
    if (lane == 0) {
      const int par = wid == 0 ? 0 : 1;
      tma_prefetch_desc(&mapx);
      tma_prefetch_desc(&mapy);
      const int per_img = P.groups * P.segs_w;
      for (int k = par; k < nkb; k += 2) {
        const int s = k % NRAW;
        if (k >= NRAW) mbar_wait(&raw_empty[s], ((unsigned)(k / NRAW) & 1u) ^ 1u, 1);
        const int kb = kb0 + k;
        const int nc = kb / per_img;
        const int rem = kb - nc * per_img;
        const int g = rem / P.segs_w, sg = rem - g * P.segs_w;
        const int c0 = sg * PXB, q0 = g * XR;
        unsigned char* st = raw_ring + (size_t)s * RAW_BYTES;
        mbar_arrive_expect_tx(&raw_full[s], xraw_bytes + yraw_bytes);
        tma_load_4d(st, &mapx, &raw_full[s], c0 - 4, q0, 0, nc);
        tma_load_4d(st + XRAW_MAX, &mapy, &raw_full[s], c0, q0 - 1, 0, nc);
      }
    }
  } else if (wid == 1 || wid == 2) {
    // ===== MMA issuers: alternate k-blocks =====
    const int par = wid - 1;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | (8u << 24);
    const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | (8u << 24);
    int o = 0, kin = 0, run = 0;
    unsigned oph = 0;
    for (int k = 0; k < nkb; ++k) {
      const int runlen = (nkb - run * SEG_KB) < SEG_KB ? (nkb - run * SEG_KB) : SEG_KB;
      const int buf = run & 1;
      if (kin == 0 || kin == 1) {
        if (kin == par || (kin == 0 && runlen == 1)) {
          // first k-block of this run for this issuer: the accumulator buffer must have been zeroed
          mbar_wait(&zeroed_bar[buf], (unsigned)(run >> 1) & 1u, 6);
          asm volatile("tcgen05.fence::after_thread_sync;");
        }
      }
      const bool mine = (kin & 1) == par;
      if (mine) {
        mbar_wait(&op_full[o], oph, 5);
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (lane == 0) {
          const uint64_t base = make_desc_sw64(smem_u32(smem + (size_t)o * OP_BYTES));
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NB);
#pragma unroll
          for (int ks = 0; ks < PXB / 8; ++ks) {
            umma_bf16(d_tmem, base + (uint64_t)(ATILE / 16 + ks * 2), base + (uint64_t)((2 * ATILE + BTILE) / 16 + ks * 2), idesc_bf16, 1u);
            umma_tf32(d_tmem, base + (uint64_t)(ks * 2), base + (uint64_t)(2 * ATILE / 16 + ks * 2), idesc, 1u);
          }
          umma_commit(&op_empty[o]);
        }
        __syncwarp();
      }
      // this issuer's last k-block of the run (or none at all in a one-k-block run): tell the drain
      const bool last_mine = mine ? (kin + 2 >= runlen) : (runlen == 1);
      if (last_mine && lane == 0) umma_commit(&accum_bar[buf]);
      __syncwarp();
      if (++o == NOP) { o = 0; oph ^= 1u; }
      if (++kin == runlen) { kin = 0; ++run; }
    }
  } else if (wid >= 4) {
    // ===== transform (two groups of six warps, alternate k-blocks); warps 4-7 also drain =====
    const int grp = wid >= 10 ? 1 : 0;
    const int tid = threadIdx.x - 128 - grp * 192;             // 0 .. 191 inside the group
    const int q4 = wid & 3;                                    // drains (warps 4-7): TMEM lanes 32*q4 ..
    const int m = q4 * 32 + lane;
    float* slot = P.partial + (size_t)blockIdx.x * SLOT_FLOATS;
    const int nrows_a = XR * K * T;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    if (wid < 8) {                                             // zero both accumulator buffers
      for (int c = 0; c < 2 * NB; c += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c), "r"(0u) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&zeroed_bar[0]); mbar_arrive(&zeroed_bar[1]); }
    }
    int o = 0, s = 0, kin = 0, run = 0;
    unsigned oph = 0, sph = 0;
    for (int k = 0; k < nkb; ++k) {
      const int runlen = (nkb - run * SEG_KB) < SEG_KB ? (nkb - run * SEG_KB) : SEG_KB;
      if ((k & 1) == grp) {
        mbar_wait(&raw_full[s], sph, 3);
        if (k >= NOP) mbar_wait(&op_empty[o], oph ^ 1u, 2);
        const unsigned char* raw = raw_ring + (size_t)s * RAW_BYTES;
        unsigned char* op = smem + (size_t)o * OP_BYTES;
        if (tid < nrows_a) {
          // A row tid = ((xr * K + i) * 3 + dx): operand pixel p is x column c0 + p + dx - 1 = staged column p + dx + 3
          const int dx = tid % T, xi = tid / T;
          const int i = xi % K, xr = xi / K;
          const int sw = (tid >> 1) & 3;
          const float* xs = reinterpret_cast<const float*>(raw) + (i * XR + xr) * XRW + dx + 3;
          float xv[PXB];
#pragma unroll
          for (int p = 0; p < PXB; ++p) xv[p] = xs[p];
          unsigned char* a32 = op + tid * 64;
          unsigned char* a16 = a32 + ATILE;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<float4*>(a32 + ((c ^ sw) << 4)) = make_float4(xv[4 * c], xv[4 * c + 1], xv[4 * c + 2], xv[4 * c + 3]);
          *reinterpret_cast<uint4*>(a16 + ((0 ^ sw) << 4)) = pack8<true>(xv);
          *reinterpret_cast<uint4*>(a16 + ((1 ^ sw) << 4)) = pack8<false>(xv);
          *reinterpret_cast<uint4*>(a16 + ((2 ^ sw) << 4)) = pack8<true>(xv + 8);
          *reinterpret_cast<uint4*>(a16 + ((3 ^ sw) << 4)) = pack8<false>(xv + 8);
        } else if (tid >= 128 && tid - 128 < K * YR) {
          // B row (channel j, y row yr) exactly as the swizzled TMA box laid it out: copy + bf16 [yh | yl]
          const int nrow = tid - 128;
          const int sw = (nrow >> 1) & 3;
          float yv[PXB];
          unsigned char* b32 = op + 2 * ATILE + nrow * 64;
          unsigned char* b16 = b32 + BTILE;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int off = (c ^ sw) << 4;
            const float4 t = *reinterpret_cast<const float4*>(raw + XRAW_MAX + nrow * 64 + off);
            *reinterpret_cast<float4*>(b32 + off) = t;
            yv[4 * c] = t.x; yv[4 * c + 1] = t.y; yv[4 * c + 2] = t.z; yv[4 * c + 3] = t.w;
          }
          *reinterpret_cast<uint4*>(b16 + ((0 ^ sw) << 4)) = pack8<false>(yv);
          *reinterpret_cast<uint4*>(b16 + ((1 ^ sw) << 4)) = pack8<true>(yv);
          *reinterpret_cast<uint4*>(b16 + ((2 ^ sw) << 4)) = pack8<false>(yv + 8);
          *reinterpret_cast<uint4*>(b16 + ((3 ^ sw) << 4)) = pack8<true>(yv + 8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&op_full[o]);
          mbar_arrive(&raw_empty[s]);
        }
      }
      if (wid < 8 && kin == runlen - 1) {
        // ---- drain this run's accumulator buffer into the slot (first run stores, later runs add), then zero it ----
        const int buf = run & 1;
        mbar_wait(&accum_bar[buf], (unsigned)(run >> 1) & 1u, 4);
        asm volatile("tcgen05.fence::after_thread_sync;");
        float4* sl4 = reinterpret_cast<float4*>(slot) + m;
        for (int c8 = 0; c8 < NB / 8; c8 += 2) {
          uint32_t a[16];
          const uint32_t taddr = lane_base + (uint32_t)(buf * NB + c8 * 8);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                         "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
                       : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float4* dst = sl4 + (size_t)c8 * 256;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (run > 0) g = dst[j * 128];
            dst[j * 128] = make_float4(__uint_as_float(a[4 * j]) + g.x, __uint_as_float(a[4 * j + 1]) + g.y,
                                       __uint_as_float(a[4 * j + 2]) + g.z, __uint_as_float(a[4 * j + 3]) + g.w);
          }
        }
        for (int c = 0; c < NB; c += 8)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + buf * NB + c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0 && run + 2 < nrun) mbar_arrive(&zeroed_bar[buf]);
      }
      if (++o == NOP) { o = 0; oph ^= 1u; }
      if (++s == NRAW) { s = 0; sph ^= 1u; }
      if (++kin == runlen) { kin = 0; ++run; }
    }
    if (nkb == 0 && wid < 8)
      for (int e = threadIdx.x - 128; e < SLOT_FLOATS; e += 128) slot[e] = 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (wid == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

// J[dy][dx][i][j] = sum over CTAs and the four x rows of a block of their slot element, fp64, fixed order
__global__ void __launch_bounds__(1024)
reduce_packed10_kernel(const float* __restrict__ partial, int ncta, int K, double* __restrict__ J) {
  __shared__ double sm[32][33];
  const int le = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + le;
  const int E = T * T * K * K;
  double s = 0.0;
  if (e < E) {
    const int j = e % K, i = (e / K) % K, dx = (e / (K * K)) % T, dy = e / (K * K * T);
    for (int c = g; c < ncta; c += 32) {
      const float* src = partial + (size_t)c * SLOT_FLOATS;
#pragma unroll
      for (int xr = 0; xr < XR; ++xr)
        s += (double)__ldg(src + slot_index((xr * K + i) * T + dx, j * YR + (xr + 2 - dy)));
    }
  }
  sm[g][le] = s;
  __syncthreads();
  if (g == 0 && e < E) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += sm[q][le];
    J[e] = t;
  }
}

static bool make_map(CUtensorMap* map, const float* base, int B, int K, int H, int W, long long sn, long long sc, long long sh,
                     int box_w, int box_h, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  if ((sh * 4) % 16 != 0 || (sc * 4) % 16 != 0 || (sn * 4) % 16 != 0) return false;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)K, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sh * 4, (cuuint64_t)sc * 4, (cuuint64_t)sn * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)K, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fwdtcp10

size_t local_joint_tcp10_slot_floats(int K, int pad) {
  return ((K == 9 || K == 10) && pad == 1) ? (size_t)fwdtcp10::SLOT_FLOATS : 0;
}

// Returns 0 when launched (J_out written), < 0 when the shape is not covered, > 0 on error.
int local_joint_tcp10_try(const float* x, long long x_sn, long long x_sc, long long x_sh, const float* y, long long y_sn,
                          long long y_sc, long long y_sh, int B, int K, int H, int W, int pad, float* partial,
                          size_t partial_floats, double* J_out, cudaStream_t st) {
  using namespace fwdtcp10;
  if (local_joint_tcp10_slot_floats(K, pad) == 0 || W % PXB != 0) return -1;
  CUtensorMap mx, my;
  if (!make_map(&mx, x, B, K, H, W, x_sn, x_sc, x_sh, XRW, XR, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
  if (!make_map(&my, y, B, K, H, W, y_sn, y_sc, y_sh, PXB, YR, CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
  const int sms = sm_count_cached(current_device());
  if (sms <= 0) return -1;
  const int groups = (H + XR - 1) / XR;
  const long long nkb = (long long)B * groups * (W / PXB);
  if (nkb < 8LL * sms && !getenv("IIC_B200_TC10_FORCE")) return -1;          // small maps: the FFMA2 kernel is faster
  long long grid = sms;
  if ((long long)(partial_floats / SLOT_FLOATS) < grid) grid = (long long)(partial_floats / SLOT_FLOATS);
  if (grid > nkb) grid = nkb;
  if (grid < 1) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    IIC_CHECK_CUDA(cudaFuncSetAttribute(local_joint_tcp10_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Params P{B, H, W, K, W / PXB, groups, partial};
  local_joint_tcp10_kernel<<<(int)grid, NTHREADS, SMEM_BYTES, st>>>(mx, my, P);
  IIC_CHECK_CUDA(cudaGetLastError());
  const int E = T * T * K * K;
  reduce_packed10_kernel<<<(E + 31) / 32, 1024, 0, st>>>(partial, (int)grid, K, J_out);
  IIC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace iic

// Shared device helpers of the tensor-core (tcgen05 / UMMA) kernels: shared-memory operand descriptors, MMA issue and
// commit, bulk copies, and the fp32 -> (tf32 hi, bf16 correction) split of the two-MMA product
//   a*w = ah*wh + (al*wh + ah*wl) + al*wl,   ah = top 19 bits of a (what kind::tf32 reads), al = a - ah (exact),
// main term on kind::tf32, the two correction terms on bf16 copies in one kind::f16 MMA, al*wl dropped
// (tests/test_split_product.py bounds the error).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "tma.cuh"

namespace iic {
namespace tc {

// Bring-up knobs (IIC_TC_DBG ablates loads / transforms / stores and therefore produces WRONG results, IIC_TC_SEG and
// IIC_TC_NOP change accumulation-run lengths): only a build with -DIIC_TC_BRINGUP (the harnesses under tools/) reads
// them; in the product library they are compile-time defaults.
#ifdef IIC_TC_BRINGUP
inline int bringup_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
#else
inline int bringup_env(const char*, int dflt) { return dflt; }
#endif

// K-major, SWIZZLE_64B: 64-byte rows, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
  return d;
}
// K-major, no swizzle: ((8,n),2):((16 B, SBO), LBO) with SBO = 128, i.e. consecutive 8-row core matrices are contiguous
// and row r of a chunk sits at r * 16 bytes; lbo_bytes = distance between the two 16-byte chunks of a row
__device__ __forceinline__ uint64_t make_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
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
// always-accumulate forms (kernels whose accumulators are zeroed by the drain)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// the mbarrier gets one arrival when every MMA this thread issued before has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {     // a at the lower address
  const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&p);
}
// eight fp32 values -> 16 bytes of bf16; LO selects the tf32 remainder instead of the value
template <bool LO>
__device__ __forceinline__ uint4 pack8(const float* v) {
  float t[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) t[q] = LO ? tf32_lo(v[q]) : v[q];
  return make_uint4(pack_bf16(t[0], t[1]), pack_bf16(t[2], t[3]), pack_bf16(t[4], t[5]), pack_bf16(t[6], t[7]));
}

}  // namespace tc
}  // namespace iic

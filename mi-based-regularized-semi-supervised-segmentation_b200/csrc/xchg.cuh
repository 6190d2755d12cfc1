// Layout and primitives of the NVLink peer-memory exchange buffers (xchg.cu, finish.cu).
//   [header: flags[2][MAXR] u64, seq u64, arrive u32, done u32][pad to 1 KB][data: 2 parities x world x capacity doubles]
#pragma once
#include "common.cuh"

namespace iic {

constexpr int XCHG_MAXR = 16;
constexpr size_t XCHG_HDR_BYTES = 1024;

struct XchgHeader {
  unsigned long long flags[2][XCHG_MAXR];
  unsigned long long seq;
  unsigned int arrive;
  unsigned int done;
};

struct XchgPeers {
  unsigned char* base[XCHG_MAXR];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// slot of `rank` in the buffer at `base` for the given parity
__device__ __forceinline__ double* xchg_slot(unsigned char* base, int par, int world, int rank, long long capacity) {
  return reinterpret_cast<double*>(base + XCHG_HDR_BYTES) + ((size_t)par * world + rank) * capacity;
}
// Wait until rank `src` has published `seq` in this rank's header; false on time-out (a peer that never arrives).
__device__ __forceinline__ bool xchg_wait_flag(const XchgHeader* hdr, int par, int src, unsigned long long seq,
                                               unsigned long long timeout_ns) {
  const unsigned long long t0 = global_timer_ns();
  unsigned int spins = 0;
  while (ld_acquire_sys(&hdr->flags[par][src]) < seq) {
    __nanosleep(64);
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) return false;
  }
  return true;
}

}  // namespace iic

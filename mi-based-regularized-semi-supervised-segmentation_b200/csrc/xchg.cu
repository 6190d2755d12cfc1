// Joint exchange over NVLink peer memory: the one collective of the path.  Every rank holds the
// fp64 partial joints of its batch shard (a few KB to 1.2 MB); all ranks need their sum before the epilogue.
// Instead of an NCCL all-reduce (latency-bound at this size, ~15-20 us per call inside the step) each rank
//   1. stores its partial joints into its own slot of EVERY peer's exchange buffer (plain st.global on peer pointers
//      mapped with CUDA IPC; NVSwitch gives every peer full bandwidth),
//   2. publishes a sequence number in each peer's flag word for (parity, source rank) after a system-scope fence,
//   3. waits until its own flag words show the current sequence number from every rank, and
//   4. adds the world_size slots in rank order in fp64 -- the same order on every rank, so all ranks hold bit-identical
//      joints (and therefore bit-identical losses and epilogue coefficients).
// Two parities alternate, so a rank that races ahead into the next exchange never overwrites slots a slow peer is
// still reading: it cannot start exchange s+2 before every peer has published s+1, i.e. finished reading s.
// The sequence counter lives in device memory and is advanced by the kernel itself, so the exchange can sit inside a
// CUDA graph and be replayed.  One process per GPU; ranks must issue the same exchanges in the same order.
//
// Buffer layout (per rank, cudaMalloc'd by iic_xchg_create so that the IPC handle maps the allocation from offset 0):
//   [header: flags[2][MAXR] u64, seq u64, arrive u32, done u32][pad to 1 KB][data: 2 parities x world x capacity doubles]
#include "xchg.cuh"

namespace iic {

__global__ void __launch_bounds__(256)
xchg_allreduce_kernel(double* __restrict__ J, long long E, long long capacity, XchgPeers peers, int rank, int world,
                      unsigned long long timeout_ns, int* __restrict__ flags) {
  XchgHeader* hdr = reinterpret_cast<XchgHeader*>(peers.base[rank]);
  __shared__ unsigned long long seq_s;
  __shared__ int last_s;
  if (threadIdx.x == 0) seq_s = *reinterpret_cast<volatile unsigned long long*>(&hdr->seq) + 1;   // advanced by the last CTA below
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int par = (int)(seq & 1ull);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long e0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;

  // 1. my partial joints into my slot of every rank's buffer (my own included)
  for (long long e = e0; e < E; e += stride) {
    const double v = J[e];
    for (int p = 0; p < world; ++p) {
      double* dst = reinterpret_cast<double*>(peers.base[p] + XCHG_HDR_BYTES) + ((size_t)par * world + rank) * capacity;
      dst[e] = v;
    }
  }
  // 2. the last CTA to get here publishes the sequence number on every rank
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last_s = (atomicAdd(&hdr->arrive, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (last_s) {
    __threadfence_system();
    if (threadIdx.x < world) {
      XchgHeader* ph = reinterpret_cast<XchgHeader*>(peers.base[threadIdx.x]);
      st_release_sys(&ph->flags[par][rank], seq);
    }
  }
  // 3. wait for every rank's slot
  if (threadIdx.x < world) {
    // a peer that is only slow (checkpointing, validation) is waited for; the bound (options: xchg_timeout_ms, ten
    // minutes by default) only stops a dead peer from hanging the GPU, and is reported through the flag word
    // (IIC_FLAG_XCHG_TIMEOUT -> RuntimeError on the host) instead of a sticky context fault
    if (!xchg_wait_flag(hdr, par, threadIdx.x, seq, timeout_ns) && flags) atomicOr(flags, IIC_FLAG_XCHG_TIMEOUT);
  }
  __syncthreads();
  // 4. fixed-order sum (ld.cg: the slots were written by peers, L1 must not serve them)
  const double* src = reinterpret_cast<const double*>(peers.base[rank] + XCHG_HDR_BYTES) + (size_t)par * world * capacity;
  for (long long e = e0; e < E; e += stride) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcg(src + (size_t)r * capacity + e);
    J[e] = s;
  }
  // bookkeeping for the next exchange
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&hdr->done, 1u) == gridDim.x - 1) {
      hdr->arrive = 0;
      hdr->done = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(&hdr->seq) = seq;
    }
  }
}

}  // namespace iic

using namespace iic;

extern "C" size_t iic_xchg_buffer_bytes(int world, long long capacity) {
  if (world < 1 || world > XCHG_MAXR || capacity < 1) return 0;
  return XCHG_HDR_BYTES + (size_t)2 * world * (size_t)capacity * sizeof(double);
}

/* cudaMalloc + zero; *buf_out receives the device pointer */
extern "C" int iic_xchg_create(int world, long long capacity, void** buf_out) {
  IIC_REQUIRE(buf_out, "iic_xchg_create: null pointer");
  const size_t bytes = iic_xchg_buffer_bytes(world, capacity);
  IIC_REQUIRE(bytes > 0, "iic_xchg_create: bad world size %d or capacity %lld", world, capacity);
  void* p = nullptr;
  IIC_CHECK_CUDA(cudaMalloc(&p, bytes));
  IIC_CHECK_CUDA(cudaMemset(p, 0, bytes));
  IIC_CHECK_CUDA(cudaDeviceSynchronize());
  *buf_out = p;
  return 0;
}

/* 64-byte CUDA IPC handle of a buffer made by iic_xchg_create (host memory) */
extern "C" int iic_xchg_export(void* buf, void* handle64_host) {
  IIC_REQUIRE(buf && handle64_host, "iic_xchg_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  IIC_CHECK_CUDA(cudaIpcGetMemHandle(&h, buf));
  memcpy(handle64_host, &h, 64);
  return 0;
}

/* map a peer's buffer from its IPC handle; *peer_out receives a device pointer valid in this process */
extern "C" int iic_xchg_import(const void* handle64_host, void** peer_out) {
  IIC_REQUIRE(handle64_host && peer_out, "iic_xchg_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  IIC_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_out = p;
  return 0;
}

extern "C" int iic_xchg_release(void* buf, int imported) {
  if (!buf) return 0;
  if (imported) IIC_CHECK_CUDA(cudaIpcCloseMemHandle(buf));
  else IIC_CHECK_CUDA(cudaFree(buf));
  return 0;
}

/* J[e] <- sum over ranks of their J[e], e < E <= capacity, in place, fp64, rank order.  bufs_host[r] = this process's
 * mapping of rank r's buffer (bufs_host[rank] = the local one). */
extern "C" int iic_xchg_allreduce(double* J, long long E, long long capacity, void* const* bufs_host, int rank,
                                  int world, int* flags, void* stream) {
  IIC_REQUIRE(J && bufs_host, "iic_xchg_allreduce: null pointer");
  IIC_REQUIRE(world >= 1 && world <= XCHG_MAXR && rank >= 0 && rank < world, "iic_xchg_allreduce: bad rank %d / world %d", rank, world);
  IIC_REQUIRE(E >= 0 && E <= capacity, "iic_xchg_allreduce: %lld elements exceed the buffer capacity %lld", E, capacity);
  if (E == 0) return 0;
  XchgPeers peers;
  for (int r = 0; r < XCHG_MAXR; ++r) peers.base[r] = r < world ? (unsigned char*)bufs_host[r] : nullptr;
  for (int r = 0; r < world; ++r) IIC_REQUIRE(peers.base[r], "iic_xchg_allreduce: buffer of rank %d is not mapped", r);
  long long ctas = (E + 255) / 256;
  const int sms = sm_count_cached(current_device());
  // every CTA waits for the LAST CTA of the same launch to publish, so all of them must be resident at once: at most
  // half the SMs' worth of 256-thread CTAs (they fit beside anything that leaves 8 warps per SM free), one for small E
  long long cap = sms > 0 ? sms / 2 : 32;
  if (E <= 4096) cap = 1;
  if (ctas > cap) ctas = cap;
  const unsigned long long timeout_ns = (unsigned long long)options().xchg_timeout_ms * 1000000ull;
  if (ctas > 1) {
    // cooperative launch: the driver starts the grid only when every CTA can be resident, whatever runs on other streams
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, xchg_allreduce_kernel, J, E, capacity, peers, rank, world, timeout_ns, flags) == cudaSuccess)
      return 0;
    cudaGetLastError();          // e.g. a driver that cannot capture cooperative launches: plain launch below
  }
  xchg_allreduce_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(J, E, capacity, peers, rank, world, timeout_ns,
                                                                          flags);
  IIC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

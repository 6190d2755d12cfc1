// Process-wide plumbing of libiic_b200: error text, device queries, the run-time options and the per-device
// dynamic-shared-memory opt-in of the kernels.  No kernels here.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace iic {

// ---- error plumbing: nothing throws across the C ABI -------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int current_device() {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess) return -1;
  return d;
}
int sm_count_cached(int device) {
  static int cache[64];
  if (device < 0 || device >= 64) return -1;
  if (cache[device] > 0) return cache[device];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    set_error("cudaDeviceGetAttribute(MultiProcessorCount, %d) failed: %s", device,
              cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  cache[device] = v;
  return v;
}

// ---- options -----------------------------------------------------------------------------------------
// The dispatch switches are read from the environment ONCE, when the library first needs them, and live in
// this table afterwards; iic_b200_set_option changes them at run time (tests force a kernel family that way).
struct OptionEntry { const char* name; const char* env; int Options::*field; };
static const OptionEntry kOptions[] = {
    {"no_tma", "IIC_B200_NO_TMA", &Options::no_tma},
    {"no_tc", "IIC_B200_NO_TC", &Options::no_tc},
    {"no_tc10", "IIC_B200_NO_TC10", &Options::no_tc10},
    {"no_fast", "IIC_B200_NO_FAST", &Options::no_fast},
    {"tcp_p1", "IIC_B200_TCP_P1", &Options::tcp_p1},
    {"tcrb_p1", "IIC_B200_TCRB_P1", &Options::tcrb_p1},
    {"tc10_force", "IIC_B200_TC10_FORCE", &Options::tc10_force},
    {"no_tcj10", "IIC_B200_NO_TCJ10", &Options::no_tcj10},
    {"tc10_tf32", "IIC_B200_TC10_TF32", &Options::tc10_tf32},
    {"no_fused_epilogue", "IIC_B200_NO_FUSED_EPILOGUE", &Options::no_fused_epilogue},
    {"fin_last_cta_epilogue", "IIC_B200_FIN_LAST_CTA_EPILOGUE", &Options::fin_last_cta_epilogue},
    {"xchg_timeout_ms", "IIC_B200_XCHG_TIMEOUT_MS", &Options::xchg_timeout_ms},
};
static Options g_options;
static std::once_flag g_options_once;
static void load_options() {
  g_options.xchg_timeout_ms = 600000;        // ten minutes: a slow peer (checkpoint, validation) is not an error
  for (const OptionEntry& e : kOptions) {
    const char* v = getenv(e.env);
    if (v && *v) {
      const int n = atoi(v);
      g_options.*(e.field) = (n != 0 || v[0] == '0') ? n : 1;      // "1", "0", or any non-numeric text = on
    }
  }
}
Options& options_mut() {
  std::call_once(g_options_once, load_options);
  return g_options;
}
const Options& options() { return options_mut(); }

// ---- dynamic shared memory opt-in, per (device, kernel) ------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-context attribute: a process that drives several devices
// must set it on each, and autograd calls the backward entry points from its own threads.
int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, unsigned long long> done;      // kernel -> bit per device
  const int dev = current_device();
  if (dev < 0 || dev >= 64) {
    set_error("ensure_dyn_smem: no current CUDA device");
    return 1;
  }
  std::lock_guard<std::mutex> lock(mu);
  unsigned long long& mask = done[func];
  if (mask & (1ull << dev)) return 0;
  IIC_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  mask |= 1ull << dev;
  return 0;
}

}  // namespace iic

using namespace iic;

extern "C" int iic_b200_abi_version(void) { return IIC_B200_ABI_VERSION; }
extern "C" const char* iic_b200_last_error(void) { return get_error(); }
extern "C" int iic_b200_sm_count(int device) { return sm_count_cached(device); }

extern "C" int iic_b200_set_option(const char* name, int value) {
  IIC_REQUIRE(name, "iic_b200_set_option: null name");
  for (const OptionEntry& e : kOptions)
    if (strcmp(e.name, name) == 0) {
      options_mut().*(e.field) = value;
      return 0;
    }
  set_error("iic_b200_set_option: unknown option '%s'", name);
  return 2;
}
extern "C" int iic_b200_get_option(const char* name) {
  if (name)
    for (const OptionEntry& e : kOptions)
      if (strcmp(e.name, name) == 0) return options().*(e.field);
  set_error("iic_b200_get_option: unknown option '%s'", name ? name : "(null)");
  return -1;
}

// ABI version, error reporting, device gate.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace livae {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

static struct { void* ptr; int64_t bytes; } g_scratch[64] = {};
float* scratch_floats(int64_t nfloats) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  const auto& s = g_scratch[dev & 63];
  return (s.ptr && nfloats * 4 <= s.bytes) ? (float*)s.ptr : nullptr;
}

int require_sm100() {
  static bool ok[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && ok[dev & 63]) return 0;
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  if (major != 10) {
    set_error("liblivae_sm100 is built for sm_100a only; device has compute capability major %d", major);
    return -2;
  }
  ok[dev & 63] = true;
  return 0;
}
}  // namespace livae

extern "C" int livae_abi_version(void) { return 2; }
extern "C" int livae_set_scratch(void* ptr, int64_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { livae::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  if (((uintptr_t)ptr & 255) != 0 || bytes < 0) { livae::set_error("set_scratch: pointer must be 256-byte aligned"); return -1; }
  livae::g_scratch[dev & 63].ptr = ptr;
  livae::g_scratch[dev & 63].bytes = ptr ? bytes : 0;
  return 0;
}
extern "C" const char* livae_last_error(void) { return livae::g_err; }
namespace livae { long long launch_count(); }
extern "C" int64_t livae_launch_count(void) { return (int64_t)livae::launch_count(); }
extern "C" int livae_device_ok(void) { return livae::require_sm100() == 0 ? 1 : 0; }

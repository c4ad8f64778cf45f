// Error state, device check and launch accounting for the C ABI (include/nerfw.h).
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace nerfw {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}
bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  if (mask & (1ull << dev)) return false;
  mask |= 1ull << dev;
  return true;
}
}  // namespace nerfw

extern "C" {
const char* nerfw_last_error(void) { return nerfw::g_err; }
int nerfw_abi_version(void) { return NERFW_ABI_VERSION; }
uint64_t nerfw_launch_count(void) { return nerfw::g_launches.load(std::memory_order_relaxed); }

int nerfw_check_device(void) {
  int dev = 0, major = 0;
  NERFW_CUDA(cudaGetDevice(&dev));
  NERFW_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    nerfw::set_error("device %d has compute capability %d.x; libnerfw_sm100 only runs on sm_100 (B200)", dev, major);
    return NERFW_EDEVICE;
  }
  return NERFW_OK;
}
}

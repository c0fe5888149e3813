// common.cu -- error channel, version, device probe, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace kmu {
static thread_local char g_err[512] = "";
// process-wide: PyTorch runs backward nodes on its autograd thread, a thread_local counter would miss every backward launch
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_deterministic{-1};
bool deterministic() {
  int v = g_deterministic.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("KMU_DETERMINISTIC");
    v = (e && e[0] == '1') ? 1 : 0;
    g_deterministic.store(v, std::memory_order_relaxed);
  }
  return v == 1;
}
void set_deterministic(int on) { g_deterministic.store(on ? 1 : 0, std::memory_order_relaxed); }

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int finish_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return KMU_ERR_LAUNCH;
  }
  return KMU_OK;
}
}  // namespace kmu

extern "C" {
int kmu_version(void) { return KMU_VERSION; }
const char* kmu_last_error(void) { return kmu::g_err; }
void kmu_set_deterministic(int on) { kmu::set_deterministic(on); }
int kmu_get_deterministic(void) { return kmu::deterministic() ? 1 : 0; }
uint64_t kmu_launch_count(void) { return kmu::g_launches.load(std::memory_order_relaxed); }
int kmu_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
}

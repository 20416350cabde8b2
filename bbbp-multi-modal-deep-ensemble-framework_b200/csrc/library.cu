// Library-level entry points: ABI version, error string, device check.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include "common.cuh"

namespace bbbp {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int launch_status(const char* what) {
  note_launches(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return BBBP_ECUDA;
  }
  return BBBP_OK;
}

int current_sm_count() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}
}  // namespace bbbp

extern "C" int bbbp_abi_version(void) { return BBBP_ABI_VERSION; }
extern "C" uint64_t bbbp_launch_count(void) { return bbbp::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* bbbp_last_error(void) { return bbbp::g_error; }

extern "C" int bbbp_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    bbbp::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return BBBP_ECUDA;
  }
  if (prop.major != 10) {
    bbbp::set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    return BBBP_EUNSUPPORTED;
  }
  return BBBP_OK;
}

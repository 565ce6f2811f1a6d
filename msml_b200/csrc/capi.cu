// C-ABI plumbing: error string, launch counter, device queries.
#include <atomic>

#include "common.cuh"

namespace msml {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide: autograd backward runs on its own thread

char* err_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace msml

extern "C" int msml_abi_version(void) { return MSML_B200_ABI_VERSION; }
extern "C" const char* msml_last_error(void) { return msml::err_buf(); }
extern "C" int64_t msml_launch_count(void) { return msml::g_launches.load(); }
extern "C" void msml_launch_count_reset(void) { msml::g_launches.store(0); }

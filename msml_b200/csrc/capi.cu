// C-ABI plumbing: error string, launch counter, device queries.
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- profiling -------------------------------------------------------------------------------
struct ProfRec { const char* name; double work, bytes; cudaEvent_t a, b; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(const char* name, double work, cudaStream_t stream, double bytes) : slot(-1), st(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{name, work, bytes, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace msml

extern "C" void msml_profile_enable(int on) { msml::g_prof_on.store(on != 0); }

// Synchronises, then writes one line per kernel family: "name launches total_ms total_work total_min_bytes\n".
// Returns the number of bytes written (excluding the terminator) and clears the records.
extern "C" int64_t msml_profile_collect(char* buf, int64_t cap) {
  using namespace msml;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { int64_t n = 0; double ms = 0, work = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.n += 1; e.ms += ms; e.work += r.work; e.bytes += r.bytes;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), (long long)kv.second.n, kv.second.ms,
             kv.second.work, kv.second.bytes);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
    return (int64_t)n;
  }
  return 0;
}

extern "C" int msml_abi_version(void) { return MSML_B200_ABI_VERSION; }
extern "C" const char* msml_last_error(void) { return msml::err_buf(); }
extern "C" int64_t msml_launch_count(void) { return msml::g_launches.load(); }
extern "C" void msml_launch_count_reset(void) { msml::g_launches.store(0); }

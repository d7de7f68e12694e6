// Launch accounting and optional per-kernel device timing (include/nvfi_b200.h,
// "instrumentation").  Every kernel launch of the library goes through a ProfScope: it
// always counts the launch (bench.py reports `gpu_launches` from this counter) and, when
// profiling is enabled, brackets the launch with two CUDA events on the launching stream so
// that bench.py can attribute device time to kernels without running under a profiler.  It also
// opens an NVTX range named after the kernel around the launch (header-only NVTX3: a no-op unless a
// tool such as nsys / ncu --nvtx is attached).
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "nvfi_common.cuh"

namespace nvfi {

struct ProfRecord {
  const char* name;
  cudaEvent_t e0, e1;
};

static std::mutex g_mu;
static long long g_launches = 0;
static bool g_enabled = false;
static std::vector<ProfRecord> g_records;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(const char* name, cudaStream_t st) : st_(st), idx_(-1) {
  for (const char* p = name; *p; ++p)   // drop namespace qualifiers: "tcb::k_x" -> "k_x"
    if (p[0] == ':' && p[1] == ':') name = p + 2;
  nvtxRangePushA(name);
  std::lock_guard<std::mutex> lk(g_mu);
  ++g_launches;
  if (!g_enabled) return;
  ProfRecord r{name, get_event(), get_event()};
  cudaEventRecord(r.e0, st);
  idx_ = (int)g_records.size();
  g_records.push_back(r);
}

ProfScope::~ProfScope() {
  if (idx_ >= 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (idx_ < (int)g_records.size()) cudaEventRecord(g_records[idx_].e1, st_);
  }
  nvtxRangePop();
}

}  // namespace nvfi

using namespace nvfi;

extern "C" int64_t nvfi_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_launches;
}

extern "C" int nvfi_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_enabled = on != 0;
  return NVFI_OK;
}

extern "C" int nvfi_profile_read(NvfiProfileEntry* out, int cap, int reset) {
  if (cap < 0 || (cap > 0 && !out)) return NVFI_EINVAL;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return -100 - (int)e;
  std::lock_guard<std::mutex> lk(g_mu);
  int n = 0;
  for (const ProfRecord& r : g_records) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = 0.f;
    int j = 0;
    for (; j < n; ++j)
      if (std::string(out[j].name) == r.name) break;
    if (j == n) {
      if (n >= cap) continue;
      std::snprintf(out[n].name, sizeof(out[n].name), "%s", r.name);
      out[n].ms = 0.0;
      out[n].launches = 0;
      ++n;
    }
    out[j].ms += ms;
    out[j].launches += 1;
  }
  if (reset) {
    for (const ProfRecord& r : g_records) {
      g_pool.push_back(r.e0);
      g_pool.push_back(r.e1);
    }
    g_records.clear();
  }
  return n;
}

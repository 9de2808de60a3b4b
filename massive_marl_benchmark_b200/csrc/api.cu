// api.cu - ABI bookkeeping: version, error strings, launch counter, per-kernel event timing.
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/mmb.h"
#include "mmb_common.cuh"

namespace mmb {
namespace {
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_profile{0};
std::mutex g_mu;
struct EvPair { cudaEvent_t a, b; };
std::vector<EvPair> g_events[K_COUNT];
}  // namespace

int sm_count() {
  static std::atomic<int> cache[MMB_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= MMB_MAX_DEVICES) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

LaunchScope::LaunchScope(int id, cudaStream_t st) : id_(id), st_(st), stop_(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_profile.load(std::memory_order_relaxed)) {
    EvPair e;
    if (cudaEventCreate(&e.a) == cudaSuccess && cudaEventCreate(&e.b) == cudaSuccess) {
      cudaEventRecord(e.a, st);
      stop_ = e.b;
      std::lock_guard<std::mutex> lk(g_mu);
      g_events[id].push_back(e);
    }
  }
}
LaunchScope::~LaunchScope() {
  if (stop_) cudaEventRecord(stop_, st_);
}
}  // namespace mmb

extern "C" int32_t mmb_abi_version(void) { return MMB_ABI_VERSION; }

extern "C" uint64_t mmb_launch_count(void) { return mmb::g_launches.load(std::memory_order_relaxed); }

extern "C" int32_t mmb_profile_enable(int32_t on) {
  mmb::g_profile.store(on ? 1 : 0);
  return MMB_OK;
}

extern "C" int32_t mmb_profile_collect(int32_t kernel_id, double* total_ms, int64_t* count) {
  using namespace mmb;
  if (kernel_id < 0 || kernel_id >= K_COUNT || !total_ms || !count) return MMB_EINVAL;
  std::vector<EvPair> ev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    ev.swap(g_events[kernel_id]);
  }
  double tot = 0.0;
  int64_t n = 0;
  for (auto& e : ev) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      tot += ms;
      ++n;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  *total_ms = tot;
  *count = n;
  return MMB_OK;
}

extern "C" const char* mmb_strerror(int32_t status) {
  switch (status) {
    case MMB_OK: return "ok";
    case MMB_EINVAL: return "invalid argument (null pointer, non-positive size or inconsistent parameters)";
    case MMB_EALIGN: return "pointer or stride violates the documented alignment";
    case MMB_ECUDA: return "CUDA runtime error at launch";
    case MMB_EUNSUPPORTED: return "request not supported by this build";
    default: return "unknown mmb status";
  }
}

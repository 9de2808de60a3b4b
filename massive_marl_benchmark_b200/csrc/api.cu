// api.cu - ABI bookkeeping: version, error strings, launch counter.
#include <atomic>

#include "../../include/mmb.h"
#include "mmb_common.cuh"

namespace mmb {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace mmb

extern "C" int32_t mmb_abi_version(void) { return MMB_ABI_VERSION; }

extern "C" uint64_t mmb_launch_count(void) { return mmb::g_launches.load(std::memory_order_relaxed); }

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

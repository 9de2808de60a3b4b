// mmb_common.cuh - host-side helpers shared by the translation units of libmmb_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MMB_TEN_ANT_EPT
#define MMB_TEN_ANT_EPT 16  // environments per CTA tile of the TenAnt kernel (multiple of 4): 16 or 32
#endif
#define MMB_MAX_DEVICES 16

namespace mmb {

// kernel classes, as reported by mmb_profile_collect (keep in sync with include/mmb.h MMB_K_*)
enum KernelId {
  K_TEN_ANT = 0, K_TEN_ANT_CHAIN, K_TEN_ANT_CARRY, K_ONE_ANT, K_ONE_ANT_CHAIN, K_INGENUITY, K_INGENUITY_CHAIN,
  K_RESET, K_ROLLOUT_ADD, K_GAE_PPO, K_ADV_NORM, K_STATS, K_GAE_MARL, K_MASKS, K_GATHER, K_PERM, K_MLP_LAYER, K_LN_CAST, K_ADV_NORM_XCHG, K_EPISODE_SCAN, K_EPISODE_RING, K_GAUSS_ACT, K_PPO_LOSS, K_MAPPO_LOSS, K_ADAM_NORM, K_ADAM, K_COUNT
};

// Brackets one kernel launch: bumps the launch counter and, while profiling is enabled
// (mmb_profile_enable), records a CUDA event pair around it on the launch stream.
// SMs of the current device (cached per device; the grid caps below are multiples of it)
int sm_count();

struct LaunchScope {
  LaunchScope(int id, cudaStream_t st);
  ~LaunchScope();
  int id_;
  cudaStream_t st_;
  cudaEvent_t stop_;
};

}  // namespace mmb

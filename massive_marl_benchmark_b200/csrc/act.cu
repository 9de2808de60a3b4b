// act.cu - Gaussian policy head: sampling and log-probabilities of the rollout-time `act()` in one launch.
//
// Replaces, after the actor MLP has produced the mean:
//   PPO   ActorCritic.act (agents/algorithms/rl/ppo/module.py:73-87): MultivariateNormal(mean, scale_tril = diag(sigma^2))
//         .sample() and .log_prob() - the reference's quirk keeps: the effective standard deviation is sigma^2
//   MARL  DiagGaussian / FixedNormal (agents/algorithms/utils/distributions.py:94-117): std = sigmoid(log_std / x) * y,
//         per-dimension log-probs
// which torch runs as six to ten small kernels (randn, mul, add, pow, log, sum ...).  The standard normal draws come from
// Philox4x32-10 keyed by (seed, step) with the element index as counter + Box-Muller: same distribution as the reference's
// generator, a different stream (exact RNG parity with torch's global generator is not reproducible; the deterministic
// path, `noise` supplied by the caller, is what the parity tests use).
#include <stdlib.h>
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincosf(6.28318530717958647692f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// one warp per row: lanes stride over the action dimensions; per-row log-prob sum by shuffle
__global__ void __launch_bounds__(256) gaussian_act_kernel(const __grid_constant__ mmb_gaussian_act_params p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int A = p.num_rows > row ? p.act_dim : 0;   // rows past the end do nothing but stay for the block's barrier below
  const int64_t grp = p.std_group_rows > 0 ? (int64_t)(row / p.std_group_rows) * A : 0;
  const float* std = p.std + grp;
  // Launched as a programmatic dependent of the kernel in front of it (the MLP that writes the means): nothing is read or
  // WRITTEN before the wait - the outputs are fresh allocations, and the caching allocator may hand out a block the
  // kernel in front still reads.
  griddep_wait();
  // device-resident Philox counter (CUDA-graph replays): every block reads it here; the last block to retire advances it
  const uint64_t step = p.step_counter ? __ldcg(reinterpret_cast<const unsigned long long*>(p.step_counter)) : p.step;
  if (p.sigma_out)
    for (int j = lane; j < A; j += 32) p.sigma_out[(int64_t)row * A + j] = __ldg(p.sigma_src + grp + j);
  const float* mean = p.mean + (int64_t)row * p.mean_stride;
  float lp_sum = 0.0f;
  for (int j = lane; j < A; j += 32) {
    const float sd = __ldg(std + j);
    float z;
    if (p.noise) {
      z = __ldg(p.noise + (int64_t)row * A + j);
    } else if (p.deterministic) {
      z = 0.0f;
    } else {
      const uint64_t idx = (uint64_t)row * (uint64_t)A + (uint64_t)j;
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(idx >> 1), (uint32_t)(idx >> 33), (uint32_t)step, (uint32_t)(step >> 32)),
                                    make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
      const float2 n = box_muller(r.x, r.y);
      z = (idx & 1) ? n.y : n.x;
    }
    const float m = __ldg(mean + j);
    const float a = m + z * sd;
    p.actions[(int64_t)row * A + j] = a;
    // log N(a; m, sd) = -z^2/2 - log(sd) - log(sqrt(2 pi)) with z = (a - m) / sd as the distribution recomputes it
    const float zz = (a - m) / sd;
    const float lp = -0.5f * zz * zz - logf(sd) - 0.9189385332046727f;
    if (p.logp_per_dim) p.logp_per_dim[(int64_t)row * A + j] = lp;
    lp_sum += lp;
  }
  if (p.logp_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
    if (lane == 0 && A > 0) p.logp_sum[row] = lp_sum;
  }
  if (p.step_counter) {
    // Find the last block to retire with a two-level ticket: same-address atomics retire at ~13 ns each on B200 (512 blocks
    // on one word were 6.6 us of this kernel), so a block draws from lane blockIdx % 64 and only the last block of a lane
    // draws from the top word.  The very last block advances the step and leaves every ticket word at zero.
    __syncthreads();   // every warp of the block has read the counter
    if (threadIdx.x == 0) {
      unsigned long long* ctr = reinterpret_cast<unsigned long long*>(p.step_counter);
      constexpr unsigned LANES = MMB_ACT_TICKET_LANES;
      const unsigned lane_id = blockIdx.x % LANES;
      const unsigned lanes_used = gridDim.x < LANES ? gridDim.x : LANES;
      const unsigned in_lane = (gridDim.x - lane_id + LANES - 1) / LANES;   // blocks that draw from this lane
      __threadfence();
      if (atomicAdd(ctr + 2 + lane_id, 1ull) == (unsigned long long)in_lane - 1ull) {
        ctr[2 + lane_id] = 0ull;
        __threadfence();
        if (atomicAdd(ctr + 1, 1ull) == (unsigned long long)lanes_used - 1ull) {
          ctr[1] = 0ull;
          ctr[0] = step + 1ull;
        }
      }
    }
  }
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_gaussian_act(const mmb_gaussian_act_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_gaussian_act_params p = *pp;
  if (p.num_rows <= 0 || p.act_dim <= 0 || !p.mean || !p.std || !p.actions || p.mean_stride < p.act_dim || p.std_group_rows < 0) return MMB_EINVAL;
  if ((p.sigma_out != nullptr) != (p.sigma_src != nullptr)) return MMB_EINVAL;
  {
    LaunchScope ls(K_GAUSS_ACT, (cudaStream_t)stream);
    // optionally a programmatic dependent launch: the grid is scheduled while the kernel in front (the MLP that writes the
    // means) drains; everything that reads its output sits behind griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((int64_t)p.num_rows * 32 + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    // MMB_ACT_PDL=1: measured equal within noise to the ordinary launch behind the dual-network chain (act() graph-replayed
    // 52.0 vs 51.1 us at M = 4096), so it is off by default
    static const bool pdl = [] { const char* e = getenv("MMB_ACT_PDL"); return e && e[0] == '1'; }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, gaussian_act_kernel, p) != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

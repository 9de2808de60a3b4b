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

// Two standard normal draws from two 32-bit words.  Only the SAMPLE comes from here (the log-probability is computed from
// the action with IEEE arithmetic below), so the fast intrinsics do: |error| of __logf / __sincosf on these ranges is
// < 2^-21, far below what a random draw can resolve; the angle is taken in [-pi, pi), where __sincosf is accurate.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f) - 0.5f;     // [-0.5, 0.5)
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.28318530717958647692f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// log N(a; m, sd) = -z^2/2 - log(sd) - log(sqrt(2 pi)) with z = (a - m) / sd as the distribution recomputes it
__device__ __forceinline__ float normal_logp(float a, float m, float sd) {
  const float zz = (a - m) / sd;
  return -0.5f * zz * zz - logf(sd) - 0.9189385332046727f;
}

constexpr int ACT_THREADS = 1024;    // 32 rows in flight per block; the grid is capped at the SM count (grid-stride over rows)
constexpr unsigned ACT_TICKET_LANES = 16;
static_assert(ACT_TICKET_LANES <= MMB_ACT_TICKET_LANES, "ticket words of the caller's buffer");

// One warp per row.  PAIR (act_dim even, 8-byte aligned rows): a lane owns the element pairs (2 lane, 2 lane + 1) + 64 k -
// the two elements that share one Philox call and one Box-Muller transform, so each is computed once, with 8-byte accesses.
// Element idx of the [rows][act_dim] matrix takes word (idx & 1) of the transform of Philox counter idx >> 1: the stream
// does not depend on the path taken.
template <bool PAIR>
__global__ void __launch_bounds__(ACT_THREADS) gaussian_act_kernel(const __grid_constant__ mmb_gaussian_act_params p) {
  __shared__ unsigned long long s_step;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int A = p.act_dim;
  // Launched (optionally) as a programmatic dependent of the kernel in front of it (the MLP that writes the means): nothing is
  // read or WRITTEN before the wait - the outputs are fresh allocations, and the caching allocator may hand out a block the
  // kernel in front still reads.
  griddep_wait();
  uint64_t step = p.step;
  if (p.step_counter) {
    // Device-resident Philox step (CUDA-graph replays).  Thread 0 of every block reads it and THEN draws a ticket; the block
    // that draws the last ticket knows every block has read and advances the counter - right away, under the sampling work
    // of the whole grid, not at its end.  Two levels (same-address atomics retire at ~13 ns each on B200): a block draws from
    // lane blockIdx % 16, the last block of a lane from the top word.  Every ticket word is left at zero.
    unsigned long long* ctr = reinterpret_cast<unsigned long long*>(p.step_counter);
    if (threadIdx.x == 0) {
      const unsigned long long v = __ldcg(ctr);
      s_step = v;
      const unsigned lane_id = blockIdx.x % ACT_TICKET_LANES;
      const unsigned lanes_used = gridDim.x < ACT_TICKET_LANES ? gridDim.x : ACT_TICKET_LANES;
      const unsigned in_lane = (gridDim.x - lane_id + ACT_TICKET_LANES - 1) / ACT_TICKET_LANES;   // blocks that draw from this lane
      __threadfence();
      if (atomicAdd(ctr + 2 + lane_id, 1ull) == (unsigned long long)in_lane - 1ull) {
        ctr[2 + lane_id] = 0ull;
        __threadfence();
        if (atomicAdd(ctr + 1, 1ull) == (unsigned long long)lanes_used - 1ull) {
          ctr[1] = 0ull;
          ctr[0] = v + 1ull;
        }
      }
    }
    __syncthreads();
    step = s_step;
  }
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  const int mode = p.noise ? 0 : (p.deterministic ? 1 : 2);
  for (int64_t row = (int64_t)blockIdx.x * (ACT_THREADS / 32) + warp; row < p.num_rows; row += (int64_t)gridDim.x * (ACT_THREADS / 32)) {
    const int64_t grp = p.std_group_rows > 0 ? (row / p.std_group_rows) * A : 0;
    const float* std = p.std + grp;
    const float* mean = p.mean + row * p.mean_stride;
    const int64_t base = row * A;
    float lp_sum = 0.0f;
    if (PAIR) {
      for (int j = 2 * lane; j < A; j += 64) {
        const float2 sd = __ldg(reinterpret_cast<const float2*>(std + j));
        const float2 m = __ldg(reinterpret_cast<const float2*>(mean + j));
        float2 z = make_float2(0.0f, 0.0f);
        if (mode == 0) {
          z = __ldg(reinterpret_cast<const float2*>(p.noise + base + j));
        } else if (mode == 2) {
          const uint64_t c = (uint64_t)(base + j) >> 1;
          const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
          z = box_muller(r.x, r.y);
        }
        const float2 a = make_float2(m.x + z.x * sd.x, m.y + z.y * sd.y);
        *reinterpret_cast<float2*>(p.actions + base + j) = a;
        const float2 lp = make_float2(normal_logp(a.x, m.x, sd.x), normal_logp(a.y, m.y, sd.y));
        if (p.logp_per_dim) *reinterpret_cast<float2*>(p.logp_per_dim + base + j) = lp;
        if (p.sigma_out) *reinterpret_cast<float2*>(p.sigma_out + base + j) = __ldg(reinterpret_cast<const float2*>(p.sigma_src + grp + j));
        lp_sum += lp.x + lp.y;
      }
    } else {
      for (int j = lane; j < A; j += 32) {
        const float sd = __ldg(std + j);
        const float m = __ldg(mean + j);
        float z = 0.0f;
        if (mode == 0) {
          z = __ldg(p.noise + base + j);
        } else if (mode == 2) {
          const uint64_t idx = (uint64_t)(base + j), c = idx >> 1;
          const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
          const float2 n = box_muller(r.x, r.y);
          z = (idx & 1) ? n.y : n.x;
        }
        const float a = m + z * sd;
        p.actions[base + j] = a;
        const float lp = normal_logp(a, m, sd);
        if (p.logp_per_dim) p.logp_per_dim[base + j] = lp;
        if (p.sigma_out) p.sigma_out[base + j] = __ldg(p.sigma_src + grp + j);
        lp_sum += lp;
      }
    }
    if (p.logp_sum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
      if (lane == 0) p.logp_sum[row] = lp_sum;
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
    int64_t blocks = ((int64_t)p.num_rows + ACT_THREADS / 32 - 1) / (ACT_THREADS / 32);
    if (blocks > sm_count()) blocks = sm_count();
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(ACT_THREADS);
    cfg.stream = (cudaStream_t)stream;
    // MMB_ACT_PDL=1: measured equal within noise to the ordinary launch behind the dual-network chain (act() graph-replayed
    // 52.0 vs 51.1 us at M = 4096), so it is off by default
    static const bool pdl = [] { const char* e = getenv("MMB_ACT_PDL"); return e && e[0] == '1'; }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
    const bool pair = (p.act_dim & 1) == 0 && (p.mean_stride & 1) == 0 && al8(p.mean) && al8(p.std) && al8(p.actions) && al8(p.noise) &&
                      al8(p.logp_per_dim) && al8(p.sigma_src) && al8(p.sigma_out);
    const cudaError_t le = pair ? cudaLaunchKernelEx(&cfg, gaussian_act_kernel<true>, p) : cudaLaunchKernelEx(&cfg, gaussian_act_kernel<false>, p);
    if (le != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

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

// PAIR (act_dim even, 8-byte aligned rows): the unit of work is the element pair (2 k, 2 k + 1) - the two elements that share
// one Philox call and one Box-Muller transform, so each is computed once, with 8-byte accesses; a warp walks the pairs of R
// consecutive rows.  Otherwise: one warp per row, one element per lane and pass.
// Element idx of the [rows][act_dim] matrix takes word (idx & 1) of the transform of Philox counter idx >> 1: the stream
// does not depend on the path taken.
template <bool PAIR>
__global__ void __launch_bounds__(ACT_THREADS) gaussian_act_kernel(const __grid_constant__ mmb_gaussian_act_params p, const int R) {
  extern __shared__ float lp_s[];        // PAIR, R > 1: [warps][R * act_dim / 2] log-probs of the pairs of a warp's row group
  __shared__ unsigned long long s_step;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
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
  if (PAIR) {
    // A warp owns R consecutive rows = R * hp element pairs and walks them 32 at a time, so no lane idles when a row's pair
    // count is not a multiple of 32 (80 actions: four rows in five full passes instead of eight partial ones; 8 actions:
    // eight rows per pass).  Row sums in a fixed order (lane partials over pairs k, k + 32, ... of the row, then the tree).
    const int hp = A >> 1;
    float* my = lp_s + (size_t)warp * R * hp;
    const int l_r = lane / hp, l_c = lane - l_r * hp;        // (row, pair) of this lane's first pair; advance per pass below
    const int d_r = 32 / hp, d_c = 32 - d_r * hp;
    for (int64_t row0 = ((int64_t)blockIdx.x * warps + warp) * R; row0 < p.num_rows; row0 += (int64_t)gridDim.x * warps * R) {
      const int nr = (p.num_rows - row0) < R ? (int)(p.num_rows - row0) : R;
      const int total = nr * hp;
      int r = l_r, c = l_c;
      float lp_sum = 0.0f;
      for (int q = lane; q < total; q += 32) {
        const int row = (int)row0 + r, j = 2 * c;
        const int64_t grp = p.std_group_rows > 0 ? (int64_t)(row / p.std_group_rows) * A : 0;
        const int64_t base = (int64_t)row * A;
        const float2 sd = __ldg(reinterpret_cast<const float2*>(p.std + grp + j));
        const float2 m = __ldg(reinterpret_cast<const float2*>(p.mean + (int64_t)row * p.mean_stride + j));
        float2 z = make_float2(0.0f, 0.0f);
        if (mode == 0) {
          z = __ldg(reinterpret_cast<const float2*>(p.noise + base + j));
        } else if (mode == 2) {
          const uint64_t cnt = (uint64_t)(base + j) >> 1;
          const uint4 rnd = philox4x32_10(make_uint4((uint32_t)cnt, (uint32_t)(cnt >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
          z = box_muller(rnd.x, rnd.y);
        }
        const float2 a = make_float2(m.x + z.x * sd.x, m.y + z.y * sd.y);
        *reinterpret_cast<float2*>(p.actions + base + j) = a;
        const float2 lp = make_float2(normal_logp(a.x, m.x, sd.x), normal_logp(a.y, m.y, sd.y));
        if (p.logp_per_dim) *reinterpret_cast<float2*>(p.logp_per_dim + base + j) = lp;
        if (p.sigma_out) *reinterpret_cast<float2*>(p.sigma_out + base + j) = __ldg(reinterpret_cast<const float2*>(p.sigma_src + grp + j));
        if (R > 1) my[q] = lp.x + lp.y;
        else lp_sum += lp.x + lp.y;
        r += d_r; c += d_c;
        if (c >= hp) { c -= hp; ++r; }
      }
      if (p.logp_sum) {
        if (R > 1) {
          __syncwarp();
          for (int rr = 0; rr < nr; ++rr) {
            float s = 0.0f;
            for (int k = lane; k < hp; k += 32) s += my[rr * hp + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) p.logp_sum[row0 + rr] = s;
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
          if (lane == 0) p.logp_sum[row0] = lp_sum;
        }
      }
    }
  } else {
    for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < p.num_rows; row += (int64_t)gridDim.x * warps) {
      const int64_t grp = p.std_group_rows > 0 ? (row / p.std_group_rows) * A : 0;
      const float* std = p.std + grp;
      const float* mean = p.mean + row * p.mean_stride;
      const int64_t base = row * A;
      float lp_sum = 0.0f;
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
      if (p.logp_sum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lp_sum += __shfl_xor_sync(0xffffffffu, lp_sum, o);
        if (lane == 0) p.logp_sum[row] = lp_sum;
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
    auto al8 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
    const bool pair = (p.act_dim & 1) == 0 && (p.mean_stride & 1) == 0 && al8(p.mean) && al8(p.std) && al8(p.actions) && al8(p.noise) &&
                      al8(p.logp_per_dim) && al8(p.sigma_src) && al8(p.sigma_out);
    // rows per warp (pair path): the smallest R <= 16 whose R * act_dim / 2 pairs fill the 32-lane passes to >= 95 %, else the
    // best; the row group's log-probs go through shared memory (R > 1), which bounds R for wide rows
    // ... and never so large that the grid runs short of warps: the kernel hides its latencies (Philox chains, IEEE log /
    // divide sequences) by parallelism - at 4096 x 80 one row per warp (4096 warps, 5/8 of the lanes busy) measured 1.7 us
    // FASTER than four rows per warp (1024 warps, every lane busy)
    int R = 1;
    if (pair) {
      const int hp = p.act_dim / 2;
      const int64_t r_cap = (int64_t)p.num_rows / ((int64_t)sm_count() * 28);
      double best = 0.0;
      for (int r = 1; r <= 16 && r <= (r_cap < 1 ? 1 : r_cap); ++r) {
        if (r > 1 && (int64_t)r * hp * 4 * 4 > 40 * 1024) break;            // four warps at least
        const double eff = (double)r * hp / (32.0 * ((r * hp + 31) / 32));
        if (eff > best + 1e-9) { best = eff; R = r; }
        if (eff >= 0.95) break;
      }
    }
    // block size: the largest that still gives (nearly) every SM a block and fits the shared memory
    int threads = ACT_THREADS;
    const int64_t per_warp = R;
    while (threads > 128 && (((int64_t)p.num_rows + (threads / 32) * per_warp - 1) / ((threads / 32) * per_warp) < (int64_t)(sm_count() * 4) / 5 ||
                             (R > 1 && (int64_t)(threads / 32) * R * (p.act_dim / 2) * 4 > 40 * 1024)))
      threads >>= 1;
    cudaLaunchConfig_t cfg = {};
    int64_t blocks = ((int64_t)p.num_rows + (threads / 32) * per_warp - 1) / ((threads / 32) * per_warp);
    if (blocks > sm_count()) blocks = sm_count();
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (pair && R > 1) ? (size_t)(threads / 32) * R * (p.act_dim / 2) * 4 : 0;
    cfg.stream = (cudaStream_t)stream;
    // MMB_ACT_PDL=1: measured equal within noise to the ordinary launch behind the dual-network chain (act() graph-replayed
    // 52.0 vs 51.1 us at M = 4096), so it is off by default
    static const bool pdl = [] { const char* e = getenv("MMB_ACT_PDL"); return e && e[0] == '1'; }();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t le = pair ? cudaLaunchKernelEx(&cfg, gaussian_act_kernel<true>, p, R) : cudaLaunchKernelEx(&cfg, gaussian_act_kernel<false>, p, R);
    if (le != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

// episode.cu - device-side episode bookkeeping of the PPO runner.
//
// Replaces the per-step block of agents/algorithms/rl/ppo/ppo.py:143-157:
//     cur_reward_sum += rews; cur_episode_length += 1
//     new_ids = (dones > 0).nonzero()
//     reward_sum.extend(cur_reward_sum[new_ids].cpu().tolist()); episode_length.extend(...)      <- host sync per step
//     cur_reward_sum[new_ids] = 0; cur_episode_length[new_ids] = 0
//     rewbuffer.extend(reward_sum); lenbuffer.extend(episode_length)       # deque(maxlen=100)
// for T steps at once, without leaving the device.  The finished episodes are appended in the reference's order
// (step-major, env ascending within a step) to two rings of `window` entries = the deques whose mean is logged
// (ppo.py:198-220).
//
//   episode_scan_kernel   thread per env walks t = 0..T-1 (fp32 `+=` in the reference's order): running sum / length
//                         in, out; at a done flag the finished (sum, length) go to the [T][N] scratch planes.
//   episode_ring_kernel   one CTA per step row: counts of the earlier rows (redundantly, no inter-CTA sync), ordered
//                         rank of its own done envs by warp ballot + popc prefix, ring position = finished-before +
//                         row offset + rank; only the last `window` entries of the update are written (the earlier
//                         ones would be overwritten anyway), so every ring slot has exactly one writer.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

__device__ __forceinline__ bool done_at(const mmb_episode_params& p, int t, int e) {
  if (p.dones_u8) return __ldg(p.dones_u8 + (int64_t)t * p.dones_u8_row_stride + e) != 0;
  return __ldg(p.dones_i64 + (int64_t)t * p.dones_i64_row_stride + e) > 0;
}

__global__ void __launch_bounds__(256) episode_scan_kernel(const __grid_constant__ mmb_episode_params p) {
  const int N = p.num_envs, T = p.num_steps;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.state[1] = p.state[0];  // snapshot of "finished so far" for the ring kernel
  for (int64_t e64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e64 < N; e64 += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)e64;
    float sum = p.cur_reward_sum[e], len = p.cur_episode_length[e];
    for (int t = 0; t < T; ++t) {
      sum = fadd(sum, __ldg(p.rewards + (int64_t)t * p.rewards_row_stride + e));
      len = fadd(len, 1.0f);
      if (done_at(p, t, e)) {
        p.ep_reward[(int64_t)t * N + e] = sum;
        p.ep_length[(int64_t)t * N + e] = len;
        sum = 0.0f;
        len = 0.0f;
      }
    }
    p.cur_reward_sum[e] = sum;
    p.cur_episode_length[e] = len;
  }
}

// number of done flags in row t, by the whole CTA; result in every thread
__device__ __forceinline__ int row_count(const mmb_episode_params& p, int t, int* warp_tot) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  int c = 0;
  for (int e = tid; e < p.num_envs; e += blockDim.x) c += done_at(p, t, e) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __syncthreads();
  if (lane == 0) warp_tot[wid] = c;
  __syncthreads();
  int tot = 0;
  for (int w = 0; w < nw; ++w) tot += warp_tot[w];
  return tot;
}

__global__ void __launch_bounds__(1024) episode_ring_kernel(const __grid_constant__ mmb_episode_params p) {
  __shared__ int warp_tot[32];
  __shared__ int s_running;
  const int t = blockIdx.x, T = p.num_steps, N = p.num_envs;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int64_t before = 0, total_new = 0;
  for (int r = 0; r < T; ++r) {
    const int c = row_count(p, r, warp_tot);
    if (r < t) before += c;
    total_new += c;
  }
  const int64_t base = (int64_t)p.state[1];            // episodes finished before this update
  const int64_t first_kept = total_new - p.window;     // entries with a smaller index would be overwritten
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int e0 = 0; e0 < N; e0 += blockDim.x) {
    const int e = e0 + tid;
    const bool d = e < N && done_at(p, t, e);
    const unsigned bal = __ballot_sync(0xffffffffu, d);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = s_running;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (d) {
      const int64_t idx = before + off + __popc(bal & ((1u << lane) - 1u));   // position inside this update
      if (idx >= first_kept) {
        const int64_t slot = (base + idx) % p.window;
        p.reward_ring[slot] = p.ep_reward[(int64_t)t * N + e];
        p.length_ring[slot] = p.ep_length[(int64_t)t * N + e];
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w];
      s_running += tot;
    }
    __syncthreads();
  }
  if (t == 0 && tid == 0) p.state[0] = (uint64_t)(base + total_new);
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_episode_update(const mmb_episode_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_episode_params p = *pp;
  if (p.num_envs <= 0 || p.num_steps <= 0 || p.window <= 0) return MMB_EINVAL;
  if (!p.rewards || (!p.dones_u8 && !p.dones_i64) || !p.cur_reward_sum || !p.cur_episode_length || !p.ep_reward || !p.ep_length ||
      !p.reward_ring || !p.length_ring || !p.state)
    return MMB_EINVAL;
  if (p.num_steps > 65535) return MMB_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope ls(K_EPISODE_SCAN, st);
    int blocks = (p.num_envs + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    episode_scan_kernel<<<blocks, 256, 0, st>>>(p);
  }
  if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  {
    LaunchScope ls(K_EPISODE_RING, st);
    episode_ring_kernel<<<p.num_steps, 1024, 0, st>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

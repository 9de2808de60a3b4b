// ingenuity.cu - fused MultiIngenuity env-step kernel for sm_100a (one thread per helicopter).
//
// Replaces multi_ingenuity.py:268-339 (thrust -> rigid-body force tensor), :351-357 (observations = raw
// root rows), :359-374 + jit :381-453 (reward/reset), :341-349 (progress) and the clamps of the vec-task
// wrappers.  The reset bit depends on a computed float (`target_dist > 8.0`), so the 3-element
// sum uses the association of the torch device selected by `flavor` (mmb_math.cuh).
//
// CTA = tile of 64 envs (256 threads) of one frame.  The root tile (52 contiguous floats per env) is staged
// in shared memory with 128-bit loads and doubles as the observation tile (obs = root rows verbatim); the
// force tile (24 bodies x 3) is assembled in shared memory and leaves with 128-bit stores so every
// 288-byte env row is written whole instead of as eight scattered 12-byte pieces.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int EPT = 64;
constexpr int H = 4;
constexpr int NT = EPT * H;
constexpr int ROOT_ENV = 52;
constexpr int FORCE_ENV = 72;

template <int FLAVOR>
__global__ void __launch_bounds__(NT) ingenuity_kernel(const __grid_constant__ mmb_ingenuity_params p, const int pf_dist) {
  __shared__ __align__(16) float root_s[EPT * ROOT_ENV];
  __shared__ __align__(16) float force_s[EPT * FORCE_ENV];
  __shared__ unsigned char flagged_s[EPT];  // env had its reset flag set on entry (T == 1 only)
  const int tid = threadIdx.x, t = blockIdx.y;
  const int N = p.num_envs, T = p.num_frames;
  const int e0 = blockIdx.x * EPT;
  const int ne = min(EPT, N - e0);
  const int el = tid >> 2, h = tid & 3;
  const int e = e0 + el;
  const bool active = el < ne;

  tile_load(root_s, p.root + (int64_t)t * p.root_frame_stride + (int64_t)e0 * ROOT_ENV, ne * ROOT_ENV, tid, NT);
  if (tid == 32) {  // L2 prefetch of the inputs of the unit two CTAs per SM ahead in launch order (see ten_ant.cu)
    const int64_t u = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_dist;
    const int64_t t2 = u / gridDim.x, tile2 = u - t2 * gridDim.x;
    if (t2 < T && (tile2 + 1) * EPT <= N) {
      prefetch_range_l2(p.root + t2 * p.root_frame_stride + tile2 * EPT * ROOT_ENV, EPT * ROOT_ENV * 4);
      prefetch_range_l2(p.actions + t2 * p.actions_frame_stride + tile2 * EPT * 24, EPT * 24 * 4);
    }
  }
  for (int i = tid; i < EPT * FORCE_ENV; i += NT) force_s[i] = 0.0f;
  if (tid < EPT) flagged_s[tid] = 0;

  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float* ap = p.actions + (int64_t)t * p.actions_frame_stride + (int64_t)e * 24 + 6 * h;
    if ((reinterpret_cast<uintptr_t>(ap) & 7u) == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { float2 v = __ldg(reinterpret_cast<const float2*>(ap) + j); a[2 * j] = v.x; a[2 * j + 1] = v.y; }
    } else {
#pragma unroll
      for (int j = 0; j < 6; ++j) a[j] = __ldg(ap + j);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) a[j] = clampf(a[j], -p.clip_actions, p.clip_actions);
  }
  __syncthreads();

  float pos_r = 0.f, up_r = 0.f, spin_r = 0.f;
  bool die = false;
  if (active) {
    // thrusts (multi_ingenuity.py:272-327), rotor a = actions[6h+0..2], rotor b = actions[6h+3..5]
    const float lim = p.thrust_upper_limit, lat = p.thrust_lateral_component;
#pragma unroll
    for (int rtr = 0; rtr < 2; ++rtr) {
      const float vert = clampf(fmul(a[3 * rtr + 2], p.thrust_action_speed_scale), -lim, lim);
      const float lx = clampf(a[3 * rtr], -lat, lat), ly = clampf(a[3 * rtr + 1], -lat, lat);
      const float tz = fmul(p.dt, vert);
      float* f = force_s + el * FORCE_ENV + (6 * h + 1 + 2 * rtr) * 3;  // bodies 1,3 | 7,9 | 13,15 | 19,21
      f[0] = fmul(tz, lx); f[1] = fmul(tz, ly); f[2] = tz;
    }
    // reward terms (multi_ingenuity.py:388-436)
    const float* r = root_s + el * ROOT_ENV + h * 13;
    const float dx = fsub(p.goals[h][0], r[0]), dy = fsub(p.goals[h][1], r[1]), dz = fsub(p.goals[h][2], r[2]);
    const float td = fsqrt(sum3<FLAVOR>(fmul(dx, dx), fmul(dy, dy), fmul(dz, dz)));
    pos_r = fdiv(1.0f, fadd(1.0f, fmul(td, td)));
    const f3 ups = quat_rot<false>(f4{r[3], r[4], r[5], r[6]}, f3{0.0f, 0.0f, 1.0f});  // quat_axis(q, 2)
    const float tilt = fabsf(fsub(1.0f, ups.z));
    up_r = fdiv(5.0f, fadd(1.0f, fmul(tilt, tilt)));
    const float spin = fabsf(r[12]);
    spin_r = fdiv(1.0f, fadd(1.0f, fmul(spin, spin)));
    die = (td > 8.0f) || (r[2] < 0.5f);
  }
  // ordered sums over the 4 helicopters of an env: lanes 4m..4m+3 of a warp
  const unsigned full = 0xffffffffu;
  float p1 = __shfl_down_sync(full, pos_r, 1, 4), p2 = __shfl_down_sync(full, pos_r, 2, 4), p3 = __shfl_down_sync(full, pos_r, 3, 4);
  float u1 = __shfl_down_sync(full, up_r, 1, 4), u2 = __shfl_down_sync(full, up_r, 2, 4), u3 = __shfl_down_sync(full, up_r, 3, 4);
  float s1 = __shfl_down_sync(full, spin_r, 1, 4), s2 = __shfl_down_sync(full, spin_r, 2, 4), s3 = __shfl_down_sync(full, spin_r, 3, 4);
  const unsigned bal = __ballot_sync(full, die);
  if (active && h == 0) {
    const float pr = fadd(fadd(fadd(pos_r, p1), p2), p3);
    const float ur = fadd(fadd(fadd(up_r, u1), u2), u3);
    const float sr = fadd(fadd(fadd(spin_r, s1), s2), s3);
    const float reward = fadd(pr, fmul(pr, fadd(ur, sr)));
    const bool any_die = ((bal >> ((tid & 31) & ~3)) & 0xfu) != 0;
    if (p.rewards) p.rewards[(int64_t)t * p.rewards_frame_stride + e] = reward;
    if (T == 1) {
      int64_t prog = p.progress_buf[e] + 1;
      const bool flagged = p.reset_buf[e] != 0;
      if (flagged) prog = 0;
      int64_t rs = any_die ? 1 : 0;
      if ((float)prog >= (float)((double)p.max_episode_length - 1.0)) rs = 1;
      p.progress_buf[e] = prog;
      p.reset_buf[e] = rs;
      if (p.dones_i64) p.dones_i64[e] = rs;
      if (p.dones_u8) p.dones_u8[e] = (uint8_t)rs;
      // task.forces after the step: reset_idx zeroes the rows of envs flagged on entry (multi_ingenuity.py:243-244)
      flagged_s[el] = flagged ? 1 : 0;
    } else {
      if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] = any_die ? 1 : 0;
      else p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] = any_die ? 1 : 0;
    }
  }
  __syncthreads();

  // ---- tiles out ----
  const float clip = p.clip_obs;
  const int n = ne * ROOT_ENV;
  if (p.obs_raw) {
    float* g = p.obs_raw + (int64_t)t * p.obs_raw_frame_stride + (int64_t)e0 * ROOT_ENV;
    if (aligned16(g)) { for (int i = tid; i < (n >> 2); i += NT) stg4(g + 4 * i, reinterpret_cast<const float4*>(root_s)[i]); }
    else { for (int i = tid; i < n; i += NT) g[i] = root_s[i]; }
  }
  if (p.obs) {
    float* g = p.obs + (int64_t)t * p.obs_frame_stride + (int64_t)e0 * ROOT_ENV;
    if (aligned16(g)) {
      for (int i = tid; i < (n >> 2); i += NT) {
        float4 v = reinterpret_cast<const float4*>(root_s)[i];
        v.x = clampf(v.x, -clip, clip); v.y = clampf(v.y, -clip, clip); v.z = clampf(v.z, -clip, clip); v.w = clampf(v.w, -clip, clip);
        stg4(g + 4 * i, v);
      }
    } else {
      for (int i = tid; i < n; i += NT) g[i] = clampf(root_s[i], -clip, clip);
    }
  }
  const int nf = ne * FORCE_ENV;
  if (p.forces) {
    float* g = p.forces + (int64_t)t * p.forces_frame_stride + (int64_t)e0 * FORCE_ENV;
    if (aligned16(g)) { for (int i = tid; i < (nf >> 2); i += NT) stg4(g + 4 * i, reinterpret_cast<const float4*>(force_s)[i]); }
    else { for (int i = tid; i < nf; i += NT) g[i] = force_s[i]; }
  }
  if (p.forces_state && t == T - 1) {
    float* g = p.forces_state + (int64_t)e0 * FORCE_ENV;
    for (int i = tid; i < nf; i += NT) g[i] = flagged_s[i / FORCE_ENV] ? 0.0f : force_s[i];
  }
}

__global__ void ingenuity_chain_kernel(const __grid_constant__ mmb_ingenuity_params p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.num_envs) return;
  const int T = p.num_frames;
  int64_t prog = p.progress_buf[e];
  bool flag = p.reset_buf[e] != 0;
  const float thr = (float)((double)p.max_episode_length - 1.0);
  bool flagged_last = false;
  for (int t = 0; t < T; ++t) {
    flagged_last = flag;
    prog = flag ? 0 : prog + 1;
    bool die = p.dones_u8 ? (p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] != 0)
                          : (p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] != 0);
    flag = die || ((float)prog >= thr);
    if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] = flag ? 1 : 0;
    if (p.dones_i64) p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] = flag ? 1 : 0;
  }
  p.progress_buf[e] = prog;
  p.reset_buf[e] = flag ? 1 : 0;
  if (p.forces_state && flagged_last)
    for (int i = 0; i < FORCE_ENV; ++i) p.forces_state[(int64_t)e * FORCE_ENV + i] = 0.0f;
}

}  // namespace
}  // namespace mmb

extern "C" int32_t mmb_ingenuity_step(const mmb_ingenuity_params* pp, void* stream) {
  using namespace mmb;
  if (!pp) return MMB_EINVAL;
  mmb_ingenuity_params p = *pp;
  if (p.num_envs <= 0 || p.num_frames <= 0) return MMB_EINVAL;
  if (!p.root || !p.actions || !p.progress_buf || !p.reset_buf) return MMB_EINVAL;
  if (p.num_frames > 65535) return MMB_EUNSUPPORTED;
  if (p.num_frames > 1 && !p.dones_u8 && !p.dones_i64) return MMB_EINVAL;
  if (p.flavor != MMB_FLAVOR_CUDA && p.flavor != MMB_FLAVOR_CPU) return MMB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((p.num_envs + EPT - 1) / EPT, p.num_frames);
  {
    LaunchScope ls(K_INGENUITY, st);
    if (p.flavor == MMB_FLAVOR_CUDA) ingenuity_kernel<FLAVOR_CUDA><<<grid, NT, 0, st>>>(p, 2 * sm_count());
    else ingenuity_kernel<FLAVOR_CPU><<<grid, NT, 0, st>>>(p, 2 * sm_count());
  }
  if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  if (p.num_frames > 1) {
    {
      LaunchScope ls(K_INGENUITY_CHAIN, st);
      ingenuity_chain_kernel<<<(p.num_envs + 255) / 256, 256, 0, st>>>(p);
    }
    if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  }
  return MMB_OK;
}

// one_ant.cu - fused OneAnt env-step kernel for sm_100a (one thread per environment).
//
// Replaces one_ant.py:396-400 (forces), :346-361 + jit :563-627 (observations, potentials, box),
// :314-344 + jit :465-560 (reward/reset), :403-415 (progress, carry) and the clamps of
// VecTaskPython.step (vec_task.py:126-131).
//
// CTA = tile of EPT envs of one frame t (grid = tiles x T).  The root tile (2 rows x 13 floats per env,
// contiguous) goes through shared memory with 128-bit loads; dof (64 B), sensor (96 B) and actions (32 B)
// are private contiguous chunks read with LDG.128; the 60-wide obs rows leave through a shared-memory
// tile with 128-bit stores.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int EPT = 128;
constexpr int ROOT_ENV = 26;
constexpr int OBS_ENV = 60;
constexpr int OBS_PAD = 61;  // odd row pitch in shared memory: conflict-free per-thread row writes

template <int FLAVOR>
__device__ __forceinline__ float potential_of(float bx, float by, float dt) {
  // one_ant.py:589-593: -norm((0-bx, 0-by, 0)) / dt
  float tx = fsub(0.0f, bx), ty = fsub(0.0f, by);
  float nrm = (FLAVOR == FLAVOR_CPU) ? fsqrt(__fmaf_rn(ty, ty, fmul(tx, tx))) : fsqrt(fadd(fmul(tx, tx), fmul(ty, ty)));
  // Tensor / python scalar: CUDA multiplies by the fp32 reciprocal, CPU divides
  return (FLAVOR == FLAVOR_CUDA) ? fmul(-nrm, fdiv(1.0f, dt)) : fdiv(-nrm, dt);
}

template <int FLAVOR>
__global__ void __launch_bounds__(EPT) one_ant_kernel(const __grid_constant__ mmb_one_ant_params p, const int pf_dist) {
  __shared__ __align__(16) float root_s[EPT * ROOT_ENV];
  __shared__ __align__(16) float obs_s[EPT * OBS_PAD];
  const int tid = threadIdx.x, t = blockIdx.y;
  const int N = p.num_envs, T = p.num_frames;
  const int e0 = blockIdx.x * EPT;
  const int ne = min(EPT, N - e0);
  const int e = e0 + tid;
  const bool active = tid < ne;
  const mmb_ant_consts& c = p.c;

  tile_load(root_s, p.root + (int64_t)t * p.root_frame_stride + (int64_t)e0 * ROOT_ENV, ne * ROOT_ENV, tid, EPT);
  if (tid == 32) {  // L2 prefetch of the inputs of the unit two CTAs per SM ahead in launch order (see ten_ant.cu)
    const int64_t u = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + pf_dist;
    const int64_t t2 = u / gridDim.x, tile2 = u - t2 * gridDim.x;
    if (t2 < T && (tile2 + 1) * EPT <= N) {
      prefetch_range_l2(p.root + t2 * p.root_frame_stride + tile2 * EPT * ROOT_ENV, EPT * ROOT_ENV * 4);
      prefetch_range_l2(p.dof + t2 * p.dof_frame_stride + tile2 * EPT * 16, EPT * 16 * 4);
      prefetch_range_l2(p.sensor + t2 * p.sensor_frame_stride + tile2 * EPT * 24, EPT * 24 * 4);
      prefetch_range_l2(p.actions + t2 * p.actions_frame_stride + tile2 * EPT * 8, EPT * 8 * 4);
    }
  }

  float raw[16], sens[24], act[8];
  float pbx = 0.f, pby = 0.f, bbx = 0.f, bby = 0.f;
  if (active) {
    const float* d = p.dof + (int64_t)t * p.dof_frame_stride + (int64_t)e * 16;
    const float* s = p.sensor + (int64_t)t * p.sensor_frame_stride + (int64_t)e * 24;
    const float* a = p.actions + (int64_t)t * p.actions_frame_stride + (int64_t)e * 8;
    if (aligned16(d)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { float4 v = ldg4(d + 4 * j); raw[4 * j] = v.x; raw[4 * j + 1] = v.y; raw[4 * j + 2] = v.z; raw[4 * j + 3] = v.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) raw[j] = __ldg(d + j);
    }
    if (aligned16(s)) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { float4 v = ldg4(s + 4 * j); sens[4 * j] = v.x; sens[4 * j + 1] = v.y; sens[4 * j + 2] = v.z; sens[4 * j + 3] = v.w; }
    } else {
#pragma unroll
      for (int j = 0; j < 24; ++j) sens[j] = __ldg(s + j);
    }
    if (aligned16(a)) {
      float4 v0 = ldg4(a), v1 = ldg4(a + 4);
      act[0] = v0.x; act[1] = v0.y; act[2] = v0.z; act[3] = v0.w; act[4] = v1.x; act[5] = v1.y; act[6] = v1.z; act[7] = v1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) act[j] = __ldg(a + j);
    }
    if (t == 0) {
      pbx = __ldg(p.pos_before + 2 * (int64_t)e); pby = __ldg(p.pos_before + 2 * (int64_t)e + 1);
      bbx = __ldg(p.box_before + 2 * (int64_t)e); bby = __ldg(p.box_before + 2 * (int64_t)e + 1);
    } else {  // carry = ant / box xy of frame t-1 (one_ant.py:412-413)
      const float* rp = p.root + (int64_t)(t - 1) * p.root_frame_stride + (int64_t)e * ROOT_ENV;
      pbx = __ldg(rp); pby = __ldg(rp + 1); bbx = __ldg(rp + 13); bby = __ldg(rp + 14);
    }
  }
  __syncthreads();

  if (active) {
    const float* r = root_s + tid * ROOT_ENV;
    f3 pos = {r[0], r[1], r[2]};
    f4 q = {r[3], r[4], r[5], r[6]};
    f3 v = {r[7], r[8], r[9]};
    f3 w = {r[10], r[11], r[12]};
    const float bx = r[13], by = r[14];
    f4 bq = {r[16], r[17], r[18], r[19]};
    AntCore o = ant_core<FLAVOR>(pos, q, v, w, f4{c.inv_start_rot[0], c.inv_start_rot[1], c.inv_start_rot[2], c.inv_start_rot[3]});
    float dps[8], dvs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      act[j] = clampf(act[j], -p.clip_actions, p.clip_actions);
      dps[j] = unscale(raw[2 * j], c.dof_lower[j], c.dof_upper[j]);
      dvs[j] = fmul(raw[2 * j + 1], c.dof_vel_scale);
    }
    float* ob = obs_s + tid * OBS_PAD;
    ob[0] = pos.z;
    ob[1] = o.vel_loc.x; ob[2] = o.vel_loc.y; ob[3] = o.vel_loc.z;
    ob[4] = o.angvel_loc.x; ob[5] = o.angvel_loc.y; ob[6] = o.angvel_loc.z;
    ob[7] = o.yaw; ob[8] = o.roll; ob[9] = o.angle_to_target; ob[10] = o.up_proj; ob[11] = o.heading_proj;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ob[12 + j] = dps[j]; ob[20 + j] = dvs[j]; ob[52 + j] = act[j]; }
#pragma unroll
    for (int j = 0; j < 24; ++j) ob[28 + j] = fmul(sens[j], c.contact_force_scale);

    if (p.forces) {
      float* f = p.forces + (int64_t)t * p.forces_frame_stride + (int64_t)e * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmul(fmul(act[j], c.joint_gears[j]), c.power_scale);
    }

    // reward (one_ant.py:498-560)
    const float quat_dist = box_quat_dist(bq, c.x_goal, c.y_goal, c.z_goal);
    const float quat_reward = fmul(c.quat_reward_scale, quat_dist);
    const float d_now = l2_dist2(pos.x, pos.y, bx, by);
    const float push = (d_now < 1.5f) ? 0.0f : 1.0f;
    const float ant_dist = fsub(l2_dist2(pbx, pby, bbx, bby), d_now);
    const float adr = fmul(fmul(c.ant_dist_reward_scale, ant_dist), push);
    const float gdb = l2_dist2(0.0f, 0.0f, bbx, bby);
    const float gd = l2_dist2(0.0f, 0.0f, bx, by);
    const bool arrive = gd < 0.5f;
    const float gdr = fmul(c.goal_dist_reward_scale, fsub(gdb, gd));
    const float up = (o.up_proj > 0.93f) ? fadd(0.0f, c.up_weight) : 0.0f;
    float sq8[8], el8[8];
    int lim = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sq8[j] = fmul(act[j], act[j]);
      el8[j] = fabsf(fmul(act[j], dvs[j]));
      lim += (dps[j] > 0.99f) ? 1 : 0;
    }
    const float actions_cost = sum8<FLAVOR>(sq8);
    const float elec = sum8<FLAVOR>(el8);
    float total = fadd(0.5f, up);
    total = fadd(total, quat_reward);
    total = fadd(total, adr);
    total = fadd(total, gdr);
    total = fadd(total, arrive ? 2.0f : 0.0f);
    total = fadd(total, (quat_dist > 0.9f && arrive) ? 10.0f : 0.0f);
    total = fsub(total, fmul(c.actions_cost, actions_cost));
    total = fsub(total, fmul(c.energy_cost, elec));
    total = fsub(total, fmul((float)lim, c.joints_at_limit_cost));
    const bool fallen = pos.z < c.termination_height;
    if (fallen) total = c.death_cost;
    if (p.rewards) p.rewards[(int64_t)t * p.rewards_frame_stride + e] = total;

    if (t == T - 1) {  // task attributes after the last frame (never read by this launch)
      if (p.up_vec) { p.up_vec[3 * (int64_t)e] = o.up_vec.x; p.up_vec[3 * (int64_t)e + 1] = o.up_vec.y; p.up_vec[3 * (int64_t)e + 2] = o.up_vec.z; }
      if (p.heading_vec) { p.heading_vec[3 * (int64_t)e] = o.heading_vec.x; p.heading_vec[3 * (int64_t)e + 1] = o.heading_vec.y; p.heading_vec[3 * (int64_t)e + 2] = o.heading_vec.z; }
      if (p.ant_pos) { p.ant_pos[2 * (int64_t)e] = pos.x; p.ant_pos[2 * (int64_t)e + 1] = pos.y; }
      if (p.box_pos) { p.box_pos[2 * (int64_t)e] = bx; p.box_pos[2 * (int64_t)e + 1] = by; }
      if (p.box_quat) { float* bqo = p.box_quat + 4 * (int64_t)e; bqo[0] = bq.x; bqo[1] = bq.y; bqo[2] = bq.z; bqo[3] = bq.w; }
    }
    if (T == 1) {
      const float pot_old = p.potentials[e];
      p.prev_potentials[e] = pot_old;                        // one_ant.py:586
      p.potentials[e] = potential_of<FLAVOR>(bx, by, c.dt);  // one_ant.py:589
      p.pos_before[2 * (int64_t)e] = pos.x; p.pos_before[2 * (int64_t)e + 1] = pos.y;
      p.box_before[2 * (int64_t)e] = bx; p.box_before[2 * (int64_t)e + 1] = by;
      int64_t prog = p.progress_buf[e] + 1;
      if (p.reset_buf[e] != 0) prog = 0;
      int64_t rs = fallen ? 1 : 0;
      if ((float)prog >= (float)((double)c.max_episode_length - 1.0)) rs = 1;
      p.progress_buf[e] = prog;
      p.reset_buf[e] = rs;
      if (p.dones_i64) p.dones_i64[e] = rs;
      if (p.dones_u8) p.dones_u8[e] = (uint8_t)rs;
    } else {
      if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] = fallen ? 1 : 0;
      else p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] = fallen ? 1 : 0;
    }
  }
  __syncthreads();

  const float clip = p.clip_obs;
  const int n = ne * OBS_ENV;
  if (p.obs_raw) {
    float* g = p.obs_raw + (int64_t)t * p.obs_raw_frame_stride + (int64_t)e0 * OBS_ENV;
    for (int i = tid; i < n; i += EPT) { int er = i / OBS_ENV; g[i] = obs_s[er * OBS_PAD + (i - er * OBS_ENV)]; }
  }
  if (p.obs) {
    float* g = p.obs + (int64_t)t * p.obs_frame_stride + (int64_t)e0 * OBS_ENV;
    for (int i = tid; i < n; i += EPT) { int er = i / OBS_ENV; g[i] = clampf(obs_s[er * OBS_PAD + (i - er * OBS_ENV)], -clip, clip); }
  }
}

// progress / reset chain and the carry after the last frame for T > 1 launches
template <int FLAVOR>
__global__ void one_ant_chain_kernel(const __grid_constant__ mmb_one_ant_params p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.num_envs) return;
  const int T = p.num_frames;
  int64_t prog = p.progress_buf[e];
  bool flag = p.reset_buf[e] != 0;
  const float thr = (float)((double)p.c.max_episode_length - 1.0);
  for (int t = 0; t < T; ++t) {
    prog = flag ? 0 : prog + 1;
    bool fallen = p.dones_u8 ? (p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] != 0)
                             : (p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] != 0);
    flag = fallen || ((float)prog >= thr);
    if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] = flag ? 1 : 0;
    if (p.dones_i64) p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] = flag ? 1 : 0;
  }
  p.progress_buf[e] = prog;
  p.reset_buf[e] = flag ? 1 : 0;
  const float* r1 = p.root + (int64_t)(T - 1) * p.root_frame_stride + (int64_t)e * ROOT_ENV;
  const float* r0 = p.root + (int64_t)(T - 2) * p.root_frame_stride + (int64_t)e * ROOT_ENV;
  p.pos_before[2 * (int64_t)e] = r1[0]; p.pos_before[2 * (int64_t)e + 1] = r1[1];
  p.box_before[2 * (int64_t)e] = r1[13]; p.box_before[2 * (int64_t)e + 1] = r1[14];
  p.prev_potentials[e] = potential_of<FLAVOR>(r0[13], r0[14], p.c.dt);
  p.potentials[e] = potential_of<FLAVOR>(r1[13], r1[14], p.c.dt);
}

}  // namespace
}  // namespace mmb

extern "C" int32_t mmb_one_ant_step(const mmb_one_ant_params* pp, void* stream) {
  using namespace mmb;
  if (!pp) return MMB_EINVAL;
  mmb_one_ant_params p = *pp;
  if (p.num_envs <= 0 || p.num_frames <= 0) return MMB_EINVAL;
  if (!p.root || !p.dof || !p.sensor || !p.actions || !p.pos_before || !p.box_before || !p.potentials ||
      !p.prev_potentials || !p.progress_buf || !p.reset_buf)
    return MMB_EINVAL;
  if (p.num_frames > 65535) return MMB_EUNSUPPORTED;
  if (p.num_frames > 1 && !p.dones_u8 && !p.dones_i64) return MMB_EINVAL;
  if (p.flavor != MMB_FLAVOR_CUDA && p.flavor != MMB_FLAVOR_CPU) return MMB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((p.num_envs + EPT - 1) / EPT, p.num_frames);
  {
    LaunchScope ls(K_ONE_ANT, st);
    if (p.flavor == MMB_FLAVOR_CUDA) one_ant_kernel<FLAVOR_CUDA><<<grid, EPT, 0, st>>>(p, 2 * sm_count());
    else one_ant_kernel<FLAVOR_CPU><<<grid, EPT, 0, st>>>(p, 2 * sm_count());
  }
  if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  if (p.num_frames > 1) {
    {
      LaunchScope ls(K_ONE_ANT_CHAIN, st);
      if (p.flavor == MMB_FLAVOR_CUDA) one_ant_chain_kernel<FLAVOR_CUDA><<<(p.num_envs + 255) / 256, 256, 0, st>>>(p);
      else one_ant_chain_kernel<FLAVOR_CPU><<<(p.num_envs + 255) / 256, 256, 0, st>>>(p);
    }
    if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  }
  return MMB_OK;
}

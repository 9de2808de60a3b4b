// ten_ant.cu - fused TenAnt env-step kernels for sm_100a.
//
// One CTA = one tile of 16 environments of one frame t (grid = tiles x T): every (tile, t) unit is independent because
// the carry of step t (pos_before / goal_before) is a pure function of frame t-1, so the horizon-batched
// launch exposes T*N*10 ant-sized work items instead of N*10.  The only sequential-in-t state, progress_buf / reset_buf,
// is a 1-byte-per-env-step chain that does not feed obs or reward: inline when T == 1; for T <= 32 one data-carrying
// 64-bit add per (env, frame), executed by the unit of the last frame (finish_env); ten_ant_post_kernel otherwise.
//
// Two kernels compute the same thing bit for bit (tests/test_gpu_ten_ant.py):
//   ten_ant_split_kernel  default: 352 threads in three warp-uniform roles (core / dof / box), see its header below
//   ten_ant_kernel        one thread per ant (MMB_TEN_ANT_VARIANT=mono), the first version, kept as the simple statement
//
// Data movement per unit (all contiguous because the Isaac layout is env-major):
//   root tile   16*143 floats  global -> smem, one 1-D TMA bulk copy + mbarrier (rows are 52 B: not vector-addressable)
//   dof         16 floats per ant, 4 x LDG.128 by the ant's dof thread, consecutive threads -> consecutive 64 B
//   actions      8 floats per ant, 2 x LDG.128
//   obs tile    16*388 floats  smem -> global, one TMA bulk store (plus clamped / per-agent copies when requested)
//   forces       8 floats per ant, 2 x STG.128
//   the inputs of the unit 1.5 x SMs CTAs ahead in launch order are pulled into L2 by cp.async.bulk.prefetch.L2
// Arithmetic: mmb_math.cuh (IEEE round-to-nearest per op in the reference's order).
//
// Replaces: ten_ant.py:886-891 (forces), :712-808 + jit :1304-1393 (observations, box goals),
// :635-710 + jit :988-1301 (reward/reset), :894-926 (progress, carry), vec_task.py:126-131 /
// multi_vec_task.py:94-144 (clamps and per-agent split).
#include <cstdlib>

#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int A = 10;
constexpr int ROOT_ENV = 143;  // 11 rows x 13
constexpr int OBS_ENV = 388;
constexpr int BOX_W = 12;
constexpr int PART_W = 7;  // adr, gdr, up, elec, asq, #joints at limit | arrived << 8, fallen (odd stride: conflict-free)

template <int EPT>
struct TenAntSmem {
  // [ obs tile | root tile (later reused for the per-ant partial terms) | box terms | mbarrier ]
  static constexpr int kObs = EPT * OBS_ENV;
  static constexpr int kRoot = EPT * ROOT_ENV;
  static constexpr int kBox = EPT * BOX_W;
  static constexpr int kPart = EPT * A * PART_W;
  static_assert(kPart <= kRoot, "partials alias the root tile");
  static constexpr int kFloats = kObs + kRoot + kBox + 4;
  static constexpr int kBytes = kFloats * 4;
};

// prefetch distance in CTAs (launch order); MMB_TEN_ANT_PREFETCH overrides (0 = off).  Swept on B200: flat optimum
// between 1 and 2 CTAs per SM ahead; a full resident wave ahead is too early (evicted / competes with demand loads).
inline int prefetch_distance() {
  static const int dist = [] {
    const char* v = getenv("MMB_TEN_ANT_PREFETCH");
    if (v) return atoi(v);
    const int sms = sm_count();
    return sms + sms / 2;
  }();
  return dist;
}

// goal offsets c_j of ten_ant.py:1365-1390 and box_targets_k of ten_ant.py:172-181
__device__ __forceinline__ float goal_offset(int k) { return 1.5f + 3.0f * (float)(k >> 1); }

// compute_box_angle + (sin, -cos) of compute_box_pos (ten_ant.py:935-947, 1353-1364)
__device__ __forceinline__ void box_dir(float qz, float qw, float& s, float& c) {
  float y = fmul(fmul(2.0f, qw), qz);
  float x = fsub(1.0f, fmul(fmul(2.0f, qz), qz));
  float ang = atanf(fdiv(y, x));
  s = sinf(ang);
  c = -cosf(ang);
}

__device__ __forceinline__ void goal_of(int k, float bx, float by, float g0x, float g0y, float& gx, float& gy) {
  float c = goal_offset(k);
  if ((k & 1) == 0) {
    gx = fadd(bx, fmul(c, g0x));
    gy = fadd(by, fmul(c, g0y));
  } else {
    gx = fsub(bx, fmul(c, g0x));
    gy = fsub(by, fmul(c, g0y));
  }
}

// carry of one (env, ant) from a root tensor (ten_ant.py:870-882)
__device__ __forceinline__ void load_carry_one(const float* __restrict__ root, int i, float* pos_before, float* goal_before,
                                               float* box_before) {
  const int e = i / A, k = i - e * A;
  const float* r = root + ((int64_t)e * 11 + k) * 13;
  const float* b = root + ((int64_t)e * 11 + 10) * 13;
  const float px = __ldg(r), py = __ldg(r + 1), bx = __ldg(b), by = __ldg(b + 1), qz = __ldg(b + 5), qw = __ldg(b + 6);
  float s, cs, gx, gy;
  box_dir(qz, qw, s, cs);
  goal_of(k, bx, by, s, cs, gx, gy);
  *reinterpret_cast<float2*>(pos_before + 2 * (int64_t)i) = make_float2(px, py);
  *reinterpret_cast<float2*>(goal_before + 2 * (int64_t)i) = make_float2(gx, gy);
  if (k == 0) *reinterpret_cast<float2*>(box_before + 2 * (int64_t)e) = make_float2(bx, by);
}

// progress / reset chain of one env over frames [t0, t0+cnt) given their `fallen` bits
// (ten_ant.py:896-901 progress += 1 and reset_idx zeroing, :1296-1299 reset rule); writes the final done flags.
__device__ __forceinline__ uint32_t chain_bits(const mmb_ten_ant_params& p, int e, int t0, int cnt, uint32_t fallen_bits,
                                               int64_t& prog, bool& flag) {
  const float thr = (float)((double)p.c.max_episode_length - 1.0);
  uint32_t out_bits = 0;
  for (int j = 0; j < cnt; ++j) {
    prog = flag ? 0 : prog + 1;
    flag = ((fallen_bits >> j) & 1u) || ((float)prog >= thr);
    out_bits |= (flag ? 1u : 0u) << j;
  }
#pragma unroll 8
  for (int j = 0; j < cnt; ++j) {
    const int t = t0 + j;
    const int v = (out_bits >> j) & 1u;
    if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + e] = (uint8_t)v;
    if (p.dones_i64) p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + e] = v;
  }
  return out_bits;
}

constexpr long long CHAIN_SPIN_CYCLES = 2000000000ll;  // ~1 s: then count an error (scratch word N) instead of guessing

// chain of one env with the `fallen` flags read back from the done plane (fallback path: T > 32 or no scratch).
// All flags of a 32-frame chunk are loaded before the dependent chain runs: one memory latency, not T.
__device__ __forceinline__ void chain_one(const mmb_ten_ant_params& p, int e) {
  const int T = p.num_frames;
  int64_t prog = p.progress_buf[e];
  bool flag = p.reset_buf[e] != 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int cnt = min(32, T - t0);
    uint32_t fallen_bits = 0;
#pragma unroll 8
    for (int j = 0; j < cnt; ++j) {
      const int t = t0 + j;
      const bool f = p.dones_u8 ? (__ldcg(p.dones_u8 + (int64_t)t * p.dones_u8_frame_stride + e) != 0)
                                : (__ldcg(p.dones_i64 + (int64_t)t * p.dones_i64_frame_stride + e) != 0);
      fallen_bits |= (f ? 1u : 0u) << j;
    }
    chain_bits(p, e, t0, cnt, fallen_bits, prog, flag);
  }
  p.progress_buf[e] = prog;
  p.reset_buf[e] = flag ? 1 : 0;
}

// smem tile -> global, 128-bit, optional clamp; trip count known at compile time for full tiles
template <int NT, int N4_FULL, bool CLAMP>
__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* __restrict__ s, int n, int tid, float clip) {
  if (aligned16(g)) {
    const int n4 = n >> 2;
    if (n4 == N4_FULL) {
      constexpr int ITERS = (N4_FULL + NT - 1) / NT;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int i = tid + it * NT;
        if (it < ITERS - 1 || i < N4_FULL) {
          float4 v = reinterpret_cast<const float4*>(s)[i];
          if (CLAMP) { v.x = clampf(v.x, -clip, clip); v.y = clampf(v.y, -clip, clip); v.z = clampf(v.z, -clip, clip); v.w = clampf(v.w, -clip, clip); }
          stg4(g + 4 * i, v);
        }
      }
    } else {
      for (int i = tid; i < n4; i += NT) {
        float4 v = reinterpret_cast<const float4*>(s)[i];
        if (CLAMP) { v.x = clampf(v.x, -clip, clip); v.y = clampf(v.y, -clip, clip); v.z = clampf(v.z, -clip, clip); v.w = clampf(v.w, -clip, clip); }
        stg4(g + 4 * i, v);
      }
    }
  } else {
    for (int i = tid; i < n; i += NT) g[i] = CLAMP ? clampf(s[i], -clip, clip) : s[i];
  }
}

// Per-env finish by one thread: ordered sums over the ten ants' partial terms (ten_ant.py:1173-1301), reward, and the
// progress / reset bookkeeping (inline for T == 1, data-carrying atomic + last-reporter chain for T <= 32).
// `root_env`: this env's 143-float root block of frame t still intact in shared memory (role-split kernel), or nullptr.
// ordered sums over the ten ants' partial terms and the reward of one env (ten_ant.py:1173-1301)
__device__ __forceinline__ float env_reward(const mmb_ant_consts& c, const float* pt, const float* bo, bool& fallen_out) {
  float adr = pt[0], gdr = pt[1], up = pt[2], elec = pt[3], asq = pt[4];
  int la = __float_as_int(pt[5]);
  int lim = la & 0xff, n_arrive = la >> 8;
  bool fallen = __float_as_int(pt[6]) != 0;
#pragma unroll
  for (int kk = 1; kk < A; ++kk) {
    const float* qq = pt + kk * PART_W;
    adr = fadd(adr, qq[0]); gdr = fadd(gdr, qq[1]); up = fadd(up, qq[2]); elec = fadd(elec, qq[3]); asq = fadd(asq, qq[4]);
    int l2 = __float_as_int(qq[5]);
    lim += l2 & 0xff; n_arrive += l2 >> 8; fallen = fallen || (__float_as_int(qq[6]) != 0);
  }
  float quat_dist = bo[4];
  float total_r = fadd(5.0f, fmul(up, 10.0f));
  total_r = fadd(total_r, fmul(c.quat_reward_scale, quat_dist));
  total_r = fadd(total_r, adr);
  total_r = fadd(total_r, gdr);
  total_r = fadd(total_r, (float)(2 * n_arrive));
  total_r = fadd(total_r, (quat_dist > 0.9f && n_arrive == A) ? 100.0f : 0.0f);
  total_r = fsub(total_r, fmul(c.actions_cost, asq));
  total_r = fsub(total_r, fmul(c.energy_cost, elec));
  total_r = fsub(total_r, fmul((float)lim, c.joints_at_limit_cost));
  if (fallen) total_r = c.death_cost;
  fallen_out = fallen;
  return total_r;
}

// carry of one env from the last frame's tile in shared memory (the executor is the frame T-1 unit: ant xy, box xy and
// the goal direction of the last frame are there already; bo[0..3] was computed from the same box row, same operations)
__device__ __forceinline__ void carry_from_tile(const mmb_ten_ant_params& p, int en, const float* bo, const float* root_env) {
#pragma unroll
  for (int kk = 0; kk < A; ++kk) {
    float gx, gy;
    goal_of(kk, bo[2], bo[3], bo[0], bo[1], gx, gy);
    *reinterpret_cast<float2*>(p.pos_before + ((int64_t)en * A + kk) * 2) = make_float2(root_env[kk * 13], root_env[kk * 13 + 1]);
    *reinterpret_cast<float2*>(p.goal_before + ((int64_t)en * A + kk) * 2) = make_float2(gx, gy);
  }
  *reinterpret_cast<float2*>(p.box_before + (int64_t)en * 2) = make_float2(bo[2], bo[3]);
}

__device__ __forceinline__ void finish_env(const mmb_ten_ant_params& p, int t, int en, const float* pt, const float* bo,
                                           const float* root_env = nullptr) {
  const mmb_ant_consts& c = p.c;
  const int T = p.num_frames;
  bool fallen;
  const float total_r = env_reward(c, pt, bo, fallen);
  if (p.rewards) p.rewards[(int64_t)t * p.rewards_frame_stride + en] = total_r;
  if (T == 1) {
    // carry out (ten_ant.py:905-925).  Role-split kernel: the whole env's carry from the tiles, here - behind barrier B3 -
    // because the dof thread of an ant reads goal_before (frame 0 without prev_root) after B2: the core thread's write,
    // which used to follow its own read at once, raced with that read (the reward's goal-distance term saw the NEW goal
    // whenever the dof warp left griddepcontrol.wait late).  One-thread-per-ant kernel: written by the ant's own thread.
    if (root_env) {
      carry_from_tile(p, en, bo, root_env);
    } else {
      p.box_before[(int64_t)en * 2] = bo[2];
      p.box_before[(int64_t)en * 2 + 1] = bo[3];
    }
    // ten_ant.py:896-901 (progress += 1; reset_idx zeroes progress/reset of flagged envs) + :1296-1299
    int64_t prog = p.progress_buf[en] + 1;
    if (p.reset_buf[en] != 0) prog = 0;
    int64_t rs = fallen ? 1 : 0;
    if ((float)prog >= (float)((double)c.max_episode_length - 1.0)) rs = 1;
    p.progress_buf[en] = prog;
    p.reset_buf[en] = rs;
    if (p.dones_i64) p.dones_i64[en] = rs;
    if (p.dones_u8) p.dones_u8[en] = (uint8_t)rs;
  } else if (p.scratch && T <= 32) {
    // Horizon-batched launch: ONE data-carrying 64-bit add per (env, frame).  The word of env `en` collects the `fallen`
    // bit of every frame (bits 0..31) and the number of frames that have reported (bits 32..).  Frames 0..T-2 use a
    // fire-and-forget reduction: nothing comes back, so their CTAs retire without waiting for an L2 round trip (a
    // returning atomic took ~2 us under load and was the tail of every CTA).  The unit of the LAST frame - launched last,
    // so the others have normally reported long before - adds its own report, waits until all T are in, runs the
    // progress / reset chain and writes the carry.  No fence, no flag read-back, no second kernel; the frame-0 unit
    // reports after its carry reads, so the carry is not rewritten under it.
    const unsigned long long mine = (1ull << 32) | ((unsigned long long)(fallen ? 1u : 0u) << t);
    unsigned long long* word = reinterpret_cast<unsigned long long*>(p.scratch) + en;
    if (t != T - 1) {
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(word), "l"(mine) : "memory");
    } else {
      unsigned long long cur = atomicAdd(word, mine) + mine;
      for (long long t0 = clock64(); (unsigned)(cur >> 32) != (unsigned)T && clock64() - t0 < CHAIN_SPIN_CYCLES;) {
        __nanosleep(100);                // earlier-launched units still in flight (forward progress as in a look-back scan)
        cur = *reinterpret_cast<volatile unsigned long long*>(word);
      }
      if ((unsigned)(cur >> 32) != (unsigned)T) {   // reports missing after ~1 s: flag it, guess nothing (mmb.h, `scratch`)
        atomicAdd(reinterpret_cast<unsigned long long*>(p.scratch) + p.num_envs, 1ull);
        return;
      }
      *word = 0ull;                      // self-resetting for the next launch / graph replay
      int64_t prog = p.progress_buf[en];
      bool flag = p.reset_buf[en] != 0;
      chain_bits(p, en, 0, T, (uint32_t)cur, prog, flag);
      p.progress_buf[en] = prog;
      p.reset_buf[en] = flag ? 1 : 0;
      if (root_env) {
        carry_from_tile(p, en, bo, root_env);
        return;
      }
      const float* last = p.root + (int64_t)(T - 1) * p.root_frame_stride;
      // carry of the whole env by this thread: the goal direction once, then ten (xy, goal) pairs
      const float* b = last + ((int64_t)en * 11 + 10) * 13;
      const float bx = __ldg(b), by = __ldg(b + 1);
      float xy[2 * A];
#pragma unroll
      for (int kk = 0; kk < A; ++kk) {
        xy[2 * kk] = __ldg(last + ((int64_t)en * 11 + kk) * 13);
        xy[2 * kk + 1] = __ldg(last + ((int64_t)en * 11 + kk) * 13 + 1);
      }
      float s, cs;
      box_dir(__ldg(b + 5), __ldg(b + 6), s, cs);
#pragma unroll
      for (int kk = 0; kk < A; ++kk) {
        float gx, gy;
        goal_of(kk, bx, by, s, cs, gx, gy);
        *reinterpret_cast<float2*>(p.pos_before + ((int64_t)en * A + kk) * 2) = make_float2(xy[2 * kk], xy[2 * kk + 1]);
        *reinterpret_cast<float2*>(p.goal_before + ((int64_t)en * A + kk) * 2) = make_float2(gx, gy);
      }
      *reinterpret_cast<float2*>(p.box_before + (int64_t)en * 2) = make_float2(bx, by);
    }
  } else {  // `fallen` only; ten_ant_post_kernel finishes the flags
    if (p.dones_u8) p.dones_u8[(int64_t)t * p.dones_u8_frame_stride + en] = fallen ? 1 : 0;
    else p.dones_i64[(int64_t)t * p.dones_i64_frame_stride + en] = fallen ? 1 : 0;
  }
}

// Outputs that are not a verbatim copy of the obs tile: clamped copies of a raw tile and the per-agent [N][10][46] view.
template <int NT, int EPT>
__device__ __forceinline__ void extra_outputs(const mmb_ten_ant_params& p, int t, int e0, int ne, int tid, const float* obs_s,
                                              bool tile_clamped, float clip) {
  const int n = ne * OBS_ENV;
  if (!tile_clamped) {  // raw tile: clamped outputs need a pass
    if (p.obs_layout == 0 && p.obs) {
      tile_store<NT, EPT * OBS_ENV / 4, true>(p.obs + (int64_t)t * p.obs_frame_stride + (int64_t)e0 * OBS_ENV, obs_s, n, tid, clip);
    } else if (p.obs_layout >= 1 && p.share_obs) {
      tile_store<NT, EPT * OBS_ENV / 4, true>(p.share_obs + (int64_t)t * p.share_obs_frame_stride + (int64_t)e0 * OBS_ENV, obs_s, n, tid, clip);
    }
  }
  if (p.obs_layout == 1 && p.obs) {  // multi_vec_task.py:105-116: per agent cat(own 38, tail 8) -> [N][10][46]
    float* g = p.obs + (int64_t)t * p.obs_frame_stride + (int64_t)e0 * 460;
    const int n2 = ne * 230;  // float2 granules: 46 and 38 are even, so a pair never straddles a boundary
    const bool al8 = (reinterpret_cast<uintptr_t>(g) & 7u) == 0;
    for (int i = tid; i < n2; i += NT) {
      int er = i / 230, r2 = i - er * 230;
      int a = r2 / 23, c2 = r2 - a * 23;
      int src = er * OBS_ENV + (c2 < 19 ? a * 38 + 2 * c2 : 380 + 2 * (c2 - 19));
      float2 v = *reinterpret_cast<const float2*>(obs_s + src);
      if (!tile_clamped) { v.x = clampf(v.x, -clip, clip); v.y = clampf(v.y, -clip, clip); }
      if (al8) *reinterpret_cast<float2*>(g + 2 * i) = v;
      else { g[2 * i] = v.x; g[2 * i + 1] = v.y; }
    }
  } else if (p.obs_layout == 2 && p.obs) {  // the same rows, agent-major: plane a holds [env][46] (shared MARL buffer slot)
    const int n2 = ne * 23;  // float2 granules of one agent's rows of this tile
    for (int a = 0; a < A; ++a) {
      float* g = p.obs + (int64_t)a * p.obs_agent_stride + (int64_t)t * p.obs_frame_stride + (int64_t)e0 * 46;
      const bool al8 = (reinterpret_cast<uintptr_t>(g) & 7u) == 0;
      for (int i = tid; i < n2; i += NT) {
        int er = i / 23, c2 = i - er * 23;
        int src = er * OBS_ENV + (c2 < 19 ? a * 38 + 2 * c2 : 380 + 2 * (c2 - 19));
        float2 v = *reinterpret_cast<const float2*>(obs_s + src);
        if (!tile_clamped) { v.x = clampf(v.x, -clip, clip); v.y = clampf(v.y, -clip, clip); }
        if (al8) *reinterpret_cast<float2*>(g + 2 * i) = v;
        else { g[2 * i] = v.x; g[2 * i + 1] = v.y; }
      }
    }
  }
}

template <int EPT>
struct UnitIdx {
  int tile, t, e0, ne;
  __device__ __forceinline__ UnitIdx(int u, int ntiles, int N) {
    t = u / ntiles;              // env tile fastest: neighbouring CTAs stream neighbouring memory of one frame
    tile = u - t * ntiles;
    e0 = tile * EPT;
    ne = min(EPT, N - e0);
  }
};

// L2 prefetch of the inputs of unit `u` (launch order: env tile fastest, then frame) by the TMA engine.  A CTA calls it
// for the unit about a third of a resident wave ahead: by the time that CTA starts, its root / dof / action tiles sit
// in L2, so its load phase sees L2 latency instead of a loaded-HBM round trip and HBM requests are issued early.
// Three instructions in one thread, no registers or shared memory held.  Measured: -9 % kernel time.
template <int EPT>
__device__ __forceinline__ void prefetch_unit(const mmb_ten_ant_params& p, int64_t u, int ntiles) {
  const int64_t t2 = u / ntiles, tile2 = u - t2 * ntiles;
  if (t2 >= p.num_frames || (tile2 + 1) * EPT > p.num_envs) return;
  const float* r2 = p.root + t2 * p.root_frame_stride + tile2 * EPT * ROOT_ENV;
  const float* d2 = p.dof + t2 * p.dof_frame_stride + tile2 * EPT * 160;
  if (!p.actions) {          // per-agent action tensors (T == 1): ten short rows per env, not worth a bulk prefetch
    if (aligned16(r2) && aligned16(d2)) { tma_prefetch_l2(r2, EPT * ROOT_ENV * 4); tma_prefetch_l2(d2, EPT * 160 * 4); }
    return;
  }
  const float* a2 = p.actions + t2 * p.actions_frame_stride + tile2 * EPT * 80;
  if (aligned16(r2) && aligned16(d2) && aligned16(a2)) {
    tma_prefetch_l2(r2, EPT * ROOT_ENV * 4);
    tma_prefetch_l2(d2, EPT * 160 * 4);
    tma_prefetch_l2(a2, EPT * 80 * 4);
  }
}

// One CTA = one unit = one tile of EPT envs of one frame.  Short-lived CTAs at 64 registers keep 6 (EPT 16) or 3
// (EPT 32) CTAs = 30 warps per SM resident: the kernel is a long dependent fp32 chain per thread, so it lives on
// thread-level parallelism (a persistent, register-prefetching variant at 96 registers measured 40% slower).
template <int FLAVOR, int EPT>
__global__ void __launch_bounds__(EPT* A, (EPT == 32) ? 3 : 6) ten_ant_kernel(const __grid_constant__ mmb_ten_ant_params p,
                                                                              const int prefetch_dist) {
  constexpr int NT = EPT * A;
  extern __shared__ __align__(128) float smem[];
  float* obs_s = smem;
  float* root_s = obs_s + TenAntSmem<EPT>::kObs;
  float* part_s = root_s;  // aliased: every root read happens before the barrier that precedes the first partial write
  float* box_s = root_s + TenAntSmem<EPT>::kRoot;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(box_s + TenAntSmem<EPT>::kBox);

  const int tid = threadIdx.x;
  const int wid = tid >> 5, lane = tid & 31;
  const int N = p.num_envs;
  const int T = p.num_frames;
  const UnitIdx<EPT> ui(blockIdx.x, (N + EPT - 1) / EPT, N);
  const int t = ui.t, e0 = ui.e0, ne = ui.ne;
  const mmb_ant_consts& c = p.c;

  // ---- root tile: one 1-D TMA bulk copy for a full, 16B-aligned tile; cooperative loads otherwise ----
  const float* root_g = p.root + (int64_t)t * p.root_frame_stride + (int64_t)e0 * ROOT_ENV;
  const bool use_tma = (ne == EPT) && aligned16(root_g);
  if (use_tma) {
    if (tid == 0) {
      mbar_init(mbar, 1);
      mbar_expect_tx(mbar, EPT * ROOT_ENV * 4);
      tma_load_1d(root_s, root_g, EPT * ROOT_ENV * 4, mbar);
    }
  } else {
    tile_load(root_s, root_g, ne * ROOT_ENV, tid, NT);
  }

  if (prefetch_dist > 0 && tid == 32) prefetch_unit<EPT>(p, (int64_t)blockIdx.x + prefetch_dist, (N + EPT - 1) / EPT);

  const int el = tid / A, k = tid - el * A;
  const int e = e0 + el;
  const bool active = el < ne;
  // the tile holds clamped values unless the unclamped task.obs_buf is requested as well
  const bool tile_clamped = (p.obs_raw == nullptr);
  const float clip = p.clip_obs;
  const float tclip = tile_clamped ? clip : __int_as_float(0x7f800000);

  float dps[8], dvs[8], act[8];
  float pbx = 0.f, pby = 0.f, gbx = 0.f, gby = 0.f;
  float pbq0 = 0.f, pbq1 = 0.f, pbq2 = 0.f, pbq3 = 1.f;  // frame t-1 box row (x, y, qz, qw), warp 1 lanes only
  const bool prev_box = (wid == 1) && (lane < ne) && (t > 0);
  if (prev_box) {  // fetched now so that its latency overlaps the wait for the root tile
    const float* b = p.root + (int64_t)(t - 1) * p.root_frame_stride + ((int64_t)(e0 + lane) * 11 + 10) * 13;
    pbq0 = __ldg(b); pbq1 = __ldg(b + 1); pbq2 = __ldg(b + 5); pbq3 = __ldg(b + 6);
  }
  if (active) {
    const float* d = p.dof + (int64_t)t * p.dof_frame_stride + ((int64_t)e * 80 + 8 * k) * 2;
    const float* a = p.actions + (int64_t)t * p.actions_frame_stride + (int64_t)e * 80 + 8 * k;
    float raw[16];
    if (aligned16(d) && aligned16(a)) {
      float4 v[4], w0, w1;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ldg4(d + 4 * j);
      w0 = ldg4(a);
      w1 = ldg4(a + 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) { raw[4 * j] = v[j].x; raw[4 * j + 1] = v[j].y; raw[4 * j + 2] = v[j].z; raw[4 * j + 3] = v[j].w; }
      act[0] = w0.x; act[1] = w0.y; act[2] = w0.z; act[3] = w0.w;
      act[4] = w1.x; act[5] = w1.y; act[6] = w1.z; act[7] = w1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) raw[j] = __ldg(d + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) act[j] = __ldg(a + j);
    }
    if (t == 0) {
      const float* pb = p.pos_before + ((int64_t)e * A + k) * 2;
      const float* gb = p.goal_before + ((int64_t)e * A + k) * 2;
      pbx = __ldg(pb); pby = __ldg(pb + 1);
      gbx = __ldg(gb); gby = __ldg(gb + 1);
    } else {  // carry of step t = ant xy of frame t-1 (ten_ant.py:905-914)
      const float* rp = p.root + (int64_t)(t - 1) * p.root_frame_stride + ((int64_t)e * 11 + k) * 13;
      pbx = __ldg(rp); pby = __ldg(rp + 1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      act[j] = clampf(act[j], -p.clip_actions, p.clip_actions);     // vec_task.py:127
      dps[j] = unscale(raw[2 * j], c.dof_lower[j], c.dof_upper[j]);  // ten_ant.py:1333
      dvs[j] = fmul(raw[2 * j + 1], c.dof_vel_scale);                // ten_ant.py:1347
    }
    if (p.forces) {  // ten_ant.py:889: actions * joint_gears * power_scale
      float* f = p.forces + (int64_t)t * p.forces_frame_stride + (int64_t)e * 80 + 8 * k;
      float fo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) fo[j] = fmul(fmul(act[j], c.joint_gears[j]), c.power_scale);
      if (aligned16(f)) {
        stg4(f, make_float4(fo[0], fo[1], fo[2], fo[3]));
        stg4(f + 4, make_float4(fo[4], fo[5], fo[6], fo[7]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fo[j];
      }
    }
  }
  __syncthreads();                       // mbarrier initialised (TMA path) / tile stores visible (fallback path)
  if (use_tma) mbar_wait(mbar, 0);

  // ---- every ant thread lifts its root row into registers; warp 0 / warp 1 derive the goal directions of
  // frame t / frame t-1 (separate warps: no divergence between the two roles) ----
  f3 pos = {0.f, 0.f, 0.f}, vel = pos, ang = pos;
  f4 q = {0.f, 0.f, 0.f, 1.f};
  if (active) {
    const float* r = root_s + el * ROOT_ENV + k * 13;
    pos = f3{r[0], r[1], r[2]};
    q = f4{r[3], r[4], r[5], r[6]};
    vel = f3{r[7], r[8], r[9]};
    ang = f3{r[10], r[11], r[12]};
  }
  if (wid == 0 && lane < ne) {
    const float* b = root_s + lane * ROOT_ENV + 10 * 13;
    float s, cs;
    box_dir(b[5], b[6], s, cs);
    float* bo = box_s + lane * BOX_W;
    bo[0] = s; bo[1] = cs; bo[2] = b[0]; bo[3] = b[1];
    bo[4] = box_quat_dist(f4{b[3], b[4], b[5], b[6]}, c.x_goal, c.y_goal, c.z_goal);
    float* tail = obs_s + lane * OBS_ENV + 380;  // ten_ant.py:806-808: box_pos, box_quat, box_targets(=0)
    tail[0] = clampf(b[0], -tclip, tclip); tail[1] = clampf(b[1], -tclip, tclip);
    tail[2] = clampf(b[3], -tclip, tclip); tail[3] = clampf(b[4], -tclip, tclip);
    tail[4] = clampf(b[5], -tclip, tclip); tail[5] = clampf(b[6], -tclip, tclip);
    tail[6] = 0.0f; tail[7] = 0.0f;
  } else if (prev_box) {
    float s, cs;
    box_dir(pbq2, pbq3, s, cs);
    float* bo = box_s + lane * BOX_W;
    bo[8] = s; bo[9] = cs; bo[10] = pbq0; bo[11] = pbq1;
  }
  __syncthreads();                       // box terms ready; all reads of the root tile are done (part_s may overwrite it)

  // ---- ant phase ----
  if (active) {
    AntCore o = ant_core<FLAVOR>(pos, q, vel, ang, f4{c.inv_start_rot[0], c.inv_start_rot[1], c.inv_start_rot[2], c.inv_start_rot[3]});
    float* ob = obs_s + el * OBS_ENV + k * 38;
    ob[0] = clampf(pos.x, -tclip, tclip); ob[1] = clampf(pos.y, -tclip, tclip); ob[2] = clampf(pos.z, -tclip, tclip);
    ob[3] = clampf(o.vel_loc.x, -tclip, tclip); ob[4] = clampf(o.vel_loc.y, -tclip, tclip); ob[5] = clampf(o.vel_loc.z, -tclip, tclip);
    ob[6] = clampf(o.angvel_loc.x, -tclip, tclip); ob[7] = clampf(o.angvel_loc.y, -tclip, tclip); ob[8] = clampf(o.angvel_loc.z, -tclip, tclip);
    ob[9] = clampf(o.yaw, -tclip, tclip); ob[10] = clampf(o.roll, -tclip, tclip); ob[11] = clampf(o.angle_to_target, -tclip, tclip);
    ob[12] = clampf(o.up_proj, -tclip, tclip); ob[13] = clampf(o.heading_proj, -tclip, tclip);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ob[14 + j] = clampf(dps[j], -tclip, tclip);
      ob[22 + j] = clampf(dvs[j], -tclip, tclip);
      ob[30 + j] = clampf(act[j], -tclip, tclip);
    }

    const float* bo = box_s + el * BOX_W;
    float gx, gy;
    goal_of(k, bo[2], bo[3], bo[0], bo[1], gx, gy);
    if (t > 0) goal_of(k, bo[10], bo[11], bo[8], bo[9], gbx, gby);

    // ten_ant.py:1073-1081 for ant k
    float d_now = l2_dist2(pos.x, pos.y, gx, gy);
    float push = (d_now < 1.5f) ? 0.0f : 1.0f;
    float ant_dist = fsub(l2_dist2(pbx, pby, gbx, gby), d_now);
    float adr = fmul(fmul(c.ant_dist_reward_scale, ant_dist), push);
    float bty = (k & 1) ? goal_offset(k) : -goal_offset(k);
    float gdb = l2_dist2(0.0f, bty, gbx, gby);
    float gd = l2_dist2(0.0f, bty, gx, gy);
    bool arrive = gd < 0.5f;
    float gdr = fmul(c.goal_dist_reward_scale, fsub(gdb, gd));
    float up = (o.up_proj > 0.93f) ? fadd(0.0f, c.up_weight) : 0.0f;  // ten_ant.py:1187
    float el8[8];
    int lim = 0;
    float asq = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      el8[j] = fabsf(fmul(act[j], dvs[j]));  // ten_ant.py:1242
      lim += (dps[j] > 0.99f) ? 1 : 0;       // ten_ant.py:1243
      asq = fadd(asq, fmul(act[j], act[j]));
    }
    float elec = sum8<FLAVOR>(el8);
    bool fallen = pos.z < c.termination_height;
    float* pt = part_s + (el * A + k) * PART_W;
    pt[0] = adr; pt[1] = gdr; pt[2] = up; pt[3] = elec; pt[4] = asq;
    pt[5] = __int_as_float(lim | (arrive ? 0x100 : 0));
    pt[6] = __int_as_float(fallen ? 1 : 0);

    if (T == 1) {  // carry out (ten_ant.py:905-925); T > 1: written once per env by the chain executor below
      float* pb = p.pos_before + ((int64_t)e * A + k) * 2;
      float* gb = p.goal_before + ((int64_t)e * A + k) * 2;
      pb[0] = pos.x; pb[1] = pos.y; gb[0] = gx; gb[1] = gy;
    }
  }
  fence_async_smem();                    // obs tile writes -> visible to the TMA store engine
  __syncthreads();

  // ---- obs tile out ----
  const int n = ne * OBS_ENV;
  float* obs_dst = nullptr;              // destination served by the TMA bulk store (a verbatim copy of the tile)
  if (tile_clamped) obs_dst = (p.obs_layout == 0) ? p.obs : p.share_obs;
  else obs_dst = p.obs_raw;
  const int64_t dst_stride = tile_clamped ? ((p.obs_layout == 0) ? p.obs_frame_stride : p.share_obs_frame_stride)
                                          : p.obs_raw_frame_stride;
  bool tma_stored = false;
  if (obs_dst) {
    float* g = obs_dst + (int64_t)t * dst_stride + (int64_t)e0 * OBS_ENV;
    if (aligned16(g)) {
      if (tid == 0) tma_store_1d(g, obs_s, (uint32_t)n * 4u);
      tma_stored = true;
    } else {
      tile_store<NT, EPT * OBS_ENV / 4, false>(g, obs_s, n, tid, clip);
    }
  }

  // ---- per-env finish: ordered sums over the ten ants (ten_ant.py:1173-1301) ----
  if (tid < ne) finish_env(p, t, e0 + tid, part_s + tid * A * PART_W, box_s + tid * BOX_W);
  extra_outputs<NT, EPT>(p, t, e0, ne, tid, obs_s, tile_clamped, clip);
  if (tma_stored && tid == 0) tma_store_wait_read();  // the tile must stay intact until the bulk store has read it
}

// ------------------------------------------------------------------------------------------------------
// Role-split variant (default).  The per-ant work is a ~1,250-instruction dependent fp32 chain, so the one-thread-per-ant
// kernel is latency-bound, not DRAM- or issue-bound (ncu: 52 % issue utilisation at 30 resident warps per SM).  Here the
// work of a tile is spread over three warp-uniform roles, 40 registers per thread, 4 CTAs = 44 warps per SM:
//   core warps 0-4  (one thread per ant): root row -> ant_core -> obs[0:14]; after the box barrier the ant-distance term
//   dof  warps 5-9  (one thread per ant): dof / action rows -> unscale, clamps, forces, energy terms -> obs[14:38]; after
//                   the box barrier the goal-distance term; warp 7 also the box-orientation term and the obs tail;
//                   16 lanes of warp 5 do the ordered 10-ant sums, the reward and the chain report (finish_env)
//   box  warp 10    goal direction of frame t (lanes 0-15) and of frame t-1 (lanes 16-31): the two ~150-instruction
//                   fdiv -> atanf -> sinf / cosf chains side by side, fed straight from L2
// What the per-CTA %globaltimer timeline (MMB_TRACE) showed on the way here is in DESIGN.md section 4.
// CTA = 16 envs x 1 frame = 352 threads, grid = (tiles, T).
// ------------------------------------------------------------------------------------------------------
// Optional per-CTA phase timeline (build with `make EXTRA=-DMMB_TRACE`, read with tools/probe/ten_ant_timeline.py): %globaltimer
// stamps of a few CTAs of frame 5.  Compiled out of the normal library.
#ifdef MMB_TRACE
}  // namespace
__device__ unsigned long long g_trace[64 * 16];
namespace {
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define MMB_TR(i) do { if ((blockIdx.x % 64) == 0 && blockIdx.y == 5) g_trace[(blockIdx.x / 64) * 16 + (i)] = gtime(); } while (0)
#else
#define MMB_TR(i) do { } while (0)
#endif

struct SplitSmem {
  static constexpr int EPT = 16;
  static constexpr int kObs = EPT * OBS_ENV;
  static constexpr int kRoot = EPT * ROOT_ENV;
  static constexpr int kPart = EPT * A * PART_W;
  static constexpr int kBox = EPT * BOX_W;
  // fused GAE (executor units only): values [EPT][33] (T <= 32 frames + bootstrap) and rewards [EPT][33] (odd stride:
  // conflict-free) ALIAS the per-ant partial terms, which are dead once the env rewards are in registers (any extra shared
  // memory costs every CTA of the launch L1 capacity: +4 KB per CTA measured +4 % kernel time); the fallen bits and the
  // error flag sit in unused words of the box rows
  static constexpr int GAE_LD = 33;
  static_assert(2 * EPT * GAE_LD <= kPart, "GAE staging aliases the partial terms");
  static constexpr int kFloats = kObs + kRoot + kPart + kBox + 4;
  static constexpr int kBytes = kFloats * 4;
};

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifndef MMB_SPLIT_MIN_CTAS
#define MMB_SPLIT_MIN_CTAS 4
#endif

template <int FLAVOR>
__global__ void __launch_bounds__(352, MMB_SPLIT_MIN_CTAS) ten_ant_split_kernel(const __grid_constant__ mmb_ten_ant_params p,
                                                                                   const int prefetch_dist) {
  constexpr int EPT = SplitSmem::EPT, NA = EPT * A, NT = 2 * NA + 32;   // core warps 0-4, dof warps 5-9, box warp 10
  extern __shared__ __align__(128) float smem[];
  float* obs_s = smem;
  float* root_s = obs_s + SplitSmem::kObs;
  float* part_s = root_s + SplitSmem::kRoot;
  float* box_s = part_s + SplitSmem::kPart;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(box_s + SplitSmem::kBox);
  constexpr int GLD = SplitSmem::GAE_LD;
  float* vals_s = part_s;                              // executor units, after the env rewards have left part_s
  float* rew_s = part_s + SplitSmem::EPT * GLD;

  const int tid = threadIdx.x;
  if (tid == 0) MMB_TR(0);
  const int wid = tid >> 5, lane = tid & 31;
  const int N = p.num_envs;
  // (launching the frames as 1..T-1, 0 so that the frame-0 units never wait on the previous kernel was measured: the
  // frame t-1 carry reads of frame 1 then miss L2 and the kernel is 4 % slower)
  const int t = blockIdx.y, e0 = blockIdx.x * EPT;
  const int ne = min(EPT, N - e0);
  const mmb_ant_consts& c = p.c;
  const bool pdl = p.overlap_prev != 0;
  if (pdl) griddep_launch_dependents();  // the next kernel in the stream may start filling SM slots as this one drains
  // the frame the carry of step t comes from (ten_ant.py:870-882,905-914): frame t - 1 of this launch, or - frame 0 - the
  // frame that preceded the launch (mmb.h, prev_root).  Without prev_root frame 0 reads the carry arrays the previous
  // launch's executors wrote, i.e. waits for that whole kernel.
  const bool has_prev = t > 0 || p.prev_root != nullptr;
  const float* prev_frame = t > 0 ? p.root + (int64_t)(t - 1) * p.root_frame_stride : p.prev_root;
  const bool box_role = tid >= 2 * NA;   // warp 10
  const bool dof_role = tid >= NA && !box_role;   // warp-uniform (NA = 5 warps)
  const int a = box_role ? 0 : (dof_role ? tid - NA : tid);
  const int el = a / A, k = a - el * A;
  const int e = e0 + el;
  const bool active = !box_role && el < ne;
  const bool tile_clamped = (p.obs_raw == nullptr);
  const float clip = p.clip_obs;
  const float tclip = tile_clamped ? clip : __int_as_float(0x7f800000);

  const float* root_g = p.root + (int64_t)t * p.root_frame_stride + (int64_t)e0 * ROOT_ENV;
  const bool use_tma = (ne == EPT) && aligned16(root_g);
  if (use_tma) {
    if (tid == 0) {
      mbar_init(mbar, 1);
      mbar_expect_tx(mbar, EPT * ROOT_ENV * 4);
      tma_load_1d(root_s, root_g, EPT * ROOT_ENV * 4, mbar);
    }
  } else {
    tile_load(root_s, root_g, ne * ROOT_ENV, tid, NT);
  }

  if (prefetch_dist > 0 && tid == NA)
    prefetch_unit<EPT>(p, (int64_t)blockIdx.y * gridDim.x + blockIdx.x + prefetch_dist, (int)gridDim.x);

  // fused GAE (mmb.h, gae_*): the unit of the LAST frame is the executor of its 16 envs
  const bool gae_on = p.gae_values != nullptr;
  const bool gae_exec = gae_on && t == p.num_frames - 1;

  if (box_role) {
    // ================= box warp: goal direction of frame t (lanes 0-15) and of frame t-1 (lanes 16-31) =================
    // both chains (division, atanf, sinf, cosf: ~150 dependent instructions) run side by side in one warp that has no
    // other work, straight from global memory (L2-prefetched), so nobody reaches the box barrier late because of them
    const int env_l = lane & 15;
    const bool prev = lane >= 16;
    const bool on = env_l < ne && !(prev && !has_prev);
    float b0 = 0.f, b1 = 0.f, b5 = 0.f, b6 = 1.f;
    if (on) {
      const float* fr = prev ? prev_frame : p.root + (int64_t)t * p.root_frame_stride;
      const float* b = fr + ((int64_t)(e0 + env_l) * 11 + 10) * 13;
      b0 = __ldg(b); b1 = __ldg(b + 1); b5 = __ldg(b + 5); b6 = __ldg(b + 6);
    }
    __syncthreads();                     // B1
    if (on) {
      float sn, cs;
      box_dir(b5, b6, sn, cs);
      float* bo = box_s + env_l * BOX_W + (prev ? 8 : 0);
      bo[0] = sn; bo[1] = cs; bo[2] = b0; bo[3] = b1;
    }
    __syncthreads();                     // B2
  } else if (dof_role) {
    // ================= dof role =================
    float raw[16], act[8];
    float gbx = 0.f, gby = 0.f;
    if (active) {
      const float* d = p.dof + (int64_t)t * p.dof_frame_stride + ((int64_t)e * 80 + 8 * k) * 2;
      // actions: one [N][80] tensor, or (per-step multi-agent path) the ten per-agent [N][8] tensors as they come from the
      // policies - the hstack of multi_vec_task.py:94-103 never materialises
      const float* ac = p.agent_actions[0] ? p.agent_actions[k] + (int64_t)e * 8
                                           : p.actions + (int64_t)t * p.actions_frame_stride + (int64_t)e * 80 + 8 * k;
      if (aligned16(d) && aligned16(ac)) {
        float4 v[4], w0, w1;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = ldg4(d + 4 * j);
        w0 = ldg4(ac);
        w1 = ldg4(ac + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { raw[4 * j] = v[j].x; raw[4 * j + 1] = v[j].y; raw[4 * j + 2] = v[j].z; raw[4 * j + 3] = v[j].w; }
        act[0] = w0.x; act[1] = w0.y; act[2] = w0.z; act[3] = w0.w;
        act[4] = w1.x; act[5] = w1.y; act[6] = w1.z; act[7] = w1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = __ldg(d + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) act[j] = __ldg(ac + j);
      }
    }
    __syncthreads();                     // B1: mbarrier initialised (TMA path) / tile stores visible (fallback path)
    if (tid == 200) MMB_TR(1);
    int lim = 0;
    if (active) {
      float* ob = obs_s + el * OBS_ENV + k * 38;   // 152-byte rows: 8-byte aligned, so pairs go out as STS.64 (conflict-free)
      float el8[8];
      float asq = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float dp[2], dv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float aj = clampf(act[j + h], -p.clip_actions, p.clip_actions);            // vec_task.py:127
          dp[h] = unscale(raw[2 * (j + h)], c.dof_lower[j + h], c.dof_upper[j + h]);       // ten_ant.py:1333
          dv[h] = fmul(raw[2 * (j + h) + 1], c.dof_vel_scale);                             // ten_ant.py:1347
          act[j + h] = aj;
          el8[j + h] = fabsf(fmul(aj, dv[h]));   // ten_ant.py:1242
          lim += (dp[h] > 0.99f) ? 1 : 0;        // ten_ant.py:1243
          asq = fadd(asq, fmul(aj, aj));
        }
        *reinterpret_cast<float2*>(ob + 14 + j) = make_float2(clampf(dp[0], -tclip, tclip), clampf(dp[1], -tclip, tclip));
        *reinterpret_cast<float2*>(ob + 22 + j) = make_float2(clampf(dv[0], -tclip, tclip), clampf(dv[1], -tclip, tclip));
        *reinterpret_cast<float2*>(ob + 30 + j) = make_float2(clampf(act[j], -tclip, tclip), clampf(act[j + 1], -tclip, tclip));
      }
      if (p.forces) {  // ten_ant.py:889: actions * joint_gears * power_scale
        float* f = p.forces + (int64_t)t * p.forces_frame_stride + (int64_t)e * 80 + 8 * k;
        float fo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) fo[j] = fmul(fmul(act[j], c.joint_gears[j]), c.power_scale);
        if (aligned16(f)) {
          stg4(f, make_float4(fo[0], fo[1], fo[2], fo[3]));
          stg4(f + 4, make_float4(fo[4], fo[5], fo[6], fo[7]));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fo[j];
        }
      }
      float* pt = part_s + a * PART_W;
      pt[3] = sum8<FLAVOR>(el8);
      pt[4] = asq;
    }
    if (tid == 200) MMB_TR(2);
    if (wid == 7 && lane < ne) {  // box orientation term and the obs tail (warp 7's share of the per-env work)
      if (use_tma) mbar_wait(mbar, 0);
      const float* b = root_s + lane * ROOT_ENV + 10 * 13;
      box_s[lane * BOX_W + 4] = box_quat_dist(f4{b[3], b[4], b[5], b[6]}, c.x_goal, c.y_goal, c.z_goal);
      float* tail = obs_s + lane * OBS_ENV + 380;  // ten_ant.py:806-808: box_pos, box_quat, box_targets(=0)
      tail[0] = clampf(b[0], -tclip, tclip); tail[1] = clampf(b[1], -tclip, tclip);
      tail[2] = clampf(b[3], -tclip, tclip); tail[3] = clampf(b[4], -tclip, tclip);
      tail[4] = clampf(b[5], -tclip, tclip); tail[5] = clampf(b[6], -tclip, tclip);
      tail[6] = 0.0f; tail[7] = 0.0f;
    }
    if (tid == 160) MMB_TR(3);
    __syncthreads();                     // B2: box terms ready
    if (tid == 160) MMB_TR(4);
    if (active) {  // goal-distance reward term and the arrival flag (ten_ant.py:1073-1081); the core thread has the ant terms
      const float* bo = box_s + el * BOX_W;
      float gx, gy;
      goal_of(k, bo[2], bo[3], bo[0], bo[1], gx, gy);
      if (has_prev) {
        goal_of(k, bo[10], bo[11], bo[8], bo[9], gbx, gby);
      } else {  // frame 0 without prev_root: the carry written by the previous launch (with overlap_prev: wait for that
                // kernel first; everything above was independent of it)
        if (pdl) griddep_wait();
        const float* gb = p.goal_before + ((int64_t)e * A + k) * 2;
        gbx = __ldcg(gb); gby = __ldcg(gb + 1);
      }
      const float bty = (k & 1) ? goal_offset(k) : -goal_offset(k);
      const float gdb = l2_dist2(0.0f, bty, gbx, gby);
      const float gd = l2_dist2(0.0f, bty, gx, gy);
      float* pt = part_s + a * PART_W;
      pt[1] = fmul(c.goal_dist_reward_scale, fsub(gdb, gd));
      pt[5] = __int_as_float(lim | ((gd < 0.5f) ? 0x100 : 0));
    }
  } else {
    // ================= core role =================
    float pbx = 0.f, pby = 0.f, gbx = 0.f, gby = 0.f;
    if (active && has_prev) {  // carry of step t = ant xy of the previous frame (ten_ant.py:905-914); else: loaded after B2
      const float* rp = prev_frame + ((int64_t)e * 11 + k) * 13;
      pbx = __ldg(rp); pby = __ldg(rp + 1);
    }
    __syncthreads();                     // B1
    if (use_tma) mbar_wait(mbar, 0);
    if (tid == 40) MMB_TR(5);
    float px = 0.f, py = 0.f, pz = 0.f, up_proj = 0.f;
    if (active) {
      const float* r = root_s + el * ROOT_ENV + k * 13;
      const f3 pos = {r[0], r[1], r[2]};
      const f4 q = {r[3], r[4], r[5], r[6]};
      const f3 vel = {r[7], r[8], r[9]};
      const f3 ang = {r[10], r[11], r[12]};
      AntCore o = ant_core<FLAVOR>(pos, q, vel, ang, f4{c.inv_start_rot[0], c.inv_start_rot[1], c.inv_start_rot[2], c.inv_start_rot[3]});
      float* ob = obs_s + el * OBS_ENV + k * 38;   // pairs as STS.64 (rows are 8-byte aligned)
      auto cl = [&](float x) { return clampf(x, -tclip, tclip); };
      *reinterpret_cast<float2*>(ob + 0) = make_float2(cl(pos.x), cl(pos.y));
      *reinterpret_cast<float2*>(ob + 2) = make_float2(cl(pos.z), cl(o.vel_loc.x));
      *reinterpret_cast<float2*>(ob + 4) = make_float2(cl(o.vel_loc.y), cl(o.vel_loc.z));
      *reinterpret_cast<float2*>(ob + 6) = make_float2(cl(o.angvel_loc.x), cl(o.angvel_loc.y));
      *reinterpret_cast<float2*>(ob + 8) = make_float2(cl(o.angvel_loc.z), cl(o.yaw));
      *reinterpret_cast<float2*>(ob + 10) = make_float2(cl(o.roll), cl(o.angle_to_target));
      *reinterpret_cast<float2*>(ob + 12) = make_float2(cl(o.up_proj), cl(o.heading_proj));
      px = pos.x; py = pos.y; pz = pos.z; up_proj = o.up_proj;
    }
    if (tid == 40) MMB_TR(6);
    __syncthreads();                     // B2: box terms ready
    if (active) {
      const float* bo = box_s + el * BOX_W;
      float gx, gy;
      goal_of(k, bo[2], bo[3], bo[0], bo[1], gx, gy);
      if (has_prev) {
        goal_of(k, bo[10], bo[11], bo[8], bo[9], gbx, gby);
      } else {
        if (pdl) griddep_wait();
        const float* pb = p.pos_before + ((int64_t)e * A + k) * 2;
        const float* gb = p.goal_before + ((int64_t)e * A + k) * 2;
        pbx = __ldcg(pb); pby = __ldcg(pb + 1);
        gbx = __ldcg(gb); gby = __ldcg(gb + 1);
      }
      // ten_ant.py:1073-1081 for ant k
      float d_now = l2_dist2(px, py, gx, gy);
      float push = (d_now < 1.5f) ? 0.0f : 1.0f;
      float ant_dist = fsub(l2_dist2(pbx, pby, gbx, gby), d_now);
      float adr = fmul(fmul(c.ant_dist_reward_scale, ant_dist), push);
      float up = (up_proj > 0.93f) ? fadd(0.0f, c.up_weight) : 0.0f;  // ten_ant.py:1187
      bool fallen = pz < c.termination_height;
      float* pt = part_s + a * PART_W;
      pt[0] = adr; pt[2] = up;
      pt[6] = __int_as_float(fallen ? 1 : 0);
    }
  }
  if (tid == 40) MMB_TR(7);
  if (tid == 200) MMB_TR(8);
  fence_async_smem();                    // obs tile writes -> visible to the TMA store engine
  __syncthreads();                       // B3
  if (tid == 0) MMB_TR(9);

  // ---- obs tile out ----
  const int n = ne * OBS_ENV;
  float* obs_dst = nullptr;
  if (tile_clamped) obs_dst = (p.obs_layout == 0) ? p.obs : p.share_obs;
  else obs_dst = p.obs_raw;
  const int64_t dst_stride = tile_clamped ? ((p.obs_layout == 0) ? p.obs_frame_stride : p.share_obs_frame_stride)
                                          : p.obs_raw_frame_stride;
  bool tma_stored = false;
  if (obs_dst) {
    float* g = obs_dst + (int64_t)t * dst_stride + (int64_t)e0 * OBS_ENV;
    if (aligned16(g)) {
      if (tid == 0) tma_store_1d(g, obs_s, (uint32_t)n * 4u);
      tma_stored = true;
    } else {
      tile_store<NT, EPT * OBS_ENV / 4, false>(g, obs_s, n, tid, clip);
    }
  }
  // the finish runs in the first dof warp: warp 0 has the bulk store to issue and to wait for
  if (!gae_on) {
    if (tid >= NA && tid - NA < ne) {
      // progress / reset / carry outputs (executor, T == 1) and the chain words are ordered behind the previous kernel;
      // with a per-set scratch (mmb.h, scratch_per_set) the reports of frames 0..T-2 need no such order
      if (pdl && (t == p.num_frames - 1 || !p.scratch_per_set || !p.scratch)) griddep_wait();
      finish_env(p, t, e0 + (tid - NA), part_s + (tid - NA) * A * PART_W, box_s + (tid - NA) * BOX_W, root_s + (tid - NA) * ROOT_ENV);
      if (tid == NA) MMB_TR(10);
    }
  } else if (!gae_exec) {
    // frames 0..T-2: reward + fallen bit of (env, frame) travel to the executor in ONE fire-and-forget 64-bit store
    // (data and flag are the same word: no fence, nothing comes back, the CTA retires at once)
    // No wait on the previous kernel here: the words belong to this call's storage set, which by the overlap_prev contract
    // (mmb.h) the preceding kernel neither reads nor writes - their last consumer finished before this launch was issued.
    if (tid >= NA && tid - NA < ne) {
      const int en = e0 + (tid - NA);
      bool fallen;
      const float r = env_reward(c, part_s + (tid - NA) * A * PART_W, box_s + (tid - NA) * BOX_W, fallen);
      if (p.rewards) p.rewards[(int64_t)t * p.rewards_frame_stride + en] = r;
      st_relaxed_gpu_u64(reinterpret_cast<unsigned long long*>(p.gae_scratch) + (int64_t)t * N + en,
                         (1ull << 63) | ((unsigned long long)(fallen ? 1u : 0u) << 32) | (unsigned long long)__float_as_uint(r));
    }
  } else if (dof_role) {
    // executor of this tile's envs (frame T-1, launched last): the five dof warps collect the (T-1) x 16 words in one
    // round trip, then one thread per env runs the progress / reset chain (ten_ant.py:896-901,1296-1299), the GAE
    // recurrence (storage.py:51-62, same operations in the same order as mmb_gae_ppo) and writes the carry
    const int T = p.num_frames;
    const int idx = tid - NA;
    unsigned long long* words = reinterpret_cast<unsigned long long*>(p.gae_scratch);
    unsigned* fallen_w = reinterpret_cast<unsigned*>(box_s);   // env el: word el * BOX_W + 5 (unused by the box terms); error flag: word 6
    if (pdl) griddep_wait();
    // Everything that needs a trip to L2 is requested FIRST and consumed later, so the executor's critical path is one round
    // trip: the env's progress / reset state (written by the previous launch, complete after the wait above), this
    // thread's (at most three) hand-over words and (at most four) of the tile's T + 1 value rows.
    int64_t prog = 0;
    bool flag = false;
    if (idx < ne) {
      prog = __ldcg(p.progress_buf + e0 + idx);
      flag = __ldcg(p.reset_buf + e0 + idx) != 0;
    }
    const int nwords = (T - 1) * EPT, nvals = (T + 1) * EPT;
    const int j0 = idx, j1 = idx + NA, j2 = idx + 2 * NA, j3 = idx + 3 * NA;     // 3 x 160 words cover T <= 31, 4 x 160 values T <= 39
    auto on_w = [&](int j) { return j < nwords && (j & (EPT - 1)) < ne; };
    auto on_v = [&](int j) { return j < nvals && (j & (EPT - 1)) < ne; };
    auto word_at = [&](int j) { return words + (int64_t)(j >> 4) * N + e0 + (j & (EPT - 1)); };
    auto value_at = [&](int j) {
      const int tt = j >> 4, e2 = j & (EPT - 1);
      return (tt < T) ? p.gae_values + (int64_t)tt * p.gae_values_frame_stride + e0 + e2 : p.gae_last_values + e0 + e2;
    };
    unsigned long long v0 = on_w(j0) ? ld_relaxed_gpu_u64(word_at(j0)) : 0ull, v1 = on_w(j1) ? ld_relaxed_gpu_u64(word_at(j1)) : 0ull,
                       v2 = on_w(j2) ? ld_relaxed_gpu_u64(word_at(j2)) : 0ull;
    const float x0 = on_v(j0) ? __ldg(value_at(j0)) : 0.f, x1 = on_v(j1) ? __ldg(value_at(j1)) : 0.f,
                x2 = on_v(j2) ? __ldg(value_at(j2)) : 0.f, x3 = on_v(j3) ? __ldg(value_at(j3)) : 0.f;
    float r_own = 0.f;
    bool fallen_own = false;
    if (idx < ne) r_own = env_reward(c, part_s + idx * A * PART_W, box_s + idx * BOX_W, fallen_own);
    if (idx < EPT) fallen_w[idx * BOX_W + 5] = fallen_own ? (1u << (T - 1)) : 0u;
    if (idx == 0) fallen_w[6] = 0u;
    asm volatile("bar.sync 1, %0;" ::"n"(NA) : "memory");   // the partial terms are dead: part_s becomes the GAE staging
    if (idx < ne) {
      if (p.rewards) p.rewards[(int64_t)t * p.rewards_frame_stride + e0 + idx] = r_own;
      rew_s[idx * GLD + T - 1] = r_own;
    }
    auto put_v = [&](int j, float x) { if (on_v(j)) vals_s[(j & (EPT - 1)) * GLD + (j >> 4)] = x; };
    put_v(j0, x0); put_v(j1, x1); put_v(j2, x2); put_v(j3, x3);
    auto take = [&](int j, unsigned long long v, bool fresh) {
      if (!on_w(j)) return;
      unsigned long long* w = word_at(j);
      if (fresh) v = ld_relaxed_gpu_u64(w);
      for (long long c0 = clock64(); !(v >> 63) && clock64() - c0 < CHAIN_SPIN_CYCLES;) {
        __nanosleep(64);                 // a unit launched before this one is still in flight
        v = ld_relaxed_gpu_u64(w);
      }
      if (v >> 63) {
        rew_s[(j & (EPT - 1)) * GLD + (j >> 4)] = __uint_as_float((unsigned)v);
        if ((v >> 32) & 1ull) atomicOr(fallen_w + (j & (EPT - 1)) * BOX_W + 5, 1u << (j >> 4));
        st_relaxed_gpu_u64(w, 0ull);     // self-resetting for the next launch / graph replay
      } else {
        fallen_w[6] = 1u;                // reports missing after ~1 s: flag it, guess nothing (mmb.h, `scratch`)
      }
    };
    take(j0, v0, false);
    take(j1, v1, false);
    take(j2, v2, false);
    for (int j = j3; j < nwords; j += NA) take(j, 0ull, true);          // T = 32 only: 16 more words than 3 x 160 threads
    asm volatile("bar.sync 1, %0;" ::"n"(NA) : "memory");
    if (idx < EPT) {                     // lanes 0-15 of warp 5
      const bool bad = fallen_w[6] != 0u;
      const bool on = idx < ne && !bad;
      const int en = e0 + idx;
      // sum and sum of squares of the raw advantages, exact to ~1e-14 without the fp64 pipe (64x slower than fp32 here; a
      // 16-step dependent DADD chain was the longest part of the executor): compensated fp32 sums (Neumaier) of a and of
      // the exact product a * a = p + fma(a, a, -p); converted to double once, after the loop
      float s1 = 0.f, c1 = 0.f, s2 = 0.f, c2 = 0.f;
      auto acc = [](float& s, float& c, float x) {
        const float t2 = __fadd_rn(s, x);
        c = __fadd_rn(c, (fabsf(s) >= fabsf(x)) ? __fadd_rn(__fsub_rn(s, t2), x) : __fadd_rn(__fsub_rn(x, t2), s));
        s = t2;
      };
      if (on) {
        const uint32_t done_bits = chain_bits(p, en, 0, T, fallen_w[idx * BOX_W + 5], prog, flag);
        p.progress_buf[en] = prog;
        p.reset_buf[en] = flag ? 1 : 0;
        const float gamma = p.gae_gamma, lam = p.gae_lam;
        float adv = 0.0f;
        float next_v = vals_s[idx * GLD + T];
#pragma unroll 1
        for (int tt = T - 1; tt >= 0; --tt) {
          const float r = rew_s[idx * GLD + tt], v = vals_s[idx * GLD + tt];
          const float mg = fmul(fsub(1.0f, (float)((done_bits >> tt) & 1u)), gamma);   // (1 - dones.float()) * gamma
          const float delta = fsub(fadd(r, fmul(mg, next_v)), v);
          adv = fadd(delta, fmul(fmul(mg, lam), adv));
          const float ret = fadd(adv, v);
          const float a = fsub(ret, v);                                                // advantages = returns - values
          p.gae_returns[(int64_t)tt * p.gae_returns_frame_stride + en] = ret;
          p.gae_advantages[(int64_t)tt * p.gae_advantages_frame_stride + en] = a;
          const float sq = __fmul_rn(a, a);
          acc(s1, c1, a);
          acc(s2, c2, sq);
          acc(s2, c2, __fmaf_rn(a, a, -sq));
          next_v = v;
        }
        carry_from_tile(p, en, box_s + idx * BOX_W, root_s + idx * ROOT_ENV);
      }
      double d1 = (double)s1 + (double)c1, d2 = (double)s2 + (double)c2;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0x0000ffffu, d1, o);
        d2 += __shfl_xor_sync(0x0000ffffu, d2, o);
      }
      if (idx == 0) {
        if (bad) {
          atomicAdd(reinterpret_cast<unsigned long long*>(p.scratch) + N, 1ull);
        } else if (p.gae_stats) {        // one fire-and-forget pair per tile, spread over MMB_STAT_SLOTS lines
          double* slot = p.gae_stats + 4 + (size_t)(blockIdx.x % MMB_STAT_SLOTS) * MMB_STAT_SLOT_STRIDE;
          asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(slot), "d"(d1) : "memory");
          asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(slot + 1), "d"(d2) : "memory");
          if (blockIdx.x == 0) atomicAdd(p.gae_stats, (double)N * (double)T);
        }
      }
    }
    if (tid == NA) MMB_TR(10);
  }
  extra_outputs<NT, EPT>(p, t, e0, ne, tid, obs_s, tile_clamped, clip);
  if (tma_stored && tid == 0) tma_store_wait_read();
  if (tid == 0) MMB_TR(11);
}

template <int FLAVOR>
int32_t launch_ten_ant_split(const mmb_ten_ant_params& p, cudaStream_t st) {
  auto kern = ten_ant_split_kernel<FLAVOR>;
  static bool attr_done[MMB_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= MMB_MAX_DEVICES) return MMB_EUNSUPPORTED;
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitSmem::kBytes) != cudaSuccess) return MMB_ECUDA;
    attr_done[dev] = true;
  }
  const unsigned tiles = (unsigned)((p.num_envs + SplitSmem::EPT - 1) / SplitSmem::EPT);
  const int prefetch_dist = prefetch_distance();
  {
    LaunchScope ls(K_TEN_ANT, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles, (unsigned)p.num_frames);
    cfg.blockDim = dim3(2 * SplitSmem::EPT * A + 32);
    cfg.dynamicSmemBytes = SplitSmem::kBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.overlap_prev ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, kern, p, prefetch_dist) != cudaSuccess) return MMB_ECUDA;
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

// Fallback for callers that pass no ticket scratch: chain (threads < N) + carry (threads < 10 N) as one kernel.
__global__ void __launch_bounds__(256) ten_ant_post_kernel(const __grid_constant__ mmb_ten_ant_params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = p.num_envs, T = p.num_frames;
  if (i < N * A) load_carry_one(p.root + (int64_t)(T - 1) * p.root_frame_stride, i, p.pos_before, p.goal_before, p.box_before);
  if (i < N) chain_one(p, i);
}

// ten_ant.py:870-882: carry from a root tensor
__global__ void ten_ant_load_carry_kernel(const float* __restrict__ root, int N, float* pos_before, float* goal_before,
                                          float* box_before) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * A) load_carry_one(root, i, pos_before, goal_before, box_before);
}

template <int FLAVOR, int EPT>
int32_t launch_ten_ant(const mmb_ten_ant_params& p, cudaStream_t st) {
  auto kern = ten_ant_kernel<FLAVOR, EPT>;
  static bool attr_done[MMB_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= MMB_MAX_DEVICES) return MMB_EUNSUPPORTED;
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TenAntSmem<EPT>::kBytes) != cudaSuccess)
      return MMB_ECUDA;
    attr_done[dev] = true;
  }
  const int64_t units = (int64_t)((p.num_envs + EPT - 1) / EPT) * p.num_frames;
  if (units > 0x7fffffffLL) return MMB_EUNSUPPORTED;
  {
    LaunchScope ls(K_TEN_ANT, st);
    kern<<<(unsigned)units, EPT * A, TenAntSmem<EPT>::kBytes, st>>>(p, prefetch_distance());
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

}  // namespace
}  // namespace mmb

extern "C" int32_t mmb_ten_ant_step(const mmb_ten_ant_params* pp, void* stream) {
  using namespace mmb;
  if (!pp) return MMB_EINVAL;
  mmb_ten_ant_params p = *pp;
  if (p.num_envs <= 0 || p.num_frames <= 0) return MMB_EINVAL;
  if (!p.root || !p.dof || !p.pos_before || !p.goal_before || !p.box_before || !p.progress_buf || !p.reset_buf)
    return MMB_EINVAL;
  if (p.agent_actions[0]) {    // ten per-agent [N][8] tensors instead of `actions`: single-frame launches of the role-split kernel
    if (p.num_frames != 1 || p.actions) return MMB_EINVAL;
    for (int k = 0; k < 10; ++k)
      if (!p.agent_actions[k]) return MMB_EINVAL;
  } else if (!p.actions) {
    return MMB_EINVAL;
  }
  if (p.num_frames > 65535) return MMB_EUNSUPPORTED;
  if (p.num_frames > 1 && !p.dones_u8 && !p.dones_i64) return MMB_EINVAL;  // the chain needs a [T][N] plane
  if (p.obs_layout < 0 || p.obs_layout > 2) return MMB_EINVAL;
  if (p.flavor != MMB_FLAVOR_CUDA && p.flavor != MMB_FLAVOR_CPU) return MMB_EINVAL;
  if (p.gae_values) {  // fused GAE: horizon-batched launches whose chain is resolved in-kernel
    if (p.num_frames < 2 || p.num_frames > 32 || !p.scratch || !p.gae_scratch || !p.gae_last_values || !p.gae_returns ||
        !p.gae_advantages)
      return MMB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(p.gae_scratch) & 7u) || (p.gae_stats && (reinterpret_cast<uintptr_t>(p.gae_stats) & 7u)))
      return MMB_EALIGN;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // tile size: 16 envs (160 threads, 6 CTAs/SM) by default; MMB_TEN_ANT_EPT=32 selects the 32-env tile (tuning knob)
  static const int ept = [] { const char* v = getenv("MMB_TEN_ANT_EPT"); return v ? ((atoi(v) == 32) ? 32 : 16) : MMB_TEN_ANT_EPT; }();
  // kernel variant: "split" (two threads per ant, default) or "mono" (one thread per ant; MMB_TEN_ANT_VARIANT=mono)
  static const bool split = [] { const char* v = getenv("MMB_TEN_ANT_VARIANT"); return !(v && v[0] == 'm'); }();
  int32_t rc;
  if ((p.gae_values || p.agent_actions[0]) && !split) return MMB_EUNSUPPORTED;   // role-split kernel only
  if (split)
    rc = (p.flavor == MMB_FLAVOR_CUDA) ? launch_ten_ant_split<FLAVOR_CUDA>(p, st) : launch_ten_ant_split<FLAVOR_CPU>(p, st);
  else if (ept == 16)
    rc = (p.flavor == MMB_FLAVOR_CUDA) ? launch_ten_ant<FLAVOR_CUDA, 16>(p, st) : launch_ten_ant<FLAVOR_CPU, 16>(p, st);
  else
    rc = (p.flavor == MMB_FLAVOR_CUDA) ? launch_ten_ant<FLAVOR_CUDA, 32>(p, st) : launch_ten_ant<FLAVOR_CPU, 32>(p, st);
  if (rc != MMB_OK) return rc;
  if (p.num_frames > 1 && !(p.scratch && p.num_frames <= 32)) {
    // chain + carry after the last frame (a (tile, T-1) CTA must not write the carry a (tile, 0) CTA reads)
    LaunchScope ls(K_TEN_ANT_CHAIN, st);
    ten_ant_post_kernel<<<(p.num_envs * A + 255) / 256, 256, 0, st>>>(p);
    if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  }
  return MMB_OK;
}

// One host call for the interactive per-step path: reset_idx of the flagged envs, then the step (TenAnt.post_physics_step,
// ten_ant.py:894-926).  Two launches, one ABI transition: at N = 4096 the step is host-bound, not kernel-bound.
extern "C" int32_t mmb_ten_ant_env_step(const mmb_reset_params* reset, const mmb_ten_ant_params* step, void* stream) {
  if (!reset || !step) return MMB_EINVAL;
  const int32_t rc = mmb_reset_compact(reset, stream);
  if (rc != MMB_OK) return rc;
  // The step kernel reads nothing the reset launch writes (index lists, staged DOF rows, counts) and the reset launch reads
  // only the flags the step kernel REPLACES at its very end, behind griddepcontrol.wait: so the step is launched as a
  // programmatic dependent of the reset launch - frame loads, observations and partial rewards overlap the compaction, the
  // task-state part (progress / reset flags / carry) is ordered behind it.  Everything older than the reset launch has
  // completed before that launch started (it is an ordinary launch), so the early part races with nothing.
  mmb_ten_ant_params s = *step;
  static const bool pdl = [] { const char* v = getenv("MMB_STEP_PDL"); return !(v && v[0] == '0'); }();   // MMB_STEP_PDL=0: plain launch order
  if (s.num_frames == 1 && pdl) s.overlap_prev = 1;
  return mmb_ten_ant_step(&s, stream);
}

#ifdef MMB_TRACE
extern "C" __attribute__((visibility("default"))) int32_t mmb_dbg_trace(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, mmb::g_trace, sizeof(unsigned long long) * 64 * 16) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int32_t mmb_ten_ant_load_carry(const float* root, int32_t num_envs, float* pos_before, float* goal_before,
                                          float* box_before, void* stream) {
  using namespace mmb;
  if (!root || !pos_before || !goal_before || !box_before || num_envs <= 0) return MMB_EINVAL;
  {
    LaunchScope ls(K_TEN_ANT_CARRY, (cudaStream_t)stream);
    ten_ant_load_carry_kernel<<<(num_envs * A + 255) / 256, 256, 0, (cudaStream_t)stream>>>(root, num_envs, pos_before,
                                                                                             goal_before, box_before);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

/*
 * mmb.h - C ABI of libmmb_b200.so: the B200 (sm_100a) kernels behind the rollout hot path of
 * SafeRL-Lab/Massive-MARL-Benchmark (per-step task tensor pipeline + rollout storage).
 *
 * The reference has no FFI: its boundary is Python duck typing (VecTask.step / RolloutStorage).  The
 * Python host classes in massive_marl_benchmark_b200/ keep those interfaces and call the entry
 * points below through ctypes.  Each entry point cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 (MMB_OK) or a negative mmb_status; no exception crosses the ABI;
 *   - all buffers are caller-owned DEVICE pointers (PyTorch allocates); the library never
 *     allocates, frees or retains a pointer after returning;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*); no host sync,
 *     no host read-back; counts the host may want are written to caller-provided device memory;
 *   - re-entrant; no global mutable state besides a per-device attribute cache;
 *   - params structs are plain-old-data, passed by pointer, copied before return;
 *   - "frames": the PhysX step is out of scope, so state arrives as Isaac-Gym-layout frames.  A call
 *     may process num_frames >= 1 consecutive frames of the same env set in one launch
 *     (horizon-batched replay); *_frame_stride are in ELEMENTS of the pointed-to type.
 *   - `flavor` selects which torch device's fp32 association order is reproduced where the two
 *     differ (3- and 8-element sum(-1), 3-element norm, division by a Python scalar): the
 *     reference's arithmetic is "whatever torch does on the device it runs on".
 */
#ifndef MMB_H_
#define MMB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMB_ABI_VERSION 4

#if defined(__GNUC__)
#define MMB_API __attribute__((visibility("default")))
#else
#define MMB_API
#endif

typedef enum {
  MMB_OK = 0,
  MMB_EINVAL = -1,       /* null pointer / non-positive size / inconsistent arguments */
  MMB_EALIGN = -2,       /* a pointer or stride violates the documented alignment */
  MMB_ECUDA = -3,        /* the CUDA runtime reported a launch/config error */
  MMB_EUNSUPPORTED = -4  /* valid request outside what this build implements */
} mmb_status;

enum { MMB_FLAVOR_CUDA = 0, /* torch eager on CUDA (the reference's production device) */
       MMB_FLAVOR_CPU = 1 /* torch on CPU (the pinned oracle) */ };

MMB_API int32_t mmb_abi_version(void);
MMB_API const char* mmb_strerror(int32_t status);
/* number of kernels launched by this library since load (all threads); bench.py reports deltas */
MMB_API uint64_t mmb_launch_count(void);

/* Optional per-kernel device timing (measurement aid, off by default): while enabled every launch is bracketed by
 * a CUDA event pair on its launch stream; mmb_profile_collect synchronises those events, returns the summed
 * duration and the number of launches of one kernel class since the previous collect, and releases the events. */
enum { MMB_K_TEN_ANT = 0, MMB_K_TEN_ANT_CHAIN, MMB_K_TEN_ANT_CARRY, MMB_K_ONE_ANT, MMB_K_ONE_ANT_CHAIN,
       MMB_K_INGENUITY, MMB_K_INGENUITY_CHAIN, MMB_K_RESET, MMB_K_ROLLOUT_ADD, MMB_K_GAE_PPO, MMB_K_ADV_NORM,
       MMB_K_STATS, MMB_K_GAE_MARL, MMB_K_MASKS, MMB_K_GATHER, MMB_K_PERM, MMB_K_MLP_LAYER, MMB_K_LN_CAST, MMB_K_ADV_NORM_XCHG, MMB_K_EPISODE_SCAN, MMB_K_EPISODE_RING, MMB_K_GAUSS_ACT, MMB_K_PPO_LOSS, MMB_K_MAPPO_LOSS, MMB_K_ADAM_NORM, MMB_K_ADAM, MMB_K_COUNT };
MMB_API int32_t mmb_profile_enable(int32_t on);
MMB_API int32_t mmb_profile_collect(int32_t kernel_id, double* total_ms, int64_t* count);

/* ------------------------------------------------------------------------------------------ */
/* Task constants: cfg/<Task>.yaml `env:` block + per-task literals     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  float dof_vel_scale;        /* env.dofVelocityScale  0.2  */
  float contact_force_scale;  /* env.contactForceScale 0.1  */
  float power_scale;          /* env.powerScale        1.0  */
  float up_weight;            /* env.upWeight          0.1  */
  float actions_cost;         /* env.actionsCost       0.005 */
  float energy_cost;          /* env.energyCost        0.05 */
  float joints_at_limit_cost; /* env.jointsAtLimitCost 0.1  */
  float death_cost;           /* env.deathCost        -2.0  */
  float termination_height;   /* env.terminationHeight 0.31 */
  float dt;                   /* sim.dt                0.0166 */
  float max_episode_length;   /* env.episodeLength     1000 (compared as progress >= len-1) */
  float quat_reward_scale;    /* ten_ant.py:57 = 0, one_ant.py:58 = 1 */
  float ant_dist_reward_scale;  /* 500 */
  float goal_dist_reward_scale; /* 500 */
  float x_goal, y_goal, z_goal; /* 0,1,0 (ten_ant.py:199-201) */
  float dof_lower[8];         /* per-ant DOF limits (rad), lower<=upper (ten_ant.py:590-596) */
  float dof_upper[8];
  float joint_gears[8];       /* motor_effort per DOF (nv_ant.xml:83-90 = 15) */
  float inv_start_rot[4];     /* quat_conjugate(start_rotation) = (-0,-0,-0,1) (ten_ant.py:166) */
  float initial_dof_pos[8];   /* ten_ant.py:133-137 */
} mmb_ant_consts;

/* ------------------------------------------------------------------------------------------ */
/* TenAnt: one or more env-steps of obs + reward + done + progress + carry + action forces.      */
/* Replaces TenAnt.pre_physics_step / post_physics_step minus reset_idx                           */
/*   (agents/tasks/ten_ant.py:886-926, :712-808, :635-710, jit fns :935-1393),                   */
/*   the clamp passes of VecTaskPython.step (agents/tasks/agent_base/vec_task.py:126-131) and     */
/*   the per-agent split of MultiVecTaskPython.step (agent_base/multi_vec_task.py:94-144).        */
/* reset_idx (ten_ant.py:810-884) is mmb_reset_compact, to be enqueued BEFORE this call on the    */
/* same stream (it reads the reset flags this call overwrites).                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_envs;    /* N */
  int32_t num_frames;  /* T >= 1 */
  int32_t flavor;      /* MMB_FLAVOR_* */
  int32_t obs_layout;  /* 0: obs rows [N][388];  1: per-agent rows [N][10][46] (+ share_obs);  2: agent-major per-agent rows
                        * obs[a * obs_agent_stride + t * obs_frame_stride + env * 46] (+ share_obs): the planes of a shared
                        * MARL rollout buffer, written in place */
  /* inputs, Isaac layout */
  const float* root;    int64_t root_frame_stride;    /* [T][11N][13] */
  const float* dof;     int64_t dof_frame_stride;     /* [T][80N][2]  */
  const float* actions; int64_t actions_frame_stride; /* [T][N][80], before the +-clip_actions clamp */
  float clip_actions;   /* VecTask clip (1.0); +inf disables */
  float clip_obs;       /* VecTask 5.0 / MultiVecTask 7.0; +inf disables */
  /* carry, in/out: values after step -1 on entry, after step T-1 on exit */
  float* pos_before;      /* [N][10][2] */
  float* goal_before;     /* [N][10][2] */
  float* box_before;      /* [N][2]     */
  int64_t* progress_buf;  /* [N] */
  int64_t* reset_buf;     /* [N] */
  /* outputs; any pointer may be NULL to skip that output */
  float* obs_raw;   int64_t obs_raw_frame_stride;   /* task.obs_buf, unclamped [T][N][388] */
  float* obs;       int64_t obs_frame_stride;       /* clamped, layout per obs_layout */
  float* share_obs; int64_t share_obs_frame_stride; /* obs_layout 1: clamped [T][N][388], once per env */
  float* rewards;   int64_t rewards_frame_stride;   /* [T][N] */
  int64_t* dones_i64; int64_t dones_i64_frame_stride; /* [T][N] reset_buf after each step */
  uint8_t* dones_u8;  int64_t dones_u8_frame_stride;  /* [T][N] same, as RolloutStorage.dones */
  float* forces;    int64_t forces_frame_stride;    /* [T][N][80] clamp(actions)*gear*power_scale */
  /* T > 1 only, optional: N + 1 zero-initialised uint64 words owned by the caller (words 0..N-1 are self-resetting; word N
   * counts chain time-outs, see below).  With it (and T <= 32) the progress/reset chain and the carry are resolved inside
   * the main kernel by the unit that processes the last frame of an env; without it (NULL) a second small kernel does the
   * same work.  The last-frame unit waits for the reports of the env's other frames (units launched before it); if they
   * have not all arrived after ~1 s (preemption, a debugger) it does NOT guess: it adds 1 to word N, leaves the env's flags
   * and carry untouched, and the host (TenAnt.chain_errors) raises. */
  uint64_t* scratch;
  /* != 0: programmatic dependent launch.  The kernel may start while the PRECEDING kernel in the stream is still
   * draining (its tail overlaps this kernel's head); it orders itself behind that kernel only where it touches the
   * task state (carry, progress_buf, reset_buf, scratch).  The caller asserts that every other tensor of this call -
   * frames, actions, all outputs - is neither written nor read by the preceding kernel in the stream (true when
   * consecutive rollouts alternate between frame / storage sets; false when the frames were just produced by a
   * simulation kernel).  0 = ordinary stream order. */
  int32_t overlap_prev;
  /* != 0 (with overlap_prev): `scratch` belongs to this call's frame / storage set like its outputs - the preceding kernel
   * in the stream does not touch it - so the chain reports of frames 0..T-2 need not wait for that kernel either.
   * 0: `scratch` is shared by consecutive launches (one per task) and every unit orders its report behind them. */
  int32_t scratch_per_set;
  /* obs_layout 2 only: element stride between the per-agent planes of `obs` (see obs_layout) */
  int64_t obs_agent_stride;
  /* Fused RolloutStorage.compute_returns (storage.py:51-62 + `returns - values`), horizon-batched launches only
   * (2 <= T <= 32, `scratch` given, `dones_u8` given).  gae_values == NULL: off.  The unit of the env's last frame already
   * gathers the T `fallen` bits; with this block the other frames hand it their reward in the same 64-bit word
   * (gae_scratch: T*N zero-initialised, self-resetting words, [T][N]), and it runs the reverse-time scan of mmb_gae_ppo
   * for its 16 envs right after the progress/reset chain: same operations in the same order, bit-identical `returns`.
   * Raw advantages go to gae_advantages, their {sum, sumsq} (fp64) to the slot area of gae_stats (one atomic pair per
   * tile, spread over MMB_STAT_SLOTS lines), the count to gae_stats[0]; normalise with MMB_NORM_SLOTS.
   * `values` / `last_values` must be final before the launch (replayed values; the interactive per-step path keeps
   * mmb_gae_ppo).  gae_scratch / gae_stats / gae_returns / gae_advantages belong to the call's outputs in the sense of
   * overlap_prev: the preceding kernel in the stream must not touch them (use one set per storage). */
  /* Optional: the root tensor [11N][13] of the frame that PRECEDED frame 0 of this call (the task's root_states before the
   * call).  The carry of a step is a pure function of the previous frame (ant xy, box row -> goals; ten_ant.py:870-882,
   * 905-914), so with it frame 0 is processed exactly like frames 1..T-1 and never reads pos_before / goal_before (still
   * written on exit): no unit but the executors of the last frame then depends on the previous launch, and with
   * overlap_prev the next rollout's head runs under this rollout's tail without stalling.  NULL: frame 0 reads the carry
   * arrays (and, with overlap_prev, waits for the preceding kernel).  Role-split kernel only. */
  const float* prev_root;
  /* Optional, T == 1: the ten per-agent action tensors [N][8] exactly as the policies return them, instead of `actions`
   * (which must then be NULL): the hstack of multi_vec_task.py:94-103 is never materialised.  Role-split kernel only. */
  const float* agent_actions[10];
  const float* gae_values;  int64_t gae_values_frame_stride;         /* [T][N] */
  const float* gae_last_values;                                      /* [N] */
  float* gae_returns;       int64_t gae_returns_frame_stride;        /* [T][N] */
  float* gae_advantages;    int64_t gae_advantages_frame_stride;     /* [T][N] raw */
  double* gae_stats;        /* [MMB_ADV_STATS_EXT_DOUBLES], zero-initialised once */
  uint64_t* gae_scratch;    /* [T][N] */
  float gae_gamma, gae_lam; /* already rounded to fp32 (float(gamma)) */
  mmb_ant_consts c;
} mmb_ten_ant_params;

MMB_API int32_t mmb_ten_ant_step(const mmb_ten_ant_params* p, void* stream);

/* Loads the carry (pos_before / goal_before / box_before) from a root tensor, as reset_idx does for
 * ALL envs from the not-yet-refreshed root_states (ten_ant.py:870-882); needed once, before the
 * first step (afterwards the step kernel's own carry is value-identical). */
MMB_API int32_t mmb_ten_ant_load_carry(const float* root /* [11N][13] */, int32_t num_envs, float* pos_before,
                               float* goal_before, float* box_before, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* OneAnt (agents/tasks/one_ant.py:396-415, :346-361, :314-344, jit fns :429-627)                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_envs;
  int32_t num_frames;
  int32_t flavor;
  int32_t reserved0;
  const float* root;    int64_t root_frame_stride;    /* [T][2N][13]: rows 2e ant, 2e+1 box */
  const float* dof;     int64_t dof_frame_stride;     /* [T][8N][2] */
  const float* sensor;  int64_t sensor_frame_stride;  /* [T][N][24] (= [4N][6]) */
  const float* actions; int64_t actions_frame_stride; /* [T][N][8] */
  float clip_actions;
  float clip_obs;
  /* carry in/out */
  float* pos_before;      /* [N][2] */
  float* box_before;      /* [N][2] */
  float* potentials;      /* [N] */
  float* prev_potentials; /* [N] */
  int64_t* progress_buf;
  int64_t* reset_buf;
  /* outputs (NULL to skip) */
  float* obs_raw;   int64_t obs_raw_frame_stride;   /* [T][N][60] */
  float* obs;       int64_t obs_frame_stride;       /* clamped [T][N][60] */
  float* rewards;   int64_t rewards_frame_stride;
  int64_t* dones_i64; int64_t dones_i64_frame_stride;
  uint8_t* dones_u8;  int64_t dones_u8_frame_stride;
  float* forces;    int64_t forces_frame_stride;    /* [T][N][8] */
  float* up_vec;       /* [N][3] after the last frame (task.up_vec) */
  float* heading_vec;  /* [N][3] */
  float* ant_pos;      /* [N][2] */
  float* box_pos;      /* [N][2] */
  float* box_quat;     /* [N][4] */
  mmb_ant_consts c;
} mmb_one_ant_params;

MMB_API int32_t mmb_one_ant_step(const mmb_one_ant_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* MultiIngenuity (agents/tasks/multi_ingenuity.py:268-374, jit :381-453)                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_envs;
  int32_t num_frames;
  int32_t flavor;
  int32_t obs_layout;   /* 0: [N][52];  1: per-agent [N][4][13] is the same memory -> identical */
  const float* root;    int64_t root_frame_stride;    /* [T][4N][13] */
  const float* actions; int64_t actions_frame_stride; /* [T][N][24] */
  float clip_actions;
  float clip_obs;
  float dt;
  float max_episode_length;
  float thrust_upper_limit;        /* 2000 */
  float thrust_lateral_component;  /* 0.2  */
  float thrust_action_speed_scale; /* 2000 */
  float goals[4][3];               /* multi_ingenuity.py:103-106 */
  int64_t* progress_buf;
  int64_t* reset_buf;
  float* obs_raw;   int64_t obs_raw_frame_stride;   /* [T][N][52] */
  float* obs;       int64_t obs_frame_stride;       /* clamped */
  float* rewards;   int64_t rewards_frame_stride;
  int64_t* dones_i64; int64_t dones_i64_frame_stride;
  uint8_t* dones_u8;  int64_t dones_u8_frame_stride;
  float* forces;    int64_t forces_frame_stride;    /* [T][N][24][3] tensor handed to apply_rigid_body_force_tensors */
  float* forces_state; /* [N][24][3] task.forces after the last frame (rows of envs reset in it zeroed) or NULL */
} mmb_ingenuity_params;

MMB_API int32_t mmb_ingenuity_step(const mmb_ingenuity_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* reset_idx for all three tasks: ordered compaction of the reset flags (== reset_buf.nonzero()),*/
/* the int32 actor-index lists of set_actor_root_state_tensor_indexed / set_dof_state_tensor_    */
/* indexed, and the DOF re-randomisation.  ten_ant.py:810-884, one_ant.py:363-391,              */
/* multi_ingenuity.py:231-266.  Rows: row f uses flags[f] (batched over replayed frames).         */
/* ------------------------------------------------------------------------------------------ */
enum { MMB_TASK_TEN_ANT = 0, MMB_TASK_ONE_ANT = 1, MMB_TASK_INGENUITY = 2 };

typedef struct {
  int32_t task;        /* MMB_TASK_* */
  int32_t num_envs;
  int32_t num_rows;    /* F >= 1 */
  int32_t noise_mode;  /* 0: rows of noise_pos/noise_vel (row i feeds the i-th reset env); 1: Philox(seed, env, step) */
  const int64_t* flags_i64; int64_t flags_i64_row_stride; /* one of flags_i64 / flags_u8 */
  const uint8_t* flags_u8;  int64_t flags_u8_row_stride;
  int64_t* env_ids;   int64_t env_ids_row_stride;   /* [F][N]      ascending env ids, first count[f] valid */
  int32_t* index_a;   int64_t index_a_row_stride;   /* [F][N*na]   root call: TenAnt 11/env, OneAnt 2, Ingenuity 4 */
  int32_t* index_b;   int64_t index_b_row_stride;   /* [F][N*nb]   dof call:  TenAnt 10/env, OneAnt 1, Ingenuity 4 */
  int32_t* counts;    /* [F] number of reset envs */
  float* dof_state;   int64_t dof_state_row_stride; /* [F][dofs*N][2] tensor the reset writes into (may be NULL) */
  const float* noise_pos; const float* noise_vel; int64_t noise_row_stride; /* [F][N][8] */
  uint64_t seed; uint64_t step;                     /* Philox key / counter base (noise_mode 1) */
  float* forces_state;                              /* Ingenuity: task.forces rows to zero, or NULL */
  /* Optional, TenAnt / OneAnt: zero-initialised, self-cleaning scratch of num_rows * (ceil(num_envs / 4096) + 1) words.
   * With it a flag row is scanned by one CTA per 4096 envs (chunk totals published to the scratch, every CTA looks back at
   * its predecessors) instead of by a single CTA: needed from ~16 k envs per row upwards.  NULL: one CTA per row. */
  uint64_t* scan_scratch;
  /* Optional (noise_mode 1): the Philox counter base lives in DEVICE memory - read by the launch instead of `step`, then
   * advanced by num_rows (by the launch itself when it is a single CTA - one row of <= 8192 envs -, else by a one-thread
   * kernel behind it) - so that a captured CUDA graph of the per-step path draws fresh numbers on every replay.
   * NULL: `step`. */
  uint64_t* step_counter;
  mmb_ant_consts c;
} mmb_reset_params;

MMB_API int32_t mmb_reset_compact(const mmb_reset_params* p, void* stream);

/* TenAnt.post_physics_step as ONE host call (ten_ant.py:894-926): mmb_reset_compact(reset) followed by
 * mmb_ten_ant_step(step) on `stream` - the interactive per-step path is bound by host time per call, not by the kernels. */
MMB_API int32_t mmb_ten_ant_env_step(const mmb_reset_params* reset, const mmb_ten_ant_params* step, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Rollout storage: PPO (agents/algorithms/rl/ppo/storage.py)                                    */
/* ------------------------------------------------------------------------------------------ */
/* RolloutStorage.add_transitions (storage.py:32-46) for callers that cannot use the step kernel's
 * direct write into the rollout slot: nine copies fused in one launch, dones int64 -> uint8.
 * A NULL source (or destination) skips that plane - the producer has written the slot itself (the step
 * kernel's observation, mmb_gaussian_act's actions / log-probs / sigma) - and the grid follows the
 * widest plane actually copied. */
typedef struct {
  int32_t num_envs, obs_dim, states_dim, act_dim;
  const float* observations; const float* states; const float* actions; const float* rewards;
  const int64_t* dones; const float* values; const float* actions_log_prob; const float* mu; const float* sigma;
  float* dst_observations; float* dst_states; float* dst_actions; float* dst_rewards; uint8_t* dst_dones;
  float* dst_values; float* dst_actions_log_prob; float* dst_mu; float* dst_sigma;   /* slot `step` of each plane */
  int64_t values_stride;                   /* elements between two envs' values (0 or 1: contiguous) - the critic column of a
                                            * wider MLP output goes in without a `.contiguous()` launch */
} mmb_rollout_add_params;
MMB_API int32_t mmb_rollout_add(const mmb_rollout_add_params* p, void* stream);

/* The planes of one team insert in ONE launch: SharedReplayBuffer.insert (agents/algorithms/marl/runner.py:229-275 loops over
 * the agents and copies share_obs, obs, actions, log-probs, value predictions, rewards and the three mask planes one by one)
 * is up to MMB_MAX_COPY_SEGS strided copies of fp32 elements,
 *     dst[i0 * dst_s0 + i1 * dst_s1 + i2] = src[i0 * src_s0 + i1 * src_s1 + i2],   i0 < n0, i1 < n1, i2 < n2
 * (the (N, A, w) tensors of a step into the agent-major slots [A][T + 1][N][w]: i0 = agent, i1 = env). */
#define MMB_MAX_COPY_SEGS 12
typedef struct {
  float* dst; const float* src;
  int32_t n0, n1, n2, _pad;
  int64_t dst_s0, dst_s1, src_s0, src_s1;   /* elements */
} mmb_copy_seg;
typedef struct {
  int32_t count, _pad;
  mmb_copy_seg seg[MMB_MAX_COPY_SEGS];
} mmb_copy_group_params;
MMB_API int32_t mmb_copy_group(const mmb_copy_group_params* p, void* stream);

/* RolloutStorage.compute_returns (storage.py:51-65): reverse-time GAE scan, raw advantages, and the
 * sum / sum-of-squares of the raw advantages (fp64) for the normalisation.  stats[0..2] =
 * {count, sum, sumsq} are ACCUMULATED (caller zeroes them), so env shards on several GPUs can be
 * all-reduced before mmb_adv_normalize. */
typedef struct {
  int32_t num_envs, num_steps;
  const float* rewards;      /* [T][N] */
  const float* values;       /* [T][N] */
  const uint8_t* dones;      /* [T][N] */
  const float* last_values;  /* [N] */
  double gamma, lam;         /* Python floats of cfg/ppo/config.yaml:30-31; each is cast to fp32 where torch does */
  float* returns;            /* [T][N] */
  float* advantages;         /* [T][N] raw (returns - values) */
  double* stats;             /* [4] ({count,sum,sumsq}, [3] reserved) or NULL */
} mmb_gae_ppo_params;
MMB_API int32_t mmb_gae_ppo(const mmb_gae_ppo_params* p, void* stream);

/* (adv - mean) / (std_unbiased + eps) in place; mean/std from stats = {count, sum, sumsq} on device.
 * eps = 1e-8 (storage.py:65) or 1e-5 (mappo_trainer.py:199).  `stats` has FOUR doubles: [3] is a ticket counter
 * owned by the library (zero-initialise once).  flags & MMB_NORM_CLEAR: the last block to read the statistics
 * clears them, so the accumulator is ready for the next rollout without a memset launch (CUDA-graph safe).
 * flags & MMB_NORM_SLOTS: `stats` is the extended accumulator of MMB_ADV_STATS_EXT_DOUBLES doubles - the four words
 * above followed by MMB_STAT_SLOTS slots of MMB_STAT_SLOT_STRIDE doubles ({sum, sumsq} at the start of each 128-byte
 * line) - filled by the fused GAE of mmb_ten_ant_step; the slots are summed in index order and cleared with the rest. */
#define MMB_STAT_SLOTS 32
#define MMB_STAT_SLOT_STRIDE 16
#define MMB_ADV_STATS_EXT_DOUBLES (4 + MMB_STAT_SLOTS * MMB_STAT_SLOT_STRIDE)
enum { MMB_NORM_CLEAR = 1, MMB_NORM_SLOTS = 2 };
MMB_API int32_t mmb_adv_normalize(float* advantages, int64_t n, double* stats, float eps, int32_t flags,
                                  void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Env-sharded multi-GPU: exchange of the advantage statistics over NVLink peer memory.            */
/* The reference is single-GPU (storage.py:65 normalises over all envs of the one device); with     */
/* envs sharded over ranks the same normalisation needs the global {count,sum,sumsq}.  Instead of a */
/* collective launch between the GAE scan and the normalisation, a one-warp kernel does the         */
/* exchange itself: lane r stores the shard's three doubles + a sequence flag straight into rank     */
/* r's mailbox, then waits on rank r's flag in its OWN mailbox (local memory) and the shards are      */
/* summed in rank order, so every rank normalises with bit-identical moments.  No host involvement:  */
/* replayable from a CUDA graph; the NVLink latency sits in that warp, off the scan path, and the     */
/* normalise launch that follows finds the global moments ready.                                     */
/*                                                                                                  */
/* Mailbox layout (per rank, device memory of that rank): [slots][world][4] 8-byte words             */
/* {count, sum, sumsq, seq flag}.  Exchange number q (1, 2, ...) uses slot q % slots; slots >= 2      */
/* (a rank can be at most one exchange ahead of a peer).  All ranks must issue the same sequence of   */
/* mmb_adv_normalize_xchg calls on one endpoint.                                                    */
/* `state` is 8 zero-initialised uint64 of local device memory: [1] exchanges completed, [3] error   */
/* count (flag wait timed out or slot overrun), [4..6] the global {count,sum,sumsq} of the last      */
/* exchange (doubles), others reserved.  Calls on one endpoint must be ordered (one stream, or        */
/* events): exchange q's moments are consumed by its normalise launch before exchange q + 1 starts.  */
/* ------------------------------------------------------------------------------------------ */
#define MMB_MAX_RANKS 16
typedef struct mmb_xchg {
  int32_t world, rank, slots;
  int32_t timeout_ms;               /* wait for the peers' flags: 0 = default (10 s), < 0 = wait for ever */
  uint64_t* state;                  /* local, [8] */
  uint64_t* mailbox[MMB_MAX_RANKS]; /* mailbox[r] = rank r's mailbox as mapped in THIS process (mailbox[rank] is local) */
} mmb_xchg;

/* Mailbox memory must be shareable between processes, so the library allocates it (the one exception to "never
 * allocates"): cudaMalloc + zero fill + cudaIpcGetMemHandle.  `handle64` receives the 64-byte IPC handle to send to
 * the other ranks (any transport; dist.StatsExchange uses torch.distributed.all_gather_object). */
MMB_API int64_t mmb_xchg_mailbox_bytes(int32_t world, int32_t slots);
MMB_API int32_t mmb_xchg_alloc(int64_t bytes, void** dev_ptr, uint8_t* handle64);
MMB_API int32_t mmb_xchg_open(const uint8_t* handle64, void** dev_ptr);  /* map a peer's mailbox (enables peer access) */
MMB_API int32_t mmb_xchg_close(void* dev_ptr);                            /* unmap a peer's mailbox */
MMB_API int32_t mmb_xchg_free(void* dev_ptr);                             /* free the local mailbox */

/* Exchange + normalisation: publishes `stats` = this shard's {count,sum,sumsq} (as accumulated by mmb_gae_ppo or the
 * fused GAE; cleared afterwards) to all ranks, waits until all `world` shards of the exchange are in the local mailbox,
 * and applies (adv - mean) / (std_unbiased + eps) in place with the global moments.  The wait is bounded
 * (xchg->timeout_ms).  On a time-out or an overrun slot NOTHING is normalised with partial moments: state[3]++ and the
 * whole advantages plane of this shard is filled with NaN, so the failure is loud in the very next loss; the caller
 * (RolloutStorage.check_exchange) raises.  flags: MMB_NORM_SLOTS as for mmb_adv_normalize (clearing is implied). */
MMB_API int32_t mmb_adv_normalize_xchg(float* advantages, int64_t n, double* stats, const mmb_xchg* xchg, float eps,
                                       int32_t flags, void* stream);

/* RolloutStorage.get_statistics (storage.py:67-73) on device: out[0] = mean trajectory length,
 * out[1] = mean reward.  No host sync. */
MMB_API int32_t mmb_rollout_statistics(const uint8_t* dones, const float* rewards, int32_t num_steps, int32_t num_envs,
                               float* out2, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Gaussian policy head of the rollout-time act(): actions = mean + z * std and their log-probs in */
/* one launch (PPO module.py:73-87 - pass std = exp(log_std)^2, the reference's scale_tril quirk -  */
/* and MARL distributions.py:94-117 - pass std = sigmoid(log_std / x) * y, per-dimension log-probs). */
/* z: rows of `noise` if given (parity mode), 0 if `deterministic`, else Philox(seed, step, index)   */
/* + Box-Muller (same distribution as the reference's generator, a different stream).               */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_rows, act_dim, deterministic;
  int32_t std_group_rows;                  /* 0: one std row for every row; else row r uses std row r / std_group_rows (a team's
                                            * agent-major means [A * N][act] with per-agent std rows [A][act]: std_group_rows = N) */
  const float* mean; int64_t mean_stride;  /* [rows][act_dim] with row stride */
  const float* std;                        /* [act_dim] (or [groups][act_dim]) standard deviation actually applied */
  const float* noise;                      /* [rows][act_dim] standard normal draws, or NULL */
  uint64_t seed, step;
  float* actions;                          /* [rows][act_dim] */
  float* logp_sum;                         /* [rows] sum over dims (PPO), or NULL */
  float* logp_per_dim;                     /* [rows][act_dim] (MARL), or NULL */
  const float* sigma_src;                  /* [act_dim] (or [groups][act_dim]) row to broadcast, or NULL: PPO's act() also returns */
  float* sigma_out;                        /* [rows][act_dim] = sigma_src per row (`log_std.repeat(N, 1)`, module.py:87), or NULL */
  uint64_t* step_counter;                  /* NULL, or MMB_ACT_COUNTER_WORDS device words {step, 0, 0, ...}: the launch takes `step` from
                                            * word 0 instead of the field above and its last block advances it by one - a CUDA-graph replay
                                            * of act() then draws fresh numbers, in the same sequence as a host counter started at the same
                                            * value.  Words 1.. are the two-level ticket that finds the last block (zero between launches) */
} mmb_gaussian_act_params;
#define MMB_ACT_TICKET_LANES 64
#define MMB_ACT_COUNTER_WORDS (2 + MMB_ACT_TICKET_LANES)
MMB_API int32_t mmb_gaussian_act(const mmb_gaussian_act_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* PPO minibatch loss, forward + backward in one launch:               */
/*   ActorCritic.evaluate's distribution part   agents/algorithms/rl/ppo/module.py:95-99,107      */
/*     (MultivariateNormal(mean, scale_tril = diag(exp(log_std)^2)): log_prob, entropy)           */
/*   PPO.update                                 agents/algorithms/rl/ppo/ppo.py:270-302           */
/*     (adaptive-KL estimate, ratio / clipped surrogate, (clipped) value loss, total loss)        */
/* Inputs are what the two MLPs hand over (mean [B][A], value [B]) and the gathered minibatch     */
/* rows; outputs are the scalar sums and the gradients of                                         */
/*   loss = mean(surrogate) + value_loss_coef * mean(value loss) - entropy_coef * mean(entropy)   */
/* with respect to mean, value and log_std, following torch's autograd rules (max: gradient split */
/* evenly on exact ties; clamp: passes on the closed interval).  `sums` (4 + act_dim doubles,     */
/* zeroed by the caller) receives {sum surrogate, sum value loss, sum kl, entropy} and then       */
/* d loss / d log_std[j].  ratio_lo / ratio_hi = (float)(1.0 -/+ clip_param) as the reference's   */
/* double-precision Python scalars round.  act_dim <= 256.                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_rows, act_dim, use_clipped_value_loss, _pad;
  const float* mu; int64_t mu_stride;      /* new action mean [B][A], row stride in elements */
  const float* log_std;                    /* [A] */
  const float* actions;                    /* [B][A] */
  const float* old_logp;                   /* [B] */
  const float* advantages;                 /* [B] */
  const float* value;                      /* [B] new value */
  const float* target_values;              /* [B] value at rollout time (clipped value loss), may be NULL otherwise */
  const float* returns;                    /* [B] */
  const float* old_mu;                     /* [B][A] or NULL (no KL estimate) */
  const float* old_sigma;                  /* [B][A] log-std rows stored at rollout time, or NULL */
  float clip_param, ratio_lo, ratio_hi, value_loss_coef, entropy_coef;
  float k_log_2pi;                         /* set by the library: (float)(act_dim * log(2 pi)) */
  float* logp;                             /* [B] new log-probabilities, or NULL */
  float* grad_mu;                          /* [B][A] d loss / d mean, or NULL */
  float* grad_value;                       /* [B] d loss / d value, or NULL */
  double* sums;                            /* [4 + A], zeroed by the caller */
  float* out;                              /* optional [5 + A]: the finalised fp32 terms {loss, mean surrogate, mean value   */
                                           /* loss, mean kl, entropy} and d loss / d log_std[A], written by the last block   */
                                           /* to finish - which then leaves `sums` and `*ticket` zeroed again, so that one  */
                                           /* scratch serves every later call on the same stream.  NULL = sums only          */
  uint32_t* ticket;                        /* with `out`: a zero-initialised word that lives beside `sums`                  */
} mmb_ppo_loss_params;
MMB_API int32_t mmb_ppo_loss(const mmb_ppo_loss_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* MAPPO minibatch losses of one agent, forward + backward in one launch:      */
/*   FixedNormal.log_probs (per dimension)        agents/algorithms/utils/distributions.py:32-35  */
/*   MAPPO.ppo_update                             agents/algorithms/marl/mappo_trainer.py:127-146 */
/*     imp_weights = exp(sum_j (logp_j - old_logp_j)), clipped surrogate, optional active masks   */
/*   MAPPO.cal_value_loss                         mappo_trainer.py:62-103                         */
/*     clipped value prediction, returns normalised with the given (mean, var) (PopArt /          */
/*     ValueNorm, popart.py:59-60; NULL = raw returns), huber (agents/utils/util.py:23-26, incl.   */
/*     its missing e < -delta branch) or mse, max with the clipped term, optional active masks    */
/* `std` is the standard deviation actually applied (sigmoid(log_std / x) * y, computed by the    */
/* caller so that autograd continues to log_std); the entropy term depends on std only and stays  */
/* with the caller.  `sums` (2 + act_dim doubles, zeroed by the caller) receives                 */
/*   {sum_b -min(surr1, surr2) [* active], sum_b value loss [* active]} and d policy_loss/d std[j]; */
/* gradients are already divided by B or by *mask_sum (= active_masks.sum(), device scalar).      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_rows, act_dim;
  int32_t use_huber_loss, use_clipped_value_loss, use_value_active_masks, use_policy_active_masks;
  const float* mean; int64_t mean_stride;  /* new action mean [B][A], row stride in elements */
  const float* std;                        /* [A] */
  const float* actions;                    /* [B][A] */
  const float* old_logp;                   /* [B][A] per-dimension log-probs stored at rollout time */
  const float* adv_targ;                   /* [B] */
  const float* values;                     /* [B] new values */
  const float* value_preds;                /* [B] values at rollout time */
  const float* returns;                    /* [B] */
  const float* active_masks;               /* [B], may be NULL when neither mask flag is set */
  const float* mask_sum;                   /* device scalar active_masks.sum(), may be NULL likewise */
  const float* ret_mean;                   /* device scalar, or NULL */
  const float* ret_var;                    /* device scalar, or NULL */
  const float* ret_mean_orig;              /* moments for the UNCLIPPED error term, or NULL = the pair above: the   */
  const float* ret_var_orig;               /* reference's PopArt updates its running moments on every call and is   */
                                           /* called once per error term (mappo_trainer.py:80-81, popart.py:38-60)  */
  float clip_param, ratio_lo, ratio_hi, huber_delta;
  float* imp_weights;                      /* [B] or NULL */
  float* logp;                             /* [B][A] new per-dimension log-probs, or NULL */
  float* grad_mean;                        /* [B][A] d policy_loss / d mean, or NULL */
  float* grad_values;                      /* [B] d value_loss / d values, or NULL */
  double* sums;                            /* [2 + A], zeroed by the caller */
  /* Optional in-kernel finalisation (as mmb_ppo_loss): the last block to finish writes out[0] = policy_loss, out[1] =
   * value_loss - the sums over the reference's denominators (B, or *mask_sum where the matching mask flag is set),
   * rounded to fp32 once - and out[2 + j] = d policy_loss / d std[j] in fp32, then hands `sums` and `*ticket` (zero before
   * the first launch) back zeroed: one launch, no follow-up kernels, the scratch is reusable at once. NULL: `sums` only. */
  float* out;                              /* [2 + A] or NULL */
  unsigned int* ticket;                    /* required with `out` */
} mmb_mappo_loss_params;
MMB_API int32_t mmb_mappo_loss(const mmb_mappo_loss_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Episode bookkeeping of the PPO runner (agents/algorithms/rl/ppo/ppo.py:143-157,198-220): running */
/* reward sum / episode length per env over T steps, finished episodes appended in the reference's  */
/* order (step-major, env ascending) to two rings of `window` entries = the deque(maxlen=100) whose */
/* mean is logged.  No host sync (the reference does .cpu() per step).                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_envs, num_steps, window, _pad;
  const float* rewards;       int64_t rewards_row_stride;    /* [T][N] */
  const uint8_t* dones_u8;    int64_t dones_u8_row_stride;   /* [T][N], or NULL and ...            */
  const int64_t* dones_i64;   int64_t dones_i64_row_stride;  /* ... [T][N] int64 (done when > 0)   */
  float* cur_reward_sum;      /* [N] in/out (cur_reward_sum, ppo.py:116) */
  float* cur_episode_length;  /* [N] in/out (cur_episode_length, ppo.py:117; float, as the reference) */
  float* ep_reward;           /* [T][N] scratch: finished-episode reward where the env was done */
  float* ep_length;           /* [T][N] scratch */
  float* reward_ring;         /* [window]: entry k of all finished episodes lives in slot k % window */
  float* length_ring;         /* [window] */
  uint64_t* state;            /* [2] zero-initialised: [0] episodes finished so far, [1] library scratch */
} mmb_episode_params;
MMB_API int32_t mmb_episode_update(const mmb_episode_params* p, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Rollout storage: MARL (agents/algorithms/marl/utils/separated_buffer.py:124-168) with the     */
/* PopArt / ValueNorm denormalisation (popart.py:30-34,64-75) folded in, plus the advantage       */
/* prologue of mappo_trainer.py:189-199.  Planes may be strided over agents (shared buffer).      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_envs, num_steps, num_agents;
  int32_t use_gae, use_proper_time_limits, use_denorm, use_popart;
  /* element (t, env, agent) of plane X lives at X[t*X_t + env*X_e + agent*X_a] */
  const float* rewards;  int64_t rew_t, rew_e, rew_a;      /* T   steps */
  float* value_preds;    int64_t val_t, val_e, val_a;      /* T+1 steps; slot T is overwritten by next_value */
  const float* masks;    int64_t msk_t, msk_e, msk_a;      /* T+1 */
  const float* bad_masks;int64_t bad_t, bad_e, bad_a;      /* T+1 (proper time limits only) */
  const float* next_value; int64_t nv_e, nv_a;             /* [N][A] */
  float* returns;        int64_t ret_t, ret_e, ret_a;      /* T+1 */
  float* advantages;     int64_t adv_t, adv_e, adv_a;      /* T, raw returns - D(value_preds); NULL to skip */
  const float* denorm_mean; const float* denorm_var;       /* [A] device scalars (running_mean_var) */
  double gamma, gae_lambda;  /* Python floats; gamma*gae_lambda is formed in double like the reference does */
  double* stats;         /* [A][4] accumulated {count, sum, sumsq, reserved} per agent, or NULL */
} mmb_gae_marl_params;
MMB_API int32_t mmb_gae_marl(const mmb_gae_marl_params* p, void* stream);

/* Runner.insert mask logic (agents/algorithms/marl/runner.py:229-255): dones [N][A] int64 ->
 * masks, active_masks [N][A] fp32 written straight into slot step+1 of the buffer planes. */
MMB_API int32_t mmb_marl_masks(const int64_t* dones, int32_t num_envs, int32_t num_agents, float* masks, int64_t m_e,
                       int64_t m_a, float* active_masks, int64_t am_e, int64_t am_a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Minibatch shuffle + gather: mini_batch_generator + the nine `.view(-1,.)[indices]` gathers of  */
/* PPO.update (storage.py:75-87, ppo.py:252-264); feed_forward_generator (separated_buffer.py:   */
/* 170-228).  One launch gathers every field of one minibatch into contiguous buffers.           */
/* ------------------------------------------------------------------------------------------ */
#define MMB_MAX_GATHER_FIELDS 16
typedef struct {
  int32_t num_fields;
  int32_t index_mode;  /* 0: indices given (int64, device); 1: stateless bijection on [0,total) keyed by seed;
                          2: identity rows batch_start.. (fused multi-field copy: buffer insert / after_update) */
  int64_t total;       /* T*N rows in the flattened planes */
  int64_t batch_start; /* first position of this minibatch in the permutation */
  int64_t batch_size;
  const int64_t* indices;   /* [batch_size] (mode 0) */
  uint64_t seed;            /* mode 1 */
  int64_t* indices_out;     /* [batch_size] the row ids used (mode 1), or NULL */
  const void* src[MMB_MAX_GATHER_FIELDS];   /* row-major [total][row_bytes] */
  void* dst[MMB_MAX_GATHER_FIELDS];         /* [batch_size][row_bytes] */
  int32_t row_bytes[MMB_MAX_GATHER_FIELDS]; /* multiple of 4 (or 1 for byte planes) */
  /* mode 1 only: 0 / 1 = the bijection permutes single rows (every random access moves row_bytes: a 4-byte plane costs a
   * whole 32-byte DRAM sector per row).  2, 4, 8, 16 = grouped shuffle: the bijection permutes GROUPS of `group` consecutive
   * rows and each group is read with its rows rotated by a hash of the group id - still a permutation of [0, total), every
   * minibatch a random set of row groups (consecutive rows are consecutive envs of one step: independent samples), and
   * every random access moves group * row_bytes contiguous bytes.  total, batch_start and batch_size must be multiples. */
  int32_t group;
  int32_t _reserved;
} mmb_gather_params;
MMB_API int32_t mmb_shuffle_gather(const mmb_gather_params* p, void* stream);
/* random permutation of [0,n) on device (fast mode of mini_batch_generator 'random'): the stateless
 * bijection evaluated for every position; a permutation by construction. */
MMB_API int32_t mmb_permutation(int64_t n, uint64_t seed, int32_t group, int64_t* out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Actor-critic MLP forward on tcgen05 tensor cores: one launch per layer                          */
/*   Y[M,N] = epilogue(X[M,K] . W[N,K]^T + bias)                                                  */
/* PPO  ActorCritic (agents/algorithms/rl/ppo/module.py:25-55): Linear -> ELU, last Linear plain  */
/* MARL Actor/Critic (agents/algorithms/marl/actor_critic.py:42-69,149-168, agents/algorithms/    */
/*      utils/mlp.py:6-65): LayerNorm(in) -> [Linear -> ELU -> LayerNorm] x3 -> Linear             */
/* Operands are bf16 (fp32 accumulation in TMEM, fp32 epilogue); the reference computes in fp32, so */
/* outputs agree to bf16 operand precision (tests state the tolerance).                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t M, N, K;      /* rows, output features, input features (unpadded) */
  int32_t Mpad;         /* rows of x / y buffers: multiple of 128 >= M (pad rows of x are zero) */
  int32_t Kpad;         /* leading dimension of x and w: multiple of 64 >= K, pad columns are zero */
  int32_t Npad;         /* rows of w: multiple of n_tile >= N, pad rows are zero */
  int32_t n_tile;       /* output columns per CTA: multiple of 32, <= 256, or 512 */
  int32_t epilogue;     /* 0: bias -> fp32;  1: bias+ELU -> bf16;  2: bias+ELU+LayerNorm -> bf16 (n_tile == Npad == N) */
  const void* x;        /* bf16 [Mpad][Kpad] */
  const void* w;        /* bf16 [Npad][Kpad] (nn.Linear weight layout, K contiguous) */
  const float* bias;    /* [N] */
  const float* ln_gamma; const float* ln_beta; float ln_eps;   /* epilogue 2 */
  int32_t stages;       /* set by the library: depth of the shared-memory ring (2..4) */
  void* y;              /* epilogue 0: fp32 [M][y_stride]; else bf16 [Mpad][y_stride] (zero-initialised by the caller) */
  int64_t y_stride;     /* elements */
  int32_t overlap_prev; /* != 0: programmatic dependent launch - prologue and the first weight tiles are fetched while the
                         * preceding kernel in the stream drains; only the loads of x wait for it.  The caller asserts
                         * that w / bias / ln_* are not written by that preceding kernel (true inside a forward chain). */
  int32_t operand_type; /* 0: bf16 operands (above).  1 (mmb_mlp_chain only): tf32 - x, w and the hidden layers' y are fp32 and
                         * nothing is padded or copied: x = the observations [M][K] / the previous layer's output [Mpad][K], w = the
                         * nn.Linear weight [N][K] as it is, Kpad = their row pitch in elements (multiple of 4), Npad = N. */
  /* Deferred LayerNorm (optional; per-layer launches only).  The MARL trunks are [Linear, ELU, LayerNorm] x 3 + head
   * (agents/algorithms/utils/mlp.py:6-65): a LayerNorm in the epilogue needs the whole row in one CTA (epilogue 2: n_tile = N =
   * 512, every tensor-memory column, two passes).  Instead the layer that PRODUCES e = ELU(.) stores it un-normalised and
   * writes, per row and per 64 columns (n_tile must be 256: an epilogue thread's share is then one or two of them), the partial
   * sums (sum e, sum e^2) of the bf16-rounded values, summed in column order (`ln_out_stats`); the layer that CONSUMES it runs on weights with the LayerNorm's gamma folded in (w[n][k] * gamma[k], by
   * the caller) and applies   LN(e) . W^T = rstd * (e . (W gamma)^T - mean * c) + W beta   in its epilogue: `ln_in_stats` = the
   * producer's partials, c[n] = sum_k w[n][k] of the folded weights AS STORED (bf16), and W beta added to `bias` by the
   * caller.  Exactly the LayerNorm of the rounded activations, with any n_tile. */
  const float* ln_in_stats; /* [Mpad][ln_in_parts][2] (sum, sum of squares) partials of the input rows, or NULL */
  int32_t ln_in_parts;      /* partials per row (the producer's Npad / 64) */
  int32_t ln_in_n;          /* features per input row the LayerNorm runs over (the producer's N) */
  float ln_in_eps;
  const float* ln_c;        /* [N], required with ln_in_stats */
  float* ln_out_stats;      /* [Mpad][Npad / 64][2]; epilogue 1, n_tile 256 and Npad % 256 == 0 only; or NULL */
} mmb_mlp_layer_params;
MMB_API int32_t mmb_mlp_layer(const mmb_mlp_layer_params* p, void* stream);

/* The same layer for `count` <= MMB_MAX_GROUP independent problems of identical geometry in ONE launch (grid z = problem):
 * the per-agent actor / critic networks of the MARL policies (runner.py:205-217 runs 2 x num_agents small forwards per
 * env step).  `params` is an array of `count` structs; M, N, K, paddings, n_tile, epilogue and y_stride must agree. */
#define MMB_MAX_GROUP 16
MMB_API int32_t mmb_mlp_layer_group(const mmb_mlp_layer_params* params, int32_t count, void* stream);
/* The whole Linear-ELU chain of a PPO network (module.py:25-55) in ONE launch: `layers` holds `count` networks x
 * `num_layers` layer descriptions (row-major; network 0 defines the geometry, the others must agree: actor and critic side
 * by side).  Layer l + 1's `x` must be layer l's `y` (bf16 [Mpad][y_stride], y_stride = Kpad of the next layer); hidden
 * layers: epilogue 1, N a multiple of 256 and <= 1024; last layer: epilogue 0, fp32 y with 16-byte aligned base and pitch.
 * `n_tile`, `stages` are ignored (a cluster of 4 CTAs per 128-row block splits every layer's columns four ways; the layer
 * boundary is a cluster barrier, activations pass through L2).  `overlap_prev` of layers[0] as in mmb_mlp_layer.
 * MMB_EUNSUPPORTED for other geometries: run the layers with mmb_mlp_layer / mmb_mlp_layer_group instead. */
#define MMB_MLP_MAX_LAYERS 6
/* x_fp32 != NULL: `count` pointers to the fp32 [M][K] inputs (row pitch K, K % 4 == 0, 16-byte aligned): the cast to the
 * bf16 operand of layer 0 happens inside the kernel (the otherwise idle epilogue warps write the swizzled tiles) and
 * layers[.][0].x is ignored - no mmb_ln_cast launch, nothing to wait for. */
MMB_API int32_t mmb_mlp_chain(const mmb_mlp_layer_params* layers, int32_t num_layers, int32_t count, const float* const* x_fp32,
                              void* stream);
/* Diagnostics: with MMB_CHAIN_TRACE=1 in the environment the first eight CTAs of every mmb_mlp_chain launch stamp
   %globaltimer (ns): [cta 0..7][layer 0..5][event 0..7: layer begin (producer), first operands in shared memory, last MMA
   issued, accumulator complete, output slices computed and their stores issued, stores landed (wait_group 0), cluster
   barrier passed, -], followed by [cta 0..7][layer 0..5][i 0..15]: operands of the i-th k-block in shared memory.
   Synchronises the device and copies the 8 x 6 x 24 words of the latest launch to `out`; MMB_EINVAL when tracing is off. */
MMB_API int32_t mmb_mlp_chain_trace(uint64_t* out, int32_t words);
/* diagnostic of the experimental cta_group::2 mode (MMB_MLP_PAIR=1): first barrier wait that timed out {code, block x, block y,
 * parity}, all zero if none; clears the record */
MMB_API int32_t mmb_mlp_debug_status(uint32_t* out4);

/* fp32 [M][K] -> optional LayerNorm over K (mlp.py:58-59 feature_norm) -> bf16 [Mpad][Kpad], zero padded: the
 * A operand of the first layer. */
MMB_API int32_t mmb_ln_cast(const float* x, int32_t M, int32_t Mpad, int32_t K, int32_t Kpad, const float* gamma,
                            const float* beta, float eps, int32_t use_ln, void* y_bf16, void* stream);
MMB_API int32_t mmb_ln_cast_group(const float* const* x, int32_t count, int32_t M, int32_t Mpad, int32_t K, int32_t Kpad,
                                  const float* const* gamma, const float* const* beta, float eps, int32_t use_ln,
                                  void* const* y_bf16, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Grouped optimiser step: clip_grad_norm_ + torch.optim.Adam for MANY networks in two launches.   */
/* Replaces the 2 x num_agents Adam instances of the MARL policies                                  */
/* (agents/algorithms/marl/mappo_policy.py:32-37, ippo_policy.py:39-45, happo_policy.py) and the     */
/* per-network clip + step of mappo_trainer.py:143-170 / ippo_trainer.py / happo_trainer.py.         */
/* All parameters / gradients / moments are slices of four flat fp32 buffers; group g = one network  */
/* = elements [group_start[g], group_start[g+1]) (slices start at multiples of 4 elements; padding   */
/* elements have zero gradient and stay zero).  Arithmetic: torch.optim.Adam (amsgrad off) after      */
/* clip_grad_norm_(max_norm, 2); see csrc/adam.cu.  The host passes the per-step scalars              */
/* (step_size = lr / (1 - beta1^step), bc2_sqrt = sqrt(1 - beta2^step)) computed in double.           */
/* ------------------------------------------------------------------------------------------ */
#define MMB_ADAM_MAX_GROUPS 64
typedef struct {
  int32_t num_groups, _pad;
  int64_t total;                            /* elements in each flat buffer (multiple of 4) */
  int64_t group_start[MMB_ADAM_MAX_GROUPS]; /* ascending, [0] = 0 */
  float* params;                            /* [total], 16-byte aligned */
  const float* grads;                       /* [total] */
  float* exp_avg;                           /* [total] */
  float* exp_avg_sq;                        /* [total] */
  double* sumsq;                            /* [num_groups]: mmb_grad_sumsq_group ACCUMULATES into it (caller zeroes);
                                               mmb_adam_group reads it for the clip coefficient */
  float step_size[MMB_ADAM_MAX_GROUPS];     /* lr / bias_correction1 */
  float bc2_sqrt[MMB_ADAM_MAX_GROUPS];      /* sqrt(bias_correction2) */
  float eps[MMB_ADAM_MAX_GROUPS];
  float weight_decay[MMB_ADAM_MAX_GROUPS];
  float max_grad_norm[MMB_ADAM_MAX_GROUPS]; /* <= 0: no clipping for that group */
  float one_minus_beta1, beta2, one_minus_beta2, _pad2;
} mmb_adam_params;
MMB_API int32_t mmb_grad_sumsq_group(const mmb_adam_params* p, void* stream);
MMB_API int32_t mmb_adam_group(const mmb_adam_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMB_H_ */

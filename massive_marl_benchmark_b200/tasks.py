"""Host-side mirror of the reference task classes (agents/tasks/{ten_ant,one_ant,multi_ingenuity}.py on
top of agents/tasks/agent_base/base_task.py), driving the sm_100a kernels through the C ABI.

Same constructor signature, attributes and buffers as the reference (`Task(cfg, sim_params,
physics_engine, device_type, device_id, headless, is_multi_agent=False)`; `obs_buf, states_buf,
rew_buf, reset_buf(int64), progress_buf(int64), randomize_buf, extras, num_envs, num_obs, num_states,
num_actions, device, cfg`; `step / pre_physics_step / post_physics_step / reset_idx /
compute_observations / compute_reward`), so `VecTaskPython(task, rl_device)` and the reference's PPO /
MARL runners work unchanged.  Two deliberate differences, both documented in DESIGN.md:

* the simulator is a FrameProvider (PhysX is out of scope): `provider=` keyword, default = synthetic
  frames seeded by cfg["seed"];
* one `step` = two launches (mmb_reset_compact, mmb_<task>_step) instead of ~2,200 torch ops, and no
  host sync: `len(env_ids)` never reaches the host; the count lives in `reset_count` (device int32).

There is no CPU path: every method that computes calls the CUDA library and fails loudly without it.
"""
import math
import os
from typing import Optional

import torch

from . import _lib as L
from . import synthetic
from .providers import FrameProvider, ReplayProvider

INF = float("inf")
_TEN_ANT_MONO = os.environ.get("MMB_TEN_ANT_VARIANT", "")[:1] == "m"   # the one-thread-per-ant kernel reads the carry arrays


def _device_of(device_type, device_id):
    if device_type in ("cuda", "GPU"):
        return "cuda:%d" % device_id
    raise L.MmbError("massive_marl_benchmark_b200 runs on CUDA devices only (device_type=%r)" % (device_type,))


def _reset_scan_scratch(task, rows):
    """Zero-initialised, self-cleaning scratch for the multi-CTA reset scan (include/mmb.h, `scan_scratch`): only for flag
    rows of more than 8192 envs, where one CTA per row would serialise the compaction."""
    N = task.num_envs
    if N <= 8192:
        return None
    cache = task.__dict__.setdefault("_reset_scratch", {})
    if rows not in cache:
        cache[rows] = torch.zeros(rows * ((N + 4095) // 4096 + 1), dtype=torch.int64, device=task.device)
    return cache[rows]


class BaseTask:
    """Buffers and `step` of reference base_task.py:24-149 (viewer / domain randomisation are PhysX-side
    and out of scope)."""

    def __init__(self, cfg, provider: Optional[FrameProvider], flavor: int):
        self.device = _device_of(cfg.get("device_type", "cuda"), cfg.get("device_id", 0))
        self.device_id = cfg.get("device_id", 0)
        self.headless = cfg.get("headless", True)
        self.num_envs = cfg["env"]["numEnvs"]
        self.num_obs = cfg["env"]["numObservations"]
        self.num_states = cfg["env"].get("numStates", 0)
        self.num_actions = cfg["env"]["numActions"]
        self.control_freq_inv = cfg["env"].get("controlFrequencyInv", 1)
        dev, N = self.device, self.num_envs
        self.obs_buf = torch.zeros((N, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((N, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(N, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(N, device=dev, dtype=torch.long)
        self.progress_buf = torch.zeros(N, device=dev, dtype=torch.long)
        self._randomize_buf = torch.zeros(N, device=dev, dtype=torch.long)
        self._randomize_pending = 0      # `randomize_buf += 1` of every step, applied when somebody looks (see the property)
        self.extras = {}
        self.provider = provider
        self.flavor = flavor
        # fused wrapper clamps (set by VecTaskPython / MultiVecTaskPython; inf = the bare task semantics)
        self.clip_actions = INF
        self.clip_obs = INF
        self.obs_layout = 0
        self.reset_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.env_ids = torch.zeros(N, device=dev, dtype=torch.long)
        self.reset_noise = None          # parity mode: (positions [N,8], velocities [N,8]); None -> Philox
        self.reset_seed = int(cfg.get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
        self._step_count = 0
        self._step_counter_dev = None    # device-resident Philox counter (CUDA-graph replays of the per-step path)
        L.lib()  # fail here, loudly, if the CUDA library is missing

    def use_device_step_counter(self):
        """From now on the reset noise's Philox counter lives on the device (read and incremented by the launches themselves,
        include/mmb.h `step_counter`): a captured CUDA graph of `step` then draws fresh numbers on every replay.  The
        sequence continues where the host-side counter stood."""
        if self._step_counter_dev is None:
            self._step_counter_dev = torch.tensor([self._step_count], dtype=torch.int64, device=self.device)
        return self._step_counter_dev

    @property
    def randomize_buf(self):
        """base_task.py:67 / ten_ant.py:897: a per-env step counter only domain randomisation reads (out of scope).  The
        per-step `+= 1` is a torch launch the host-bound step path can do without: increments are counted on the host and
        applied when the buffer is read."""
        if self._randomize_pending:
            self._randomize_buf += self._randomize_pending
            self._randomize_pending = 0
        return self._randomize_buf

    @randomize_buf.setter
    def randomize_buf(self, value):
        self._randomize_buf, self._randomize_pending = value, 0

    def step(self, actions):
        self.pre_physics_step(actions)
        for _ in range(self.control_freq_inv):
            self.provider.simulate()
        self.post_physics_step()

    def get_states(self):
        return self.states_buf

    _agent_actions = None

    def pre_physics_step(self, actions):
        # forces are produced by the fused step kernel from the same read of `actions`
        self._agent_actions = None
        self.actions = actions if actions.device == self.rew_buf.device else actions.to(self.device)
        if not self.actions.is_contiguous():
            self.actions = self.actions.contiguous()

    def _fresh_out(self, *shape):
        """Output tensor of one step, owned by the CALLER from then on: the reference returns a fresh `torch.clamp(...)`
        tensor per step (vec_task.py:130) and its PPO.run keeps and mutates it (`current_obs = reset()`, then
        `current_obs.copy_(next_obs)` every step, ppo.py:128-139), so the task must never write into a tensor it has
        handed out.  The caching allocator recycles the block once the caller drops it: no copy, no kernel."""
        return torch.empty(shape, device=self.device, dtype=torch.float)


class TenAnt(BaseTask):
    """reference agents/tasks/ten_ant.py.  The 10-ant internal layout is always used (SURVEY finding 6):
    `is_multi_agent=True` -> 10 agents x 8 actions, else one agent with the flat 80 actions / 388 obs."""

    def __init__(self, cfg, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 is_multi_agent=False, provider=None, flavor=L.FLAVOR_CUDA):
        self.cfg = cfg
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.is_multi_agent = is_multi_agent
        env = cfg["env"]
        self.max_episode_length = env.get("episodeLength", 1000)
        self.num_agents = 10 if is_multi_agent else 1
        cfg["env"]["numObservations"] = 388                      # width of obs_buf (ten_ant.py:806-808)
        cfg["env"]["numActions"] = 8 if is_multi_agent else 80    # ten_ant.py:61-67
        cfg["device_type"], cfg["device_id"], cfg["headless"] = device_type, device_id, headless
        N = env["numEnvs"]
        if provider is None:
            fr = synthetic.ten_ant_frames(N, 32, seed=1234 + int(cfg.get("seed", 0)))
            provider = ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=_device_of(device_type, device_id))
        super().__init__(cfg, provider, flavor)
        self.num_actions_total = 80
        dev = self.device
        self.dt = cfg.get("sim", {}).get("dt", 0.0166)
        self.consts = L.default_ant_consts(env, quat_reward_scale=0.0, dt=self.dt)
        self.dof_limits_lower = torch.tensor(list(self.consts.dof_lower), device=dev)
        self.dof_limits_upper = torch.tensor(list(self.consts.dof_upper), device=dev)
        self.initial_root_states = synthetic.ten_ant_initial_root(N).to(dev)   # ten_ant.py:99-100 (velocities 0)
        self.root_states = self.initial_root_states.clone()
        self.dof_state = torch.zeros(N * 80, 2, device=dev)
        self.dof_reset_staging = torch.zeros(N * 80, 2, device=dev)
        self.pos_before = torch.zeros(N, 10, 2, device=dev)
        self.goal_before = torch.zeros(N, 10, 2, device=dev)
        self.box_before = torch.zeros(N, 2, device=dev)
        self.forces = torch.zeros(N, 80, device=dev)
        self.ant_box_indices = torch.zeros(11 * N, device=dev, dtype=torch.int32)
        self.ant_indices = torch.zeros(10 * N, device=dev, dtype=torch.int32)
        self.obs_clamped = torch.zeros(N, 388, device=dev)
        self.obs_all = torch.zeros(N, 10, 46, device=dev)
        self.keep_raw_obs = True
        self.actions = torch.zeros(N, 80, device=dev)
        self._chain_words = torch.zeros(N + 1, device=dev, dtype=torch.int64)   # replay(): per-env flag/count words + error count
        self.use_prev_root = True
        self._p_reset = None
        self._p_step = None
        # reset_idx at the first step reloads the carry from the not-yet-refreshed root tensor
        L.check(L.lib().mmb_ten_ant_load_carry(L.ptr(self.root_states), N, L.ptr(self.pos_before),
                                               L.ptr(self.goal_before), L.ptr(self.box_before), L.stream_ptr()),
                "mmb_ten_ant_load_carry")

    # -- reference method names ---------------------------------------------------------------
    def reset_idx(self, env_ids=None):
        """ten_ant.py:810-884 for the envs flagged in `reset_buf` (env_ids is recomputed on the device)."""
        p = self._reset_params()
        L.check(L.lib().mmb_reset_compact(p, L.stream_ptr()), "mmb_reset_compact")
        self.provider.set_actor_root_state_tensor_indexed(self.initial_root_states, self.ant_box_indices, self.reset_count)
        self.provider.set_dof_state_tensor_indexed(self.dof_reset_staging, self.ant_indices, self.reset_count)

    def _reset_params(self):
        p = self._p_reset
        if p is None:   # built once: only the noise source and the step counter change between calls
            p = self._p_reset = L.ResetParams()
            p.task, p.num_envs, p.num_rows = L.TASK_TEN_ANT, self.num_envs, 1
            p.flags_i64 = L.ptr(self.reset_buf)
            p.env_ids, p.index_a, p.index_b, p.counts = (L.ptr(self.env_ids), L.ptr(self.ant_box_indices),
                                                         L.ptr(self.ant_indices), L.ptr(self.reset_count))
            p.dof_state = L.ptr(self.dof_reset_staging)
            p.seed = self.reset_seed
            p.c = self.consts
            p.scan_scratch = L.ptr(_reset_scan_scratch(self, 1))
        if self.reset_noise is not None:
            p.noise_mode = 0
            self._noise_keep = tuple(t.contiguous() for t in self.reset_noise)
            p.noise_pos, p.noise_vel = L.ptr(self._noise_keep[0]), L.ptr(self._noise_keep[1])
        else:
            p.noise_mode, p.step = 1, self._step_count
            p.step_counter = L.ptr(self._step_counter_dev)
        return p

    def _prev_root_for_replay(self):
        """The frame that preceded the replay = the task's current `root_states` (the last frame processed, or the initial
        root before the first step): frame 0's carry is computed from it inside the kernel (include/mmb.h, prev_root), which
        equals the stored carry bit for bit.  Only the role-split kernel takes it; it must be a contiguous [11N, 13] tensor
        that stays untouched until the launch has run (frames of a replay set / provider are)."""
        r = self.root_states
        if _TEN_ANT_MONO or r is None or not r.is_contiguous() or r.shape != (11 * self.num_envs, 13) or not self.use_prev_root:
            return None
        self._prev_root_keep = r
        return r

    def step_agent_actions(self, action_list):
        """`step` for the ten per-agent (N, 8) action tensors of MultiVecTaskPython.step (multi_vec_task.py:94-103 hstacks
        them first): the kernel reads them through a pointer list, `task.actions` becomes a lazy hstack nobody pays for unless
        it is read."""
        dev = self.rew_buf.device
        al = [a if (a.device == dev and a.dtype == torch.float32 and a.is_contiguous()) else a.to(dev, torch.float32).contiguous()
              for a in action_list]
        if len(al) != 10 or any(a.shape != (self.num_envs, 8) for a in al) or _TEN_ANT_MONO:
            return self.step(torch.cat(tuple(al), dim=1))
        self._agent_actions, self._actions_cat = al, None
        for _ in range(self.control_freq_inv):
            self.provider.simulate()
        self.post_physics_step()

    @property
    def actions(self):
        if self._agent_actions is not None:
            if self._actions_cat is None:
                self._actions_cat = torch.cat(tuple(self._agent_actions), dim=1)
            return self._actions_cat
        return self._actions_t

    @actions.setter
    def actions(self, value):
        self._actions_t = value

    def chain_errors(self) -> int:
        """Horizon-batched launches whose in-kernel progress / reset chain gave up waiting for a frame's report (~1 s:
        preemption, a debugger); the affected envs' flags, carry and returns were left untouched.  Host sync."""
        n = int(self._chain_words[self.num_envs].item())
        for w in getattr(self, "_extra_chain_words", ()):
            n += int(w[self.num_envs].item())
        return n

    def _launch(self, root, dof, actions, T, strides, obs_raw, obs, share_obs, rewards, dones_i64, dones_u8, forces,
                out_strides, overlap_prev=False, obs_layout=None, obs_agent_stride=0, gae=None, prev_root=None, chain_scratch=None):
        p = L.TenAntParams()
        p.prev_root = L.ptr(prev_root)
        if gae is not None:
            v, r, a = gae["values"], gae["returns"], gae["advantages"]
            p.gae_values, p.gae_values_frame_stride = L.ptr(v), v.stride(0)
            p.gae_last_values = L.ptr(gae["last_values"])
            p.gae_returns, p.gae_returns_frame_stride = L.ptr(r), r.stride(0)
            p.gae_advantages, p.gae_advantages_frame_stride = L.ptr(a), a.stride(0)
            p.gae_stats, p.gae_scratch = L.ptr(gae["stats"]), L.ptr(gae["scratch"])
            p.gae_gamma, p.gae_lam = gae["gamma"], gae["lam"]
        p.overlap_prev = 1 if overlap_prev else 0
        p.obs_agent_stride = obs_agent_stride
        p.num_envs, p.num_frames, p.flavor = self.num_envs, T, self.flavor
        p.obs_layout = self.obs_layout if obs_layout is None else obs_layout
        p.root, p.dof, p.actions = L.ptr(root), L.ptr(dof), L.ptr(actions)
        p.root_frame_stride, p.dof_frame_stride, p.actions_frame_stride = strides
        p.clip_actions, p.clip_obs = self.clip_actions, self.clip_obs
        p.pos_before, p.goal_before, p.box_before = L.ptr(self.pos_before), L.ptr(self.goal_before), L.ptr(self.box_before)
        p.progress_buf, p.reset_buf = L.ptr(self.progress_buf), L.ptr(self.reset_buf)
        p.obs_raw, p.obs, p.share_obs, p.rewards = L.ptr(obs_raw), L.ptr(obs), L.ptr(share_obs), L.ptr(rewards)
        p.dones_i64, p.dones_u8, p.forces = L.ptr(dones_i64), L.ptr(dones_u8), L.ptr(forces)
        (p.obs_raw_frame_stride, p.obs_frame_stride, p.share_obs_frame_stride, p.rewards_frame_stride,
         p.dones_i64_frame_stride, p.dones_u8_frame_stride, p.forces_frame_stride) = out_strides
        if chain_scratch is not None:     # a [N + 1] int64 scratch of the caller's frame / storage set (see replay)
            p.scratch, p.scratch_per_set = L.ptr(chain_scratch), 1
            self._chain_words_last = chain_scratch
        else:
            p.scratch = L.ptr(self._chain_words) if T > 1 else None
            self._chain_words_last = self._chain_words
        p.c = self.consts
        L.check(L.lib().mmb_ten_ant_step(p, L.stream_ptr()), "mmb_ten_ant_step")

    # True: `post_physics_step` launches the step kernel only - the reset compaction of the step (ten_ant.py:899-901) was
    # launched ahead by the caller through `reset_idx()`.  The step kernel reads nothing that launch writes, and that launch
    # reads only the flags the previous step left, so it may run any time between the two step kernels.
    reset_ahead = False

    def post_physics_step(self):
        """ten_ant.py:894-926 (+ the fused pre_physics force scaling and wrapper clamps)."""
        self._randomize_pending += 1
        fr = self.provider.frame()
        if self.obs_layout == 0:
            obs = self.obs_clamped = self._fresh_out(self.num_envs, 388)
            share = None
        else:
            obs = self.obs_all = self._fresh_out(self.num_envs, 10, 46)
            share = self.obs_clamped = self._fresh_out(self.num_envs, 388)
        p = self._p_step
        if p is None:   # built once: the per-step call only refreshes the pointers that move
            p = self._p_step = L.TenAntParams()
            p.num_envs, p.num_frames, p.flavor = self.num_envs, 1, self.flavor
            p.pos_before, p.goal_before, p.box_before = L.ptr(self.pos_before), L.ptr(self.goal_before), L.ptr(self.box_before)
            p.progress_buf, p.reset_buf = L.ptr(self.progress_buf), L.ptr(self.reset_buf)
            p.rewards, p.forces = L.ptr(self.rew_buf), L.ptr(self.forces)
            p.c = self.consts
        p.obs_layout, p.clip_actions, p.clip_obs = self.obs_layout, self.clip_actions, self.clip_obs
        self.root_states, self.dof_state = fr["root"], fr["dof"]
        p.root, p.dof = self.root_states.data_ptr(), self.dof_state.data_ptr()
        al = self._agent_actions
        if al is not None:               # ten per-agent tensors: pointers go to the kernel, no hstack
            p.actions = None
            for k in range(10):
                p.agent_actions[k] = al[k].data_ptr()
        else:
            p.actions = self.actions.data_ptr()
            if p.agent_actions[0]:
                for k in range(10):
                    p.agent_actions[k] = None
        p.obs_raw = self.obs_buf.data_ptr() if self.keep_raw_obs else None
        p.obs = obs.data_ptr()
        p.share_obs = share.data_ptr() if share is not None else None
        if self.reset_ahead:
            # the caller has launched this step's reset_idx() already (ppo_rollout: on a side stream, under the policy forward,
            # ordered before this launch by an event): the step kernel alone
            L.check(L.lib().mmb_ten_ant_step(p, L.stream_ptr()), "mmb_ten_ant_step")
        else:
            # reset_idx of the flagged envs (ten_ant.py:899-901) and the step in ONE host call (mmb_ten_ant_env_step)
            pr = self._reset_params()
            L.check(L.lib().mmb_ten_ant_env_step(pr, p, L.stream_ptr()), "mmb_ten_ant_env_step")
        prov = self.provider
        prov.set_actor_root_state_tensor_indexed(self.initial_root_states, self.ant_box_indices, self.reset_count)
        prov.set_dof_state_tensor_indexed(self.dof_reset_staging, self.ant_indices, self.reset_count)
        prov.set_dof_actuation_force_tensor(self.forces)
        self._step_count += 1

    def compute_observations(self):
        raise L.MmbError("compute_observations is fused into post_physics_step (mmb_ten_ant_step)")

    def compute_reward(self, actions=None):
        raise L.MmbError("compute_reward is fused into post_physics_step (mmb_ten_ant_step)")

    # -- horizon-batched replay (B200-native addition, SURVEY.md section 7 hard part 1) ---------
    def replay(self, frames, actions, obs_out, rewards_out, dones_u8_out=None, dones_i64_out=None, forces_out=None,
               share_obs_out=None, obs_raw_out=None, overlap_prev=False, agent_major_obs_out=None, gae=None, chain_scratch=None):
        """Process T consecutive frames in ONE launch (+ the 1-byte/env-step progress chain).

        gae = `RolloutStorage.fused_gae(last_values, gamma, lam)`: the same launch also runs the storage's
        compute_returns scan (returns, raw advantages, their statistics) - follow it with `normalize_advantages()`.

        chain_scratch = `RolloutStorage.chain_scratch()`: the in-kernel progress / reset chain collects its per-frame reports
        in words of the storage set instead of the task's; with overlap_prev the reports then need not wait for the preceding
        launch (which uses another set).

        frames: dict root [T,11N,13], dof [T,80N,2]; actions [T,N,80]; outputs are [T, ...] planes, e.g.
        slices of a rollout storage so that obs / reward / done land in their slots without a copy pass.
        Reset side effects (index lists, DOF re-randomisation) of the T steps are produced afterwards by
        `reset_replay` from the emitted done flags.

        overlap_prev=True launches with programmatic stream serialisation: the kernel may begin while the previous
        kernel in the stream (normally the previous rollout's step kernel) drains, and orders itself behind it only
        for the task state.  Only valid when frames / actions / outputs are not touched by that previous kernel
        (include/mmb.h, `overlap_prev`).

        agent_major_obs_out [A, T, N, 46] (instead of obs_out; with share_obs_out [T, N, 388]): the per-agent rows of
        multi_vec_task.py:105-116 written straight into the agent-major planes of a `SharedReplayBuffer`
        (`buf.obs[:, 1:]`, `buf.share_obs[1:]`), no insert pass."""
        T = actions.shape[0]
        root, dof = frames["root"], frames["dof"]
        s = lambda x: 0 if x is None else x.stride(0)
        layout, agent_stride = None, 0
        if agent_major_obs_out is not None:
            am = agent_major_obs_out
            if obs_out is not None or am.dim() != 4 or am.shape[0] != 10 or am.shape[3] != 46 or am.stride(3) != 1 or am.stride(2) != 46:
                raise ValueError("agent_major_obs_out must be [10, T, N, 46] with contiguous [N, 46] slots (and obs_out None)")
            layout, agent_stride = 2, am.stride(0)
            obs_out = am[0]          # [T, N, 46] view of agent 0: base pointer and frame stride
        self._launch(root, dof, actions, T, (root.stride(0), dof.stride(0), actions.stride(0)),
                     obs_raw_out, obs_out, share_obs_out, rewards_out, dones_i64_out, dones_u8_out, forces_out,
                     (s(obs_raw_out), s(obs_out), s(share_obs_out), s(rewards_out), s(dones_i64_out), s(dones_u8_out),
                      s(forces_out)), overlap_prev=overlap_prev, obs_layout=layout, obs_agent_stride=agent_stride, gae=gae,
                     prev_root=self._prev_root_for_replay(), chain_scratch=chain_scratch)
        if chain_scratch is not None and all(chain_scratch is not w for w in self.__dict__.setdefault("_extra_chain_words", [])):
            self._extra_chain_words.append(chain_scratch)
        self.root_states, self.dof_state = root[T - 1], dof[T - 1]
        self._step_count += T


class OneAnt(BaseTask):
    """reference agents/tasks/one_ant.py"""

    def __init__(self, cfg, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 is_multi_agent=False, provider=None, flavor=L.FLAVOR_CUDA):
        self.cfg = cfg
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.is_multi_agent = is_multi_agent
        env = cfg["env"]
        self.max_episode_length = env.get("episodeLength", 1000)
        self.num_agents = 1
        cfg["env"]["numObservations"] = 60
        cfg["env"]["numActions"] = 8
        cfg["device_type"], cfg["device_id"], cfg["headless"] = device_type, device_id, headless
        N = env["numEnvs"]
        if provider is None:
            fr = synthetic.one_ant_frames(N, 32, seed=1234 + int(cfg.get("seed", 0)))
            provider = ReplayProvider(fr, device=_device_of(device_type, device_id))
        super().__init__(cfg, provider, flavor)
        dev = self.device
        self.dt = cfg.get("sim", {}).get("dt", 0.0166)
        self.consts = L.default_ant_consts(env, quat_reward_scale=1.0, dt=self.dt)
        self.dof_limits_lower = torch.tensor(list(self.consts.dof_lower), device=dev)
        self.dof_limits_upper = torch.tensor(list(self.consts.dof_upper), device=dev)
        self.initial_root_states = synthetic.one_ant_initial_root(N).to(dev)
        self.root_states = self.initial_root_states.clone()
        self.dof_state = torch.zeros(N * 8, 2, device=dev)
        self.vec_sensor_tensor = torch.zeros(N, 24, device=dev)
        self.dof_reset_staging = torch.zeros(N * 8, 2, device=dev)
        # reset_idx at the first step loads the carry from the initial root tensor (one_ant.py:388-389)
        self.pos_before = self.root_states[0::2, :2].clone()
        self.box_before = self.root_states[1::2, :2].clone()
        self.potentials = torch.full((N,), 0.0, device=dev)
        self.potentials += torch.tensor([-4 / self.dt], dtype=torch.float32).to(dev)    # one_ant.py:144
        self.prev_potentials = self.potentials.clone()
        self.up_vec = torch.tensor([0.0, 0.0, 1.0], device=dev).repeat(N, 1)
        self.heading_vec = torch.tensor([1.0, 0.0, 0.0], device=dev).repeat(N, 1)
        self.ant_pos = torch.zeros(N, 2, device=dev)
        self.box_pos = torch.zeros(N, 2, device=dev)
        self.box_quat = torch.zeros(N, 4, device=dev)
        self.forces = torch.zeros(N, 8, device=dev)
        self.ant_box_indices = torch.zeros(2 * N, device=dev, dtype=torch.int32)
        self.ant_indices = torch.zeros(N, device=dev, dtype=torch.int32)
        self.obs_clamped = torch.zeros(N, 60, device=dev)
        self.keep_raw_obs = True
        self.actions = torch.zeros(N, 8, device=dev)

    def reset_idx(self, env_ids=None):
        """one_ant.py:363-391"""
        p = self.__dict__.get("_p_reset")
        if p is None:   # built once: only the noise source and the step counter change between calls
            p = self._p_reset = L.ResetParams()
            p.task, p.num_envs, p.num_rows = L.TASK_ONE_ANT, self.num_envs, 1
            p.flags_i64 = L.ptr(self.reset_buf)
            p.env_ids, p.index_a, p.index_b, p.counts = (L.ptr(self.env_ids), L.ptr(self.ant_box_indices),
                                                         L.ptr(self.ant_indices), L.ptr(self.reset_count))
            p.dof_state = L.ptr(self.dof_reset_staging)
            p.seed = self.reset_seed
            p.c = self.consts
            p.scan_scratch = L.ptr(_reset_scan_scratch(self, 1))
        if self.reset_noise is not None:
            p.noise_mode = 0
            self._noise_keep = tuple(t.contiguous() for t in self.reset_noise)
            p.noise_pos, p.noise_vel = L.ptr(self._noise_keep[0]), L.ptr(self._noise_keep[1])
        else:
            p.noise_mode, p.step = 1, self._step_count
            p.step_counter = L.ptr(self._step_counter_dev)
        L.check(L.lib().mmb_reset_compact(p, L.stream_ptr()), "mmb_reset_compact")
        self.provider.set_actor_root_state_tensor_indexed(self.initial_root_states, self.ant_box_indices, self.reset_count)
        self.provider.set_dof_state_tensor_indexed(self.dof_reset_staging, self.ant_indices, self.reset_count)

    def _launch(self, root, dof, sensor, actions, T, strides, obs_raw, obs, rewards, dones_i64, dones_u8, forces,
                out_strides):
        p = L.OneAntParams()
        p.num_envs, p.num_frames, p.flavor = self.num_envs, T, self.flavor
        p.root, p.dof, p.sensor, p.actions = L.ptr(root), L.ptr(dof), L.ptr(sensor), L.ptr(actions)
        p.root_frame_stride, p.dof_frame_stride, p.sensor_frame_stride, p.actions_frame_stride = strides
        p.clip_actions, p.clip_obs = self.clip_actions, self.clip_obs
        p.pos_before, p.box_before = L.ptr(self.pos_before), L.ptr(self.box_before)
        p.potentials, p.prev_potentials = L.ptr(self.potentials), L.ptr(self.prev_potentials)
        p.progress_buf, p.reset_buf = L.ptr(self.progress_buf), L.ptr(self.reset_buf)
        p.obs_raw, p.obs, p.rewards = L.ptr(obs_raw), L.ptr(obs), L.ptr(rewards)
        p.dones_i64, p.dones_u8, p.forces = L.ptr(dones_i64), L.ptr(dones_u8), L.ptr(forces)
        (p.obs_raw_frame_stride, p.obs_frame_stride, p.rewards_frame_stride, p.dones_i64_frame_stride,
         p.dones_u8_frame_stride, p.forces_frame_stride) = out_strides
        p.up_vec, p.heading_vec, p.ant_pos = L.ptr(self.up_vec), L.ptr(self.heading_vec), L.ptr(self.ant_pos)
        p.box_pos, p.box_quat = L.ptr(self.box_pos), L.ptr(self.box_quat)
        p.c = self.consts
        L.check(L.lib().mmb_one_ant_step(p, L.stream_ptr()), "mmb_one_ant_step")

    def post_physics_step(self):
        """one_ant.py:403-415"""
        self._randomize_pending += 1
        self.reset_idx()
        fr = self.provider.frame()
        self.root_states, self.dof_state = fr["root"], fr["dof"]
        self.vec_sensor_tensor = fr["sensor"].view(self.num_envs, 24)
        obs = self.obs_clamped = self._fresh_out(self.num_envs, 60)
        self._launch(self.root_states, self.dof_state, self.vec_sensor_tensor, self.actions, 1, (0, 0, 0, 0),
                     self.obs_buf if self.keep_raw_obs else None, obs, self.rew_buf, None, None, self.forces,
                     (0, 0, 0, 0, 0, 0))
        self.provider.set_dof_actuation_force_tensor(self.forces)
        self._step_count += 1

    def replay(self, frames, actions, obs_out, rewards_out, dones_u8_out=None, dones_i64_out=None, forces_out=None,
               obs_raw_out=None):
        T = actions.shape[0]
        root, dof, sensor = frames["root"], frames["dof"], frames["sensor"]
        s = lambda x: 0 if x is None else x.stride(0)
        self._launch(root, dof, sensor, actions, T, (root.stride(0), dof.stride(0), sensor.stride(0), actions.stride(0)),
                     obs_raw_out, obs_out, rewards_out, dones_i64_out, dones_u8_out, forces_out,
                     (s(obs_raw_out), s(obs_out), s(rewards_out), s(dones_i64_out), s(dones_u8_out), s(forces_out)))
        self.root_states, self.dof_state = root[T - 1], dof[T - 1]
        self._step_count += T


class MultiIngenuity(BaseTask):
    """reference agents/tasks/multi_ingenuity.py"""

    def __init__(self, cfg, sim_params=None, physics_engine=None, device_type="cuda", device_id=0, headless=True,
                 is_multi_agent=False, provider=None, flavor=L.FLAVOR_CUDA):
        self.cfg = cfg
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.is_multi_agent = is_multi_agent
        env = cfg["env"]
        self.max_episode_length = env.get("episodeLength", 1000)
        self.num_agents = 4 if is_multi_agent else 1
        cfg["env"]["numObservations"] = 52                      # width of obs_buf (multi_ingenuity.py:351-357)
        cfg["env"]["numActions"] = 6 if is_multi_agent else 24   # multi_ingenuity.py:54-61
        cfg["device_type"], cfg["device_id"], cfg["headless"] = device_type, device_id, headless
        N = env["numEnvs"]
        if provider is None:
            fr = synthetic.ingenuity_frames(N, 32, seed=1234 + int(cfg.get("seed", 0)))
            provider = ReplayProvider({"root": fr["root"]}, device=_device_of(device_type, device_id))
        super().__init__(cfg, provider, flavor)
        dev = self.device
        self.dt = getattr(sim_params, "dt", None) or cfg.get("sim", {}).get("dt", 0.0166)
        self.initial_root_states = synthetic.ingenuity_initial_root(N).to(dev)
        self.root_states = self.initial_root_states.clone()
        self.dof_state = torch.zeros(N * 16, 2, device=dev)
        self.thrust_upper_limit, self.thrust_lateral_component = 2000.0, 0.2
        self.forces = torch.zeros(N, 24, 3, device=dev)            # task.forces (post-step state)
        self.forces_applied = torch.zeros(N, 24, 3, device=dev)    # tensor handed to apply_rigid_body_force_tensors
        self.actor_indices = torch.zeros(4 * N, device=dev, dtype=torch.int32)
        self.obs_clamped = torch.zeros(N, 52, device=dev)
        self.keep_raw_obs = True
        self.goals = ((4.0, 2.0, 1.0), (4.0, -2.0, 1.0), (4.0, 6.0, 1.0), (4.0, -6.0, 1.0))
        self.actions = torch.zeros(N, 24, device=dev)

    def reset_idx(self, env_ids=None):
        """multi_ingenuity.py:231-266"""
        p = self.__dict__.get("_p_reset")
        if p is None:   # nothing in it changes between calls (no noise in this reset)
            p = self._p_reset = L.ResetParams()
            p.task, p.num_envs, p.num_rows = L.TASK_INGENUITY, self.num_envs, 1
            p.flags_i64 = L.ptr(self.reset_buf)
            p.env_ids, p.index_a, p.counts = L.ptr(self.env_ids), L.ptr(self.actor_indices), L.ptr(self.reset_count)
            p.noise_mode = 1
        p.dof_state = L.ptr(self.dof_state)
        L.check(L.lib().mmb_reset_compact(p, L.stream_ptr()), "mmb_reset_compact")
        self.provider.set_actor_root_state_tensor_indexed(self.initial_root_states, self.actor_indices, self.reset_count)
        self.provider.set_dof_state_tensor_indexed(self.dof_state, self.actor_indices, self.reset_count)

    def _launch(self, root, actions, T, strides, obs_raw, obs, rewards, dones_i64, dones_u8, forces, out_strides):
        p = L.IngenuityParams()
        p.num_envs, p.num_frames, p.flavor = self.num_envs, T, self.flavor
        p.root, p.actions = L.ptr(root), L.ptr(actions)
        p.root_frame_stride, p.actions_frame_stride = strides
        p.clip_actions, p.clip_obs = self.clip_actions, self.clip_obs
        p.dt, p.max_episode_length = self.dt, self.max_episode_length
        p.thrust_upper_limit, p.thrust_lateral_component = self.thrust_upper_limit, self.thrust_lateral_component
        p.thrust_action_speed_scale = 2000.0
        for h in range(4):
            for j in range(3):
                p.goals[h][j] = self.goals[h][j]
        p.progress_buf, p.reset_buf = L.ptr(self.progress_buf), L.ptr(self.reset_buf)
        p.obs_raw, p.obs, p.rewards = L.ptr(obs_raw), L.ptr(obs), L.ptr(rewards)
        p.dones_i64, p.dones_u8, p.forces = L.ptr(dones_i64), L.ptr(dones_u8), L.ptr(forces)
        (p.obs_raw_frame_stride, p.obs_frame_stride, p.rewards_frame_stride, p.dones_i64_frame_stride,
         p.dones_u8_frame_stride, p.forces_frame_stride) = out_strides
        p.forces_state = L.ptr(self.forces)
        L.check(L.lib().mmb_ingenuity_step(p, L.stream_ptr()), "mmb_ingenuity_step")

    def post_physics_step(self):
        """multi_ingenuity.py:341-349 (+ fused thrust mapping of :268-339 and wrapper clamps)"""
        self.reset_idx()
        fr = self.provider.frame()
        self.root_states = fr["root"]
        obs = self.obs_clamped = self._fresh_out(self.num_envs, 52)
        self._launch(self.root_states, self.actions, 1, (0, 0), self.obs_buf if self.keep_raw_obs else None, obs,
                     self.rew_buf, None, None, self.forces_applied, (0, 0, 0, 0, 0, 0))
        self.provider.apply_rigid_body_force_tensors(self.forces_applied)
        self._step_count += 1

    def replay(self, frames, actions, obs_out, rewards_out, dones_u8_out=None, dones_i64_out=None, forces_out=None,
               obs_raw_out=None):
        T = actions.shape[0]
        root = frames["root"]
        s = lambda x: 0 if x is None else x.stride(0)
        self._launch(root, actions, T, (root.stride(0), actions.stride(0)), obs_raw_out, obs_out, rewards_out,
                     dones_i64_out, dones_u8_out, forces_out,
                     (s(obs_raw_out), s(obs_out), s(rewards_out), s(dones_i64_out), s(dones_u8_out), s(forces_out)))
        self.root_states = root[T - 1]
        self._step_count += T


def reset_replay(task, flags_u8, dof_out=None, noise=None, out=None):
    """Batched reset_idx over the rows of a [F,N] uint8 flag plane (the done flags a replay emitted): returns
    (env_ids [F,N], index_a [F,N*na], index_b [F,N*nb], counts [F]).  Row f = the reset that step f+1
    performs (flags are those left by step f).  `out` = a previous return value to reuse its buffers."""
    F, N = flags_u8.shape
    dev = flags_u8.device
    kind = {TenAnt: L.TASK_TEN_ANT, OneAnt: L.TASK_ONE_ANT, MultiIngenuity: L.TASK_INGENUITY}[type(task)]
    na, nb = {L.TASK_TEN_ANT: (11, 10), L.TASK_ONE_ANT: (2, 1), L.TASK_INGENUITY: (4, 4)}[kind]
    if out is not None:
        env_ids, ia, ib, counts = out
    else:
        env_ids = torch.zeros(F, N, device=dev, dtype=torch.long)
        ia = torch.zeros(F, N * na, device=dev, dtype=torch.int32)
        ib = torch.zeros(F, N * nb, device=dev, dtype=torch.int32)
        counts = torch.zeros(F, device=dev, dtype=torch.int32)
    p = L.ResetParams()
    p.task, p.num_envs, p.num_rows = kind, N, F
    p.flags_u8, p.flags_u8_row_stride = L.ptr(flags_u8), flags_u8.stride(0)
    p.env_ids, p.env_ids_row_stride = L.ptr(env_ids), env_ids.stride(0)
    p.index_a, p.index_a_row_stride = L.ptr(ia), ia.stride(0)
    p.index_b, p.index_b_row_stride = L.ptr(ib), ib.stride(0)
    p.counts = L.ptr(counts)
    if dof_out is not None:
        p.dof_state, p.dof_state_row_stride = L.ptr(dof_out), dof_out.stride(0)
    if noise is not None:
        p.noise_mode = 0
        p.noise_pos, p.noise_vel, p.noise_row_stride = L.ptr(noise[0]), L.ptr(noise[1]), noise[0].stride(0)
    else:
        p.noise_mode, p.seed, p.step = 1, task.reset_seed, task._step_count
        p.step_counter = L.ptr(task._step_counter_dev)      # set -> the launch reads and advances the device-side counter
    if kind != L.TASK_INGENUITY:
        p.c = task.consts
        p.scan_scratch = L.ptr(_reset_scan_scratch(task, F))
    L.check(L.lib().mmb_reset_compact(p, L.stream_ptr()), "mmb_reset_compact")
    return env_ids, ia, ib, counts

"""The rollout phase of the reference's PPO loop as ONE CUDA-graph replay (B200-native addition: the reference has no
counterpart - its loop is ~60 eager torch launches and three host syncs per env step).

What is captured is, launch for launch, the loop of agents/algorithms/rl/ppo/ppo.py:127-157 on the replacement classes:

    for _ in range(num_transitions_per_env):
        actions, actions_log_prob, values, mu, sigma = actor_critic.act(current_obs, current_states)   # ppo.py:132
        next_obs, rews, dones, infos = vec_env.step(actions)                                           # ppo.py:134
        storage.add_transitions(current_obs, current_states, actions, rews, dones, values, ...)        # ppo.py:138
        current_obs.copy_(next_obs)                                                                    # ppo.py:139
    _, _, last_values, _, _ = actor_critic.act(current_obs, current_states)                            # ppo.py:153
    storage.compute_returns(last_values, gamma, lam)                                                   # ppo.py:157

i.e. per env step: the dual-network MLP chain (tcgen05), the sampling / log-prob kernel, the reset compaction, the fused
step kernel and the insert (on a side branch of the graph, under the next step's policy forward); per rollout
the GAE scan and the advantage normalisation.  `current_obs.copy_` has no launch here, and the insert carries four small
planes only: the step kernel writes observation t + 1 straight into `storage.obs_slots[t + 1]`, the sampling kernel writes
actions / log-probs / sigma into their storage slots.  One graph per phase of the frame ring
(captured when the phase first comes up, after one eager rollout), so `run()` costs one `cudaGraphLaunch`.

What makes it replayable: both random streams keep their Philox counters in device memory, advanced by the launches
themselves (`BaseTask.use_device_step_counter`, `PPOActorCriticForward.use_device_step_counter`); the storage slots are
fixed per captured step; the policy's kernel-side weights are refreshed IN PLACE (`sync_parameters`, called before every
replay, so an `optimizer.step()` between two rollouts is seen); everything the host tracks (`provider.cursor`, step counts,
`storage.step`) is re-applied per replay.  The eager loop with the same seeds stores bit-identical planes
(tests/test_gpu_ppo_rollout.py)."""
import os

import torch

from . import _lib as L


class GraphedPPORollout:
    """`run()` = one rollout of `storage.num_transitions_per_env` env steps + `compute_returns`; afterwards `storage` is
    full (`storage.step == T`: iterate `mini_batch_generator`, then `storage.clear()`, as ppo.py:160-166 does) and
    `current_obs` holds the observation the next rollout starts from.

    env      a `VecTaskPython` over a task whose frames are resident in HBM (a looping `ReplayProvider`), already `reset()`
             or not (`run()` resets on first use)
    policy   `mlp.PPOActorCriticForward`
    storage  `storage.RolloutStorage` on the same device
    tracker  optional `episodes.EpisodeTracker`: the episode bookkeeping of ppo.py:143-151 for the rollout's T steps, inside
             the same graph (`tracker.means()` afterwards, no sync)"""

    def __init__(self, env, policy, storage, gamma, lam, tracker=None):
        from .providers import ReplayProvider
        task = env.task
        prov = task.provider
        if not isinstance(prov, ReplayProvider) or not prov.loop:
            raise TypeError("a graphed rollout needs frames resident in HBM (a looping ReplayProvider)")
        rl, td = torch.device(env.rl_device), torch.device(task.device)
        if rl.type != "cuda" or (rl.index is not None and td.index is not None and rl.index != td.index):
            raise TypeError("a graphed rollout needs rl_device == the task's CUDA device (got %s / %s)" % (rl, td))
        if getattr(task, "obs_layout", 0) != 0:
            raise TypeError("a graphed rollout drives the single-agent (flat observation) wrapper")
        if storage.process_group is not None and storage.stats_exchange is None:
            raise TypeError("env-sharded storages: use the peer-memory statistics exchange (dist.StatsExchange), which is "
                            "graph-replayable; the NCCL all-reduce of the moments is a host-issued collective")
        self.env, self.task, self.policy, self.storage = env, task, policy, storage
        self.gamma, self.lam, self.tracker = float(gamma), float(lam), tracker
        self.T = storage.num_transitions_per_env
        self.states = torch.zeros(task.num_envs, 0, device=task.device)
        self.current_obs = None
        self._graphs = {}
        self._side = torch.cuda.Stream(device=task.device)
        # the graph's main branch is captured on a high-priority stream (kernel nodes keep it): when a step ends, the next
        # policy forward - whole SMs per CTA - and the side branch become ready together, and the forward must go first
        self._main = torch.cuda.Stream(device=task.device, priority=-1) if os.environ.get("MMB_ROLLOUT_PRIO", "1") != "0" else None
        self._pool = None
        self._eager_left = 1              # one eager rollout first: modules loaded, caches built before any capture
        self._ahead = hasattr(task, "reset_ahead")
        self._steps_seen = None           # task._step_count when the last run() ended (somebody stepped the env in between?)
        self.captures = 0
        task.use_device_step_counter()
        policy.use_device_step_counter()

    _FRAME_ATTRS = ("root_states", "dof_state", "vec_sensor_tensor")

    def _next_out(self, *shape):
        """`task._fresh_out` during a rollout: the step kernel writes the next observation straight into the storage slot it
        belongs to (`storage.step` counts the inserts made so far)."""
        slot = self.storage.obs_slots[self.storage.step + 1]
        if tuple(shape) != tuple(slot.shape):
            raise L.MmbError("unexpected step output %r (observation slots are %r)" % (tuple(shape), tuple(slot.shape)))
        return slot

    def _body(self):
        """The loop of ppo.py:127-157 without its copies: observation t lives in `storage.obs_slots[t]` (written there by the
        step kernel; slot T of the previous rollout becomes slot 0 first), the sampling kernel writes actions / log-probs /
        sigma into their slots, so the insert of step t only carries mu, values, rewards and dones.  It runs on a side stream
        under the policy forward of step t + 1; the step kernel of t + 1 waits for it, because it overwrites the reward /
        reset buffers the insert reads.  So does (TenAnt) the reset compaction of step t + 1, which needs nothing but the
        flags step t left (`TenAnt.reset_ahead`): the invariant at every rollout boundary is "the next step's reset_idx() has
        been launched"."""
        t, pol, st = self.task, self.policy, self.storage
        main, side = torch.cuda.current_stream(), self._side
        slots = st.obs_slots
        slots[0].copy_(slots[self.T])
        keep, inserted = [], None        # the policy outputs stay allocated until the last insert has been joined
        ahead = self._ahead
        for k in range(self.T):
            cur = slots[k]
            outs = pol.act(cur, self.states, out=(st.actions[k], st.actions_log_prob[k], st.sigma[k]))
            keep.append(outs)
            actions, logp, values, mu, sigma = outs
            if inserted is not None:
                main.wait_event(inserted)
            t.step(actions)
            stepped = torch.cuda.Event()
            stepped.record(main)
            side.wait_event(stepped)
            with torch.cuda.stream(side):
                if ahead:                 # the NEXT step's reset compaction: it reads only the flags this step has just written
                    t.reset_idx()
                st.add_transitions(cur, self.states, actions, t.rew_buf, t.reset_buf, values, logp, mu, sigma)
                inserted = torch.cuda.Event()
                inserted.record(side)
        self.current_obs = slots[self.T]
        last_values = pol.act(self.current_obs, self.states)[2]
        main.wait_event(inserted)
        st.compute_returns(last_values, self.gamma, self.lam)
        if self.tracker is not None:
            self.tracker.update(st.rewards, st.dones)
        del keep

    def start_from(self, obs):
        """Start (or restart) from `obs` = what `env.reset()` / the last `env.step()` returned, instead of resetting on the
        first `run()`."""
        self.storage.obs_slots[self.T].copy_(obs)
        self.current_obs = self.storage.obs_slots[self.T]

    def run(self):
        t, st, prov = self.task, self.storage, self.task.provider
        if st.step != 0:
            raise AssertionError("Rollout buffer overflow")          # storage.py:36 (the storage was not cleared)
        if self.current_obs is None:
            self.start_from(self.env.reset())
        self.policy.sync_parameters()
        if self._ahead and self._steps_seen != t._step_count:
            t.reset_idx()                 # first run, or the env was stepped from outside: (re-)establish the invariant
        try:
            t.reset_ahead = self._ahead
            self._run()
        finally:
            t.reset_ahead = False
            self._steps_seen = t._step_count

    def _run(self):
        t, st, prov = self.task, self.storage, self.task.provider
        if self._eager_left > 0:
            self._eager_left -= 1
            t._fresh_out = self._next_out
            try:
                self._body()
            finally:
                del t._fresh_out
            return
        slot = (prov.cursor + t.control_freq_inv) % prov.num_frames
        rec = self._graphs.get(slot)
        if rec is None:
            before = (prov.cursor, t._step_count, t._randomize_pending, getattr(self.policy, "_calls", 0))
            g = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            t._fresh_out = self._next_out
            try:
                with torch.cuda.graph(g, pool=self._pool, **({"stream": self._main} if self._main is not None else {})):
                    self._body()          # advances the host-side state once; the kernels run at the replay below
            finally:
                del t._fresh_out
            rec = self._graphs[slot] = (g, tuple((k, getattr(t, k)) for k in self._FRAME_ATTRS if hasattr(t, k)),
                                        (prov.cursor - before[0], t._step_count - before[1], t._randomize_pending - before[2],
                                         getattr(self.policy, "_calls", 0) - before[3]))
            self.captures += 1
        else:
            d = rec[2]
            prov.cursor += d[0]
            t._step_count += d[1]
            t._randomize_pending += d[2]
            self.policy._calls = getattr(self.policy, "_calls", 0) + d[3]
            for k, v in rec[1]:
                setattr(t, k, v)
            st.step = self.T
            t.obs_clamped = self.current_obs = st.obs_slots[self.T]
        rec[0].replay()
        if st.stats_exchange is not None:
            st._xchg_pending = True

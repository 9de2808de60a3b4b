"""`RolloutStorage` with the reference's interface (agents/algorithms/rl/ppo/storage.py) on the B200 kernels.

Same constructor, same public tensors (`observations, states, rewards, actions, dones(uint8),
actions_log_prob, values, returns, advantages, mu, sigma`, all `[T, N, .]`), same methods
(`add_transitions, clear, compute_returns, get_statistics, mini_batch_generator`) and the same
`AssertionError("Rollout buffer overflow")`, so the reference's unmodified `PPO` runs on it.

What changes underneath:
  add_transitions      9 copy_ launches            -> 1 launch (mmb_rollout_add)
  compute_returns      ~8 ops x T + 5 (Python loop) -> 2 launches (mmb_gae_ppo reverse scan with fused
                                                      fp64 sum/sumsq, mmb_adv_normalize)
  get_statistics       dones.cpu() host sync        -> 1 launch, result stays on the device
  mini_batch_generator Python list[int] per batch   -> CUDA int64 index tensors (device permutation);
                                                      `gather_minibatch` gathers all fields in 1 launch

`observations` is a view of a `[T+1, N, obs]` allocation: the fused task kernels write the observation
that follows step t straight into slot t+1 (`obs_slots`), so a rollout needs no add_transitions copy of
the 1.5 KB/env observation at all; `roll_last_obs()` moves slot T to slot 0 between rollouts.

Multi-GPU (env-sharded): every shard normalises with the global moments.  Set `stats_exchange` (a
`dist.StatsExchange`) and the normalise kernel itself publishes the shard's (count, sum, sumsq) into every rank's
mailbox over NVLink and waits on its own mailbox for the peers' - no collective launch; or set `process_group`
for the baseline (NCCL all-reduce of the 3 doubles between the two launches).
"""
import torch

from . import _lib as L


class _BatchIterable:
    """What `mini_batch_generator` returns: iterated once per epoch (ppo.py:247-252), reshuffles on every
    `__iter__` when sampler == 'random' (BatchSampler(SubsetRandomSampler) semantics, drop_last=True)."""

    def __init__(self, storage, mini_batch_size):
        self.storage = storage
        self.mb = mini_batch_size

    def __len__(self):
        return self.storage.batch_size // self.mb

    def __iter__(self):
        st = self.storage
        order = st._epoch_order()
        for i in range(0, st.batch_size - self.mb + 1, self.mb):
            yield order[i:i + self.mb]


class RolloutStorage:
    def __init__(self, num_envs, num_transitions_per_env, obs_shape, states_shape, actions_shape, device='cuda:0',
                 sampler='sequential'):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.MmbError("RolloutStorage needs a CUDA device (there is no CPU path); got %r" % (device,))
        L.lib()
        self.device = device
        self.sampler = sampler
        T, N = num_transitions_per_env, num_envs
        self.obs_slots = torch.zeros(T + 1, N, *obs_shape, device=dev)
        self.observations = self.obs_slots[:T]
        self.states = torch.zeros(T, N, *states_shape, device=dev)
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.actions = torch.zeros(T, N, *actions_shape, device=dev)
        self.dones = torch.zeros(T, N, 1, device=dev, dtype=torch.uint8)
        self.actions_log_prob = torch.zeros(T, N, 1, device=dev)
        self.values = torch.zeros(T, N, 1, device=dev)
        self.returns = torch.zeros(T, N, 1, device=dev)
        self.advantages = torch.zeros(T, N, 1, device=dev)
        self.mu = torch.zeros(T, N, *actions_shape, device=dev)
        self.sigma = torch.zeros(T, N, *actions_shape, device=dev)
        self.num_transitions_per_env = T
        self.num_envs = N
        self.step = 0
        self.batch_size = T * N
        self.process_group = None
        self.stats_exchange = None
        # count, sum, sumsq + library ticket, then the slot area the fused GAE of the step kernel accumulates into
        # (include/mmb.h, MMB_ADV_STATS_EXT_DOUBLES)
        self._adv_stats4 = torch.zeros(L.ADV_STATS_EXT_DOUBLES, device=dev, dtype=torch.float64)
        self.adv_stats = self._adv_stats4[:3]
        self._gae_words = None            # [T, N] uint64 hand-over words of the fused GAE (allocated on first use)
        self._stats_in_slots = False      # the pending statistics sit in the slot area (fused GAE) -> MMB_NORM_SLOTS
        self._stats_out = torch.zeros(2, device=dev)
        self.shuffle_seed = 0
        # 'random' sampler, device shuffle: 1 = the bijection permutes single transitions; 2 / 4 / 8 / 16 = grouped shuffle
        # (include/mmb.h, mmb_gather_params.group): groups of that many consecutive envs of one step travel together, so a
        # gathered minibatch reads whole DRAM sectors of every plane (the 4-byte planes at group 8) instead of one 32-byte
        # sector per 4-byte value.  Every minibatch is still a random subset; the order inside it is irrelevant to the loss.
        self.shuffle_group = 1
        self._epoch = 0
        self.permutation_override = None   # parity mode: a host-supplied permutation (e.g. torch.randperm)
        self._obs_dim = int(self.obs_slots[0, 0].numel())
        self._states_dim = int(self.states[0, 0].numel()) if self.states.numel() else 0
        self._act_dim = int(self.actions[0, 0].numel())
        self._p_add = None

    # ------------------------------------------------------------------------------------------
    def add_transitions(self, observations, states, actions, rewards, dones, values, actions_log_prob, mu, sigma):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        c = lambda t: t if t.is_contiguous() else t.contiguous()
        # values as one column of a wider matrix (the critic column of the actor + critic pair's output): read with its row
        # stride by the kernel, no `.contiguous()` launch
        v_stride = 1
        if values.dim() == 2 and values.shape[1] == 1 and values.stride(0) > 1 and values.dtype == torch.float32:
            v_stride = values.stride(0)
        keep = [c(observations), c(states), c(actions), c(rewards), c(dones), values if v_stride > 1 else c(values),
                c(actions_log_prob), c(mu), c(sigma)]
        if keep[4].dtype != torch.int64:
            keep[4] = keep[4].to(torch.int64)
        p = self._p_add
        if p is None:   # built once; destination pointers of every slot are precomputed
            p = self._p_add = L.RolloutAddParams()
            p.num_envs, p.obs_dim, p.states_dim, p.act_dim = self.num_envs, self._obs_dim, self._states_dim, self._act_dim
            planes = (self.observations, self.states, self.actions, self.rewards, self.dones, self.values,
                      self.actions_log_prob, self.mu, self.sigma)
            self._dst_ptrs = [[(pl[t].data_ptr() if pl.numel() else None) for pl in planes]
                              for t in range(self.num_transitions_per_env)]
        dst = self._dst_ptrs[s]
        # a plane the producer has written straight into this slot (the step kernel's observation, the sampling kernel's
        # actions / log-probs / sigma: ppo_rollout.GraphedPPORollout) is not copied onto itself
        (p.observations, p.states, p.actions, p.rewards, p.dones, p.values, p.actions_log_prob, p.mu,
         p.sigma) = [(t.data_ptr() if t.numel() and t.data_ptr() != d else None) for t, d in zip(keep, dst)]
        (p.dst_observations, p.dst_states, p.dst_actions, p.dst_rewards, p.dst_dones, p.dst_values, p.dst_actions_log_prob,
         p.dst_mu, p.dst_sigma) = dst
        p.values_stride = v_stride
        L.check(L.lib().mmb_rollout_add(p, L.stream_ptr()), "mmb_rollout_add")
        self._keep_add = keep
        self.step += 1

    def clear(self):
        self.step = 0

    def roll_last_obs(self):
        """Fused-rollout mode: the observation after the last step becomes slot 0 of the next rollout."""
        self.obs_slots[0].copy_(self.obs_slots[self.num_transitions_per_env])

    # ------------------------------------------------------------------------------------------
    def compute_returns(self, last_values, gamma, lam):
        self.compute_returns_scan(last_values, gamma, lam)
        self.normalize_advantages()

    def compute_returns_scan(self, last_values, gamma, lam):
        self._stats_in_slots = False
        """First half of compute_returns (storage.py:51-62 + the raw `returns - values`): the reverse-time scan and
        the fp64 (count, sum, sumsq) of the raw advantages.  Split out so a caller can overlap the second half
        (statistics all-reduce + normalisation) with the next rollout on another stream."""
        T, N = self.num_transitions_per_env, self.num_envs
        lv = last_values if last_values.is_contiguous() else last_values.contiguous()
        p = L.GaePpoParams()
        p.num_envs, p.num_steps = N, T
        p.rewards, p.values, p.dones, p.last_values = L.ptr(self.rewards), L.ptr(self.values), L.ptr(self.dones), L.ptr(lv)
        p.gamma, p.lam = float(gamma), float(lam)
        p.returns, p.advantages, p.stats = L.ptr(self.returns), L.ptr(self.advantages), L.ptr(self._adv_stats4)
        L.check(L.lib().mmb_gae_ppo(p, L.stream_ptr()), "mmb_gae_ppo")

    def fused_gae(self, last_values, gamma, lam):
        """Arguments for `TenAnt.replay(..., gae=...)`: the step kernel then runs compute_returns_scan itself, in the unit
        that resolves the progress / reset chain of an env (include/mmb.h, `gae_*`), so a horizon-batched rollout is ONE
        launch + normalize_advantages.  `values` (this storage's plane) and `last_values` must be final before the
        launch."""
        T, N = self.num_transitions_per_env, self.num_envs
        if self._gae_words is None:
            self._gae_words = torch.zeros(T, N, device=self.rewards.device, dtype=torch.int64)
        lv = last_values if last_values.is_contiguous() else last_values.contiguous()
        self._stats_in_slots = True
        return {"values": self.values.view(T, N), "last_values": lv, "returns": self.returns.view(T, N),
                "advantages": self.advantages.view(T, N), "stats": self._adv_stats4, "scratch": self._gae_words,
                "gamma": float(gamma), "lam": float(lam)}

    def chain_scratch(self):
        """[N + 1] int64 words for `TenAnt.replay(..., chain_scratch=...)`: the in-kernel progress / reset chain of a
        horizon-batched launch into this storage (include/mmb.h, `scratch` / `scratch_per_set`)."""
        if getattr(self, "_chain_scratch", None) is None:
            self._chain_scratch = torch.zeros(self.num_envs + 1, device=self.rewards.device, dtype=torch.int64)
        return self._chain_scratch

    def normalize_advantages(self):
        """Second half (storage.py:65): [all-reduce of the statistics over env shards] + (adv - mean) / (std + 1e-8)."""
        T, N = self.num_transitions_per_env, self.num_envs
        flags = L.NORM_CLEAR | (L.NORM_SLOTS if self._stats_in_slots else 0)
        if self.stats_exchange is not None:
            L.check(L.lib().mmb_adv_normalize_xchg(L.ptr(self.advantages), T * N, L.ptr(self._adv_stats4), self.stats_exchange.desc, 1e-8,
                                                   flags, L.stream_ptr()), "mmb_adv_normalize_xchg")
            self._xchg_pending = True
            return
        if self.process_group is not None:
            from . import dist as mdist
            if self._stats_in_slots:   # fold the slot area into the three base words first (one tiny torch op; baseline path)
                sl = self._adv_stats4[4:].view(L.STAT_SLOTS, L.STAT_SLOT_STRIDE)
                self._adv_stats4[1:3] += sl[:, :2].sum(0)
                sl.zero_()
                flags = L.NORM_CLEAR
            mdist.all_reduce_stats(self.adv_stats, self.process_group)
        # the normalise launch clears the accumulator for the next rollout (no memset launch; graph-replay safe)
        L.check(L.lib().mmb_adv_normalize(L.ptr(self.advantages), T * N, L.ptr(self._adv_stats4), 1e-8, flags, L.stream_ptr()),
                "mmb_adv_normalize")

    def check_exchange(self):
        """Raises if a peer-memory statistics exchange of this storage's endpoint timed out or was overrun (the kernel
        then filled the advantages with NaN instead of normalising with partial moments).  One host sync; called once per
        update from `mini_batch_generator`."""
        if self.stats_exchange is not None and getattr(self, "_xchg_pending", False):
            self._xchg_pending = False
            n = self.stats_exchange.errors
            if n:
                raise L.MmbError("advantage-statistics exchange failed %d time(s): a peer did not publish its moments within "
                                 "the time-out (or a mailbox slot was overrun); the advantages of that rollout are NaN. "
                                 "Raise StatsExchange(timeout_ms=...) or fall back to process_group (NCCL all-reduce)." % n)

    def get_statistics(self):
        L.check(L.lib().mmb_rollout_statistics(L.ptr(self.dones), L.ptr(self.rewards), self.num_transitions_per_env,
                                               self.num_envs, L.ptr(self._stats_out), L.stream_ptr()),
                "mmb_rollout_statistics")
        return self._stats_out[0], self._stats_out[1]

    # ------------------------------------------------------------------------------------------
    def _epoch_order(self):
        n = self.batch_size
        dev = self.rewards.device
        if self.sampler == "sequential":
            if not hasattr(self, "_arange") or self._arange.numel() != n:
                self._arange = torch.arange(n, device=dev)
            return self._arange
        if self.permutation_override is not None:
            return torch.as_tensor(self.permutation_override, device=dev, dtype=torch.int64)
        out = torch.empty(n, device=dev, dtype=torch.int64)
        seed = (self.shuffle_seed * 0x9E3779B97F4A7C15 + self._epoch) & 0xFFFFFFFFFFFFFFFF
        self._epoch += 1
        L.check(L.lib().mmb_permutation(n, seed, self._group(), L.ptr(out), L.stream_ptr()), "mmb_permutation")
        return out

    def _group(self):
        g = int(self.shuffle_group)
        if g not in (1, 2, 4, 8, 16) or self.batch_size % g:
            raise ValueError("shuffle_group must be 1, 2, 4, 8 or 16 and divide T * N = %d (got %r)" % (self.batch_size, self.shuffle_group))
        return g

    def mini_batch_generator(self, num_mini_batches):
        mini_batch_size = self.batch_size // num_mini_batches
        if self.sampler not in ("sequential", "random"):
            raise ValueError("unknown sampler %r" % (self.sampler,))
        self.check_exchange()
        return _BatchIterable(self, mini_batch_size)

    FIELDS = ("observations", "states", "actions", "values", "returns", "actions_log_prob", "advantages", "mu", "sigma")

    def new_epoch(self):
        """Draws the seed of the next epoch's device permutation for `gather_epoch_minibatch` (the analogue of re-iterating
        the BatchSampler, ppo.py:248-252)."""
        self._epoch_seed = (self.shuffle_seed * 0x9E3779B97F4A7C15 + self._epoch) & 0xFFFFFFFFFFFFFFFF
        self._epoch += 1
        return self._epoch_seed

    def gather_epoch_minibatch(self, k, num_mini_batches, out=None, want_indices=False):
        """Minibatch k of the current epoch (`new_epoch()`), shuffle AND gather in one launch and without an index array:
        position j of the epoch's permutation is computed inside the gather kernel (stateless bijection, include/mmb.h
        index_mode 1, grouped by `shuffle_group`).  Equals `gather_minibatch(order[k*mb:(k+1)*mb])` for the `order` that
        `mini_batch_generator` would have materialised with the same seed.  Returns the dict of `gather_minibatch` (plus
        "indices" with want_indices)."""
        mb = self.batch_size // num_mini_batches
        g = self._group()
        if mb % g:
            raise ValueError("minibatch size %d is not a multiple of shuffle_group %d" % (mb, g))
        fields = [f for f in self.FIELDS if getattr(self, f).numel()]
        dev = self.rewards.device
        if out is None:
            out = {f: torch.empty((mb,) + tuple(getattr(self, f).shape[2:]), device=dev) for f in fields}
        p = L.GatherParams()
        p.num_fields, p.index_mode, p.total, p.batch_start, p.batch_size = len(fields), 1, self.batch_size, k * mb, mb
        p.seed, p.group = self._epoch_seed, g
        if want_indices:
            if "indices" not in out:
                out["indices"] = torch.empty(mb, device=dev, dtype=torch.int64)
            p.indices_out = L.ptr(out["indices"])
        for i, f in enumerate(fields):
            src = getattr(self, f)
            p.src[i], p.dst[i] = src.data_ptr(), out[f].data_ptr()
            p.row_bytes[i] = int(src[0, 0].numel()) * src.element_size()
        L.check(L.lib().mmb_shuffle_gather(p, L.stream_ptr()), "mmb_shuffle_gather")
        return out

    def gather_minibatch(self, indices, out=None):
        """All fields of one minibatch in ONE launch (the 9 gathers of ppo.py:253-264): dict name -> [B, .]."""
        B = int(indices.numel())
        fields = [f for f in self.FIELDS if getattr(self, f).numel()]
        if out is None:
            out = {f: torch.empty((B,) + tuple(getattr(self, f).shape[2:]), device=self.rewards.device) for f in fields}
        p = L.GatherParams()
        p.num_fields, p.index_mode, p.total, p.batch_start, p.batch_size = len(fields), 0, self.batch_size, 0, B
        idx = indices if indices.dtype == torch.int64 and indices.is_contiguous() else indices.to(torch.int64).contiguous()
        p.indices = L.ptr(idx)
        for i, f in enumerate(fields):
            src = getattr(self, f)
            p.src[i], p.dst[i] = src.data_ptr(), out[f].data_ptr()
            p.row_bytes[i] = int(src[0, 0].numel()) * src.element_size()
        L.check(L.lib().mmb_shuffle_gather(p, L.stream_ptr()), "mmb_shuffle_gather")
        return out

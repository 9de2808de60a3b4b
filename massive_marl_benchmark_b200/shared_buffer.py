"""Shared on-device rollout buffer for the MARL runners.

The reference gives every agent its own `SeparatedReplayBuffer` (agents/algorithms/marl/runner.py:53-62,
utils/separated_buffer.py:36-43): the centralised-critic `share_obs` plane - identical for all agents, it is the
`state_all` replica of multi_vec_task.py:118-125 - is stored once PER AGENT (10 x 17 x N x 1,552 B for TenAnt), and
two recurrent-state planes of 2 KB per (step, env, agent) are kept although `use_recurrent_policy` is False
(cfg/mappo/config.yaml:41-42).  `SharedReplayBuffer` stores

    share_obs                              [T+1, N, S]        once per env
    obs, value_preds, returns, masks, ...  [A, T+1, N, .]     agent-major, so every agent's planes are contiguous
    actions, action_log_probs, rewards ... [A, T, N, .]
    rnn_states / rnn_states_critic         one shared all-zero [T+1, N, recurrent_N, hidden] plane (read-only)

and hands out per-agent views with the full `SeparatedReplayBuffer` interface (`agent(i)`: same tensor names, same
methods, same generator tuples), so the reference's trainers run on them unchanged, while the whole-team operations are
single launches:

    insert           one strided copy per field for ALL agents  (reference: 9 copy_ x A)
    compute_returns  one `mmb_gae_marl` launch over (env, agent) with per-agent PopArt / ValueNorm moments
                     (reference: ~12 ops x T x A)
    after_update     one fused row copy

At TenAnt N = 4096, T = 16, A = 10 this is 0.97 GB less share_obs and 2.8 GB less recurrent state than the per-agent
buffers; the values every agent reads are bit-identical (tests/test_gpu_storage.py).
"""
import torch

from . import _lib as L
from .separated_buffer import SeparatedReplayBuffer, _act_dim, _multi_copy, _shape_from_space


class SharedReplayBuffer:
    def __init__(self, config, num_agents, obs_space, share_obs_space, act_space, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.MmbError("SharedReplayBuffer needs a CUDA device (there is no CPU path)")
        L.lib()
        self.config, self.num_agents, self.device = config, num_agents, device
        T, N, A = config["episode_length"], config["n_rollout_threads"], num_agents
        self.episode_length, self.n_rollout_threads = T, N
        obs_shape, share_shape, act = _shape_from_space(obs_space), _shape_from_space(share_obs_space), _act_dim(act_space)
        self.share_obs = torch.zeros(T + 1, N, *share_shape, device=dev)
        self.obs = torch.zeros(A, T + 1, N, *obs_shape, device=dev)
        self.value_preds = torch.zeros(A, T + 1, N, 1, device=dev)
        self.returns = torch.zeros(A, T + 1, N, 1, device=dev)
        self.masks = torch.ones(A, T + 1, N, 1, device=dev)
        self.bad_masks = torch.ones(A, T + 1, N, 1, device=dev)
        self.active_masks = torch.ones(A, T + 1, N, 1, device=dev)
        self.actions = torch.zeros(A, T, N, act, device=dev)
        self.action_log_probs = torch.zeros(A, T, N, act, device=dev)
        self.rewards = torch.zeros(A, T, N, 1, device=dev)
        self.factor = torch.ones(A, T, N, 1, device=dev)
        self.raw_advantages = torch.zeros(A, T, N, 1, device=dev)
        self._zero_rnn = torch.zeros(T + 1, N, config["recurrent_N"], config["hidden_size"], device=dev)
        self._adv_stats4 = torch.zeros(A, 4, device=dev, dtype=torch.float64)
        self.step = 0
        self._agents = [self._make_view(i) for i in range(A)]

    # -- per-agent SeparatedReplayBuffer-compatible views ------------------------------------------------
    def _make_view(self, i):
        v = object.__new__(SeparatedReplayBuffer)
        c = self.config
        v.episode_length, v.n_rollout_threads = self.episode_length, self.n_rollout_threads
        v.rnn_hidden_size, v.recurrent_N = c["hidden_size"], c["recurrent_N"]
        v.gamma, v.gae_lambda = c["gamma"], c["gae_lambda"]
        v._use_gae, v._use_popart, v._use_valuenorm = c["use_gae"], c["use_popart"], c["use_valuenorm"]
        v._use_proper_time_limits = c["use_proper_time_limits"]
        v.device = self.device
        v.share_obs = self.share_obs                       # ONE plane for all agents
        v.rnn_states = v.rnn_states_critic = self._zero_rnn
        for name in ("obs", "value_preds", "returns", "masks", "bad_masks", "active_masks", "actions", "action_log_probs",
                     "rewards", "factor", "raw_advantages"):
            setattr(v, name, getattr(self, name)[i])
        v.available_actions = None
        v._adv_stats4 = self._adv_stats4[i]
        v.adv_stats = v._adv_stats4[:3]
        v.process_group, v.permutation_override = None, None
        v.step = 0
        v._shared_parent = self
        return v

    def agent(self, i):
        v = self._agents[i]
        v.step = self.step
        return v

    def __len__(self):
        return self.num_agents

    # -- whole-team operations -----------------------------------------------------------------------------
    def insert(self, share_obs, obs, actions, action_log_probs, value_preds, rewards, masks, bad_masks=None,
               active_masks=None):
        """share_obs (N, S) once; everything else (N, A, .) as `MultiVecTaskPython.step` / the policies return it
        (runner.py:229-275 loops over agents and inserts share_obs[:, agent_id] each time)."""
        s = self.step
        pairs = [(self.obs, obs, s + 1), (self.actions, actions, s), (self.action_log_probs, action_log_probs, s),
                 (self.value_preds, value_preds, s), (self.rewards, rewards, s), (self.masks, masks, s + 1)]
        if bad_masks is not None:
            pairs.append((self.bad_masks, bad_masks, s + 1))
        if active_masks is not None:
            pairs.append((self.active_masks, active_masks, s + 1))
        # (N, A, .) -> agent-major slots [A][slot][N][.], share_obs once: ALL planes in one launch (`mmb_copy_group`) when they
        # are fp32 CUDA tensors with contiguous rows - nine strided torch copies otherwise.  No views are built on the way (nine
        # slicing / transpose calls cost more host time than the launch): destination addresses from the planes' strides.
        p = self.__dict__.get("_copy_params")
        if p is None:
            p = self._copy_params = L.CopyGroupParams()
        N, A = self.n_rollout_threads, self.num_agents
        ok = len(pairs) + 1 <= L.MAX_COPY_SEGS and share_obs.dtype is torch.float32 and share_obs.is_cuda and share_obs.dim() == 2 \
            and share_obs.stride(1) == 1 and self.share_obs.dtype is torch.float32
        if ok:
            g = p.seg[0]
            g.dst, g.src = self.share_obs.data_ptr() + 4 * (s + 1) * self.share_obs.stride(0), share_obs.data_ptr()
            g.n0, g.n1, g.n2 = 1, N, share_obs.shape[1]
            g.dst_s0, g.dst_s1, g.src_s0, g.src_s1 = 0, self.share_obs.stride(1), 0, share_obs.stride(0)
            for k, (dst, src, slot) in enumerate(pairs):
                w = dst.shape[3]
                if (src.dtype is not torch.float32 or dst.dtype is not torch.float32 or not src.is_cuda or src.dim() != 3 or
                        src.shape[0] != N or src.shape[1] != A or src.shape[2] != w or (w > 1 and (src.stride(2) != 1 or dst.stride(3) != 1))):
                    ok = False
                    break
                g = p.seg[k + 1]
                g.dst, g.src = dst.data_ptr() + 4 * slot * dst.stride(1), src.data_ptr()
                g.n0, g.n1, g.n2 = A, N, w
                g.dst_s0, g.dst_s1, g.src_s0, g.src_s1 = dst.stride(0), dst.stride(2), src.stride(1), src.stride(0)
        if ok:
            p.count = len(pairs) + 1
            L.check(L.lib().mmb_copy_group(p, L.stream_ptr()), "mmb_copy_group")
        else:
            self.share_obs[s + 1].copy_(share_obs)
            for dst, src, slot in pairs:
                dst[:, slot].copy_(src.transpose(0, 1))
        self.step = (self.step + 1) % self.episode_length

    def after_update(self):
        A, N = self.num_agents, self.n_rollout_threads
        self.share_obs[0].copy_(self.share_obs[-1])
        names = ("obs", "masks", "bad_masks", "active_masks")
        for n in names:
            t = getattr(self, n)
            t[:, 0].copy_(t[:, -1])

    def compute_returns(self, next_values, value_normalizers=None, advantages=True):
        """separated_buffer.py:124-168 for ALL agents in one launch.  next_values (N, A, 1) or (A, N, 1); one value
        normalizer per agent (PopArt / ValueNorm, `running_mean_var()`), or None."""
        T, N, A = self.episode_length, self.n_rollout_threads, self.num_agents
        c = self.config
        use_denorm = (c["use_popart"] or c["use_valuenorm"]) and value_normalizers is not None
        nv = next_values
        if nv.shape[0] == N and nv.shape[1] == A:
            nv = nv.transpose(0, 1)
        nv = nv.reshape(A, N).contiguous()
        p = L.GaeMarlParams()
        p.num_envs, p.num_steps, p.num_agents = N, T, A
        p.use_gae, p.use_proper_time_limits = int(c["use_gae"]), int(c["use_proper_time_limits"])
        p.use_denorm, p.use_popart = int(use_denorm), int(c["use_popart"])
        keep = [nv]
        p.rewards, p.rew_t, p.rew_e, p.rew_a = L.ptr(self.rewards), N, 1, T * N
        p.value_preds, p.val_t, p.val_e, p.val_a = L.ptr(self.value_preds), N, 1, (T + 1) * N
        p.masks, p.msk_t, p.msk_e, p.msk_a = L.ptr(self.masks), N, 1, (T + 1) * N
        p.bad_masks, p.bad_t, p.bad_e, p.bad_a = L.ptr(self.bad_masks), N, 1, (T + 1) * N
        p.next_value, p.nv_e, p.nv_a = L.ptr(nv), 1, N
        p.returns, p.ret_t, p.ret_e, p.ret_a = L.ptr(self.returns), N, 1, (T + 1) * N
        if advantages:
            self._adv_stats4.zero_()
            p.advantages, p.adv_t, p.adv_e, p.adv_a = L.ptr(self.raw_advantages), N, 1, T * N
            p.stats = L.ptr(self._adv_stats4)
        if use_denorm:
            mv = [vn.running_mean_var() for vn in value_normalizers]
            mean = torch.stack([m.reshape(-1)[0] for m, _ in mv]).float().contiguous()
            var = torch.stack([v.reshape(-1)[0] for _, v in mv]).float().contiguous()
            keep += [mean, var]
            p.denorm_mean, p.denorm_var = L.ptr(mean), L.ptr(var)
        p.gamma, p.gae_lambda = float(c["gamma"]), float(c["gae_lambda"])
        self._keep = keep
        L.check(L.lib().mmb_gae_marl(p, L.stream_ptr()), "mmb_gae_marl")

    def normalized_advantages(self, eps=1e-5):
        """mappo_trainer.py:194-199 per agent: [A, T, N, 1]."""
        adv = self.raw_advantages.clone()
        n = self.episode_length * self.n_rollout_threads
        for i in range(self.num_agents):
            L.check(L.lib().mmb_adv_normalize(L.ptr(adv[i]), n, L.ptr(self._adv_stats4[i]), eps, 0, L.stream_ptr()),
                    "mmb_adv_normalize")
        return adv

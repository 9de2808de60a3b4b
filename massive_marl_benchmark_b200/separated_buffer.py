"""`SeparatedReplayBuffer` with the reference's interface (agents/algorithms/marl/utils/
separated_buffer.py) on the B200 kernels, plus the two pieces of its callers that sit on the hot path:
the advantage prologue of the trainers (mappo_trainer.py:189-199 / happo_trainer.py:180-189) and the
mask logic of `Runner.insert` (runner.py:229-255).

Same constructor (`config, obs_space, share_obs_space, act_space, device`), same tensors
(`share_obs, obs, rnn_states, rnn_states_critic, value_preds, returns, actions, action_log_probs,
rewards, masks, bad_masks, active_masks, factor`), same methods (`insert, after_update, update_factor,
compute_returns, feed_forward_generator`) and the generator's positional 12/13-tuple
(separated_buffer.py:225-228, consumed at mappo_trainer.py:119-121).

  insert                9-12 copy_ launches            -> 1 launch (fused multi-field row copy)
  compute_returns       ~12 ops x T (Python loop)      -> 1 launch (mmb_gae_marl; PopArt/ValueNorm
                                                         denormalisation folded in as two device scalars)
  feed_forward_generator randperm + 11 gathers          -> 1 launch per minibatch (mmb_shuffle_gather)

The recurrent generators are out of scope (`use_recurrent_policy: False`, cfg/mappo/config.yaml:41-42);
the rnn-state planes are kept (zero-filled) because the reference's policies take them as arguments.
"""
import torch

from . import _lib as L


def _shape_from_space(space):
    shp = getattr(space, "shape", None)
    if shp is None:
        shp = tuple(space)
    return tuple(shp)


def _act_dim(space):
    name = space.__class__.__name__
    if name == "Discrete":
        return 1
    return int(space.shape[0])


def _multi_copy(srcs, dsts, rows):
    """One launch: for each (src, dst) pair of [rows, ...] tensors copy src -> dst."""
    p = L.GatherParams()
    p.num_fields, p.index_mode, p.total, p.batch_start, p.batch_size = len(srcs), 2, rows, 0, rows
    keep = []
    for i, (s, d) in enumerate(zip(srcs, dsts)):
        s = s if s.is_contiguous() else s.contiguous()
        if s.dtype != d.dtype:
            s = s.to(d.dtype)
        keep.append(s)
        p.src[i], p.dst[i] = s.data_ptr(), d.data_ptr()
        p.row_bytes[i] = (d.numel() // rows) * d.element_size()
    L.check(L.lib().mmb_shuffle_gather(p, L.stream_ptr()), "mmb_shuffle_gather(copy)")


class SeparatedReplayBuffer(object):
    def __init__(self, config, obs_space, share_obs_space, act_space, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.MmbError("SeparatedReplayBuffer needs a CUDA device (there is no CPU path)")
        L.lib()
        self.episode_length = config["episode_length"]
        self.n_rollout_threads = config["n_rollout_threads"]
        self.rnn_hidden_size = config["hidden_size"]
        self.recurrent_N = config["recurrent_N"]
        self.gamma = config["gamma"]
        self.gae_lambda = config["gae_lambda"]
        self._use_gae = config["use_gae"]
        self._use_popart = config["use_popart"]
        self._use_valuenorm = config["use_valuenorm"]
        self._use_proper_time_limits = config["use_proper_time_limits"]
        self.device = device
        T, N = self.episode_length, self.n_rollout_threads
        obs_shape = _shape_from_space(obs_space)
        share_obs_shape = _shape_from_space(share_obs_space)
        self.share_obs = torch.zeros(T + 1, N, *share_obs_shape, device=dev)
        self.obs = torch.zeros(T + 1, N, *obs_shape, device=dev)
        self.rnn_states = torch.zeros(T + 1, N, self.recurrent_N, self.rnn_hidden_size, device=dev)
        self.rnn_states_critic = torch.zeros_like(self.rnn_states)
        self.value_preds = torch.zeros(T + 1, N, 1, device=dev)
        self.returns = torch.zeros(T + 1, N, 1, device=dev)
        self.available_actions = None
        act_shape = _act_dim(act_space)
        self.actions = torch.zeros(T, N, act_shape, device=dev)
        self.action_log_probs = torch.zeros(T, N, act_shape, device=dev)
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.masks = torch.ones(T + 1, N, 1, device=dev)
        self.bad_masks = torch.ones_like(self.masks)
        self.active_masks = torch.ones_like(self.masks)
        self.factor = torch.ones(T, N, 1, device=dev)
        self.step = 0
        self.raw_advantages = torch.zeros(T, N, 1, device=dev)
        self._adv_stats4 = torch.zeros(4, device=dev, dtype=torch.float64)
        self.adv_stats = self._adv_stats4[:3]
        self.process_group = None
        self.permutation_override = None
        self._one = torch.ones(1, device=dev)

    def update_factor(self, factor):
        self.factor.copy_(factor)

    def insert(self, share_obs, obs, rnn_states, rnn_states_critic, actions, action_log_probs, value_preds, rewards,
               masks, bad_masks=None, active_masks=None, available_actions=None):
        s = self.step
        srcs = [share_obs, obs, actions, action_log_probs, value_preds, rewards, masks]
        dsts = [self.share_obs[s + 1], self.obs[s + 1], self.actions[s], self.action_log_probs[s], self.value_preds[s],
                self.rewards[s], self.masks[s + 1]]
        if rnn_states is not None:
            srcs += [rnn_states, rnn_states_critic]
            dsts += [self.rnn_states[s + 1], self.rnn_states_critic[s + 1]]
        if bad_masks is not None:
            srcs.append(bad_masks); dsts.append(self.bad_masks[s + 1])
        if active_masks is not None:
            srcs.append(active_masks); dsts.append(self.active_masks[s + 1])
        _multi_copy(srcs, dsts, self.n_rollout_threads)
        self.step = (self.step + 1) % self.episode_length

    def after_update(self):
        names = ("share_obs", "obs", "rnn_states", "rnn_states_critic", "masks", "bad_masks", "active_masks")
        _multi_copy([getattr(self, n)[-1] for n in names], [getattr(self, n)[0] for n in names], self.n_rollout_threads)

    def compute_returns(self, next_value, value_normalizer=None, advantages=True):
        """separated_buffer.py:124-168 (all four branches) in one launch; also leaves the raw advantages of
        mappo_trainer.py:189-192 in `raw_advantages` and their (count, sum, sumsq) in `adv_stats`."""
        T, N = self.episode_length, self.n_rollout_threads
        use_denorm = (self._use_popart or self._use_valuenorm) and value_normalizer is not None
        p = L.GaeMarlParams()
        p.num_envs, p.num_steps, p.num_agents = N, T, 1
        p.use_gae, p.use_proper_time_limits = int(self._use_gae), int(self._use_proper_time_limits)
        p.use_denorm, p.use_popart = int(use_denorm), int(self._use_popart)
        nv = next_value if next_value.is_contiguous() else next_value.contiguous()
        self._keep = [nv]
        p.rewards, p.rew_t, p.rew_e = L.ptr(self.rewards), N, 1
        p.value_preds, p.val_t, p.val_e = L.ptr(self.value_preds), N, 1
        p.masks, p.msk_t, p.msk_e = L.ptr(self.masks), N, 1
        p.bad_masks, p.bad_t, p.bad_e = L.ptr(self.bad_masks), N, 1
        p.next_value, p.nv_e = L.ptr(nv), 1
        p.returns, p.ret_t, p.ret_e = L.ptr(self.returns), N, 1
        if advantages:
            self._adv_stats4.zero_()
            p.advantages, p.adv_t, p.adv_e, p.stats = L.ptr(self.raw_advantages), N, 1, L.ptr(self._adv_stats4)
        if use_denorm:
            mean, var = value_normalizer.running_mean_var()
            mean, var = mean.reshape(-1).contiguous().float(), var.reshape(-1).contiguous().float()
            self._keep += [mean, var]
            p.denorm_mean, p.denorm_var = L.ptr(mean), L.ptr(var)
        p.gamma, p.gae_lambda = float(self.gamma), float(self.gae_lambda)
        L.check(L.lib().mmb_gae_marl(p, L.stream_ptr()), "mmb_gae_marl")

    def normalized_advantages(self, eps=1e-5):
        """mappo_trainer.py:194-199: (adv - mean) / (std + 1e-5) of the raw advantages left by compute_returns."""
        if self.process_group is not None:
            from . import dist as mdist
            mdist.all_reduce_stats(self.adv_stats, self.process_group)
        adv = self.raw_advantages.clone()
        L.check(L.lib().mmb_adv_normalize(L.ptr(adv), adv.numel(), L.ptr(self._adv_stats4), eps, 0, L.stream_ptr()),
                "mmb_adv_normalize")
        return adv

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        T, N = self.rewards.shape[0:2]
        batch_size = N * T
        if mini_batch_size is None:
            assert batch_size >= num_mini_batch, (
                "PPO requires the number of processes ({}) * number of steps ({}) = {} to be greater than or equal to "
                "the number of PPO mini batches ({}).".format(N, T, N * T, num_mini_batch))
            mini_batch_size = batch_size // num_mini_batch
        dev = self.rewards.device
        if self.permutation_override is not None:
            rand = torch.as_tensor(self.permutation_override, dtype=torch.int64)
        else:
            rand = torch.randperm(batch_size)   # CPU generator, as the reference (separated_buffer.py:183)
        rand = rand.to(dev)
        fields = [("share_obs", self.share_obs[:-1]), ("obs", self.obs[:-1]), ("rnn_states", self.rnn_states[:-1]),
                  ("rnn_states_critic", self.rnn_states_critic[:-1]), ("actions", self.actions),
                  ("value_preds", self.value_preds[:-1]), ("returns", self.returns[:-1]), ("masks", self.masks[:-1]),
                  ("active_masks", self.active_masks[:-1]), ("action_log_probs", self.action_log_probs)]
        if advantages is not None:
            fields.append(("advantages", advantages.reshape(T, N, 1).contiguous()))
        if self.factor is not None:
            fields.append(("factor", self.factor))
        for i in range(num_mini_batch):
            idx = rand[i * mini_batch_size:(i + 1) * mini_batch_size].contiguous()
            B = idx.numel()
            out = {}
            p = L.GatherParams()
            p.num_fields, p.index_mode, p.total, p.batch_start, p.batch_size = len(fields), 0, batch_size, 0, B
            p.indices = L.ptr(idx)
            for j, (name, src) in enumerate(fields):
                out[name] = torch.empty((B,) + tuple(src.shape[2:]), device=dev)
                p.src[j], p.dst[j] = src.data_ptr(), out[name].data_ptr()
                p.row_bytes[j] = int(src[0, 0].numel()) * src.element_size()
            L.check(L.lib().mmb_shuffle_gather(p, L.stream_ptr()), "mmb_shuffle_gather")
            tup = (out["share_obs"], out["obs"], out["rnn_states"], out["rnn_states_critic"], out["actions"],
                   out["value_preds"], out["returns"], out["masks"], out["active_masks"], out["action_log_probs"],
                   out.get("advantages"), None)
            if self.factor is None:
                yield tup
            else:
                yield tup + (out["factor"],)

    def naive_recurrent_generator(self, advantages, num_mini_batch):
        raise NotImplementedError("recurrent generators are out of scope (use_recurrent_policy: False)")

    def recurrent_generator(self, advantages, num_mini_batch, data_chunk_length):
        raise NotImplementedError("recurrent generators are out of scope (use_recurrent_policy: False)")


def runner_insert_masks(dones, masks_out=None, active_masks_out=None):
    """Runner.insert mask logic (runner.py:229-255) in one launch: dones (N,A) int64 -> masks, active_masks
    (N,A,1) fp32.  The outputs may be slices of buffer planes (strided over agents)."""
    N, A = dones.shape
    dev = dones.device
    d = dones if (dones.dtype == torch.int64 and dones.is_contiguous()) else dones.to(torch.int64).contiguous()
    if masks_out is None:
        masks_out = torch.empty(N, A, 1, device=dev)
    if active_masks_out is None:
        active_masks_out = torch.empty(N, A, 1, device=dev)
    L.check(L.lib().mmb_marl_masks(L.ptr(d), N, A, L.ptr(masks_out), masks_out.stride(0), masks_out.stride(1),
                                   L.ptr(active_masks_out), active_masks_out.stride(0), active_masks_out.stride(1),
                                   L.stream_ptr()), "mmb_marl_masks")
    return masks_out, active_masks_out

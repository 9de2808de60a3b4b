"""One PPO update of a whole MAPPO / IPPO team, grouped and data-parallel.

The reference updates the agents one after the other (agents/algorithms/marl/runner.py:257-317): for each agent its trainer's
`train(buffer)` runs `ppo_epoch` x `num_mini_batch` minibatch updates (mappo_trainer.py:172-230 / ippo_trainer.py), each with
two backward passes, two `clip_grad_norm_` and two `Adam.step()` on that agent's own actor and critic: with ten agents that
is 20 optimisers stepped in turn.  In MAPPO and IPPO the agents share nothing during the update (own networks, own buffers,
own value normaliser), so the order of (agent, epoch, minibatch) is immaterial, and this module runs the update
minibatch-major:

    for epoch, minibatch k:
        opt.zero_grad()                                          one memset for the 20 networks
        for agent a:  losses(a, minibatch k).backward()          the reference's modules + the fused loss kernel
                      opt.collect_grads(groups of a)             one multi-tensor copy
                      reducer.reduce_async(slice of a)           NCCL all-reduce of agent a's 5 MB, overlapped with agent a+1
        reducer.wait(); opt.step()                               clip + Adam for all 20 networks: two launches

* `GroupedAdam` keeps every parameter / gradient / moment of the team in four flat buffers; the gradient buffer is the
  all-reduce buffer (no flatten / unflatten).
* Env-sharded data parallel (BASELINE configs[3]): every rank owns a slice of the envs and its own buffers; gradients are
  averaged per agent while the next agent computes; the advantage statistics and the value normalisers' batch moments are
  averaged too, so all replicas take bit-identical steps (`replica_checksum`).
* HAPPO's sequential scheme (each agent's surrogate carries the probability ratios of the agents updated before it,
  happo_trainer.py:135-141) cannot be reordered: use the per-agent drop-in `mappo_update.happo_ppo_update`.

Equivalence with the reference on one GPU (same minibatch permutations): tests/test_gpu_team_update.py.
"""
import torch

from . import _lib as L
from . import dist as mdist
from .grouped_adam import GroupedAdam
from .mappo_update import evaluate_losses


class TeamUpdate:
    def __init__(self, trainers, buffers, config, algorithm="mappo", group=None, data_parallel=None):
        if algorithm not in ("mappo", "ippo"):
            raise NotImplementedError("TeamUpdate reorders the per-agent updates, which is only valid for mappo / ippo (got %r)" % algorithm)
        self.trainers, self.buffers, self.config, self.algorithm = list(trainers), list(buffers), config, algorithm
        self.A = len(self.trainers)
        self.opt = GroupedAdam.for_marl_policies([t.policy for t in self.trainers], config)
        import torch.distributed as dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if data_parallel is None:
            data_parallel = self.world > 1
        self.dp = data_parallel and self.world > 1
        self.group = group
        self.reducer = mdist.FlatGradReducer(self.opt.flat_grads, group) if self.dp else None
        if self.dp:
            sync = lambda t: mdist.mean_over_ranks(t, group)          # noqa: E731
            for t in self.trainers:
                t.moment_sync = sync
        self.ppo_epoch, self.num_mini_batch = config["ppo_epoch"], config["num_mini_batch"]
        self.last_grad_norms = None

    # ------------------------------------------------------------------------------------------------------------
    def advantages(self):
        """mappo_trainer.py:189-199 per agent: returns - denormalised value predictions, normalised with the mean and the
        unbiased std over the WHOLE batch - all env shards when data parallel (the raw advantages and their fp64 count / sum
        / sum of squares were left by the buffers' compute_returns)."""
        stats = torch.stack([b._adv_stats4 for b in self.buffers])                # [A, 4] fp64 (views for a SharedReplayBuffer)
        if self.dp:
            import torch.distributed as dist
            s3 = stats[:, :3].contiguous()
            dist.all_reduce(s3, op=dist.ReduceOp.SUM, group=self.group)
            stats = torch.cat([s3, stats[:, 3:]], dim=1)
        out = []
        for a, b in enumerate(self.buffers):
            adv = b.raw_advantages.clone()
            st = stats[a].contiguous()
            L.check(L.lib().mmb_adv_normalize(L.ptr(adv), adv.numel(), L.ptr(st), 1e-5, 0, L.stream_ptr()), "mmb_adv_normalize")
            out.append(adv)
        return out

    def train(self):
        """One update of every agent from its buffer (whose `compute_returns` has run).  Returns the per-agent train_info
        dicts of `<Algo>.train` (value_loss, policy_loss, dist_entropy, actor_grad_norm, critic_grad_norm, ratio)."""
        ippo = self.algorithm == "ippo"
        for t in self.trainers:
            t.prep_training()
        advs = self.advantages()
        T, N = self.buffers[0].episode_length, self.buffers[0].n_rollout_threads
        factor = torch.ones(T, N, 1, device=self.opt.device)
        for b in self.buffers:
            b.update_factor(factor)
        sums = torch.zeros(self.A, 6, dtype=torch.float64, device=self.opt.device)
        for _epoch in range(self.ppo_epoch):
            gens = [b.feed_forward_generator(advs[a], self.num_mini_batch) for a, b in enumerate(self.buffers)]
            for _k in range(self.num_mini_batch):
                self.opt.zero_grad()
                for a in range(self.A):
                    tr = self.trainers[a]
                    out = evaluate_losses(tr, next(gens[a]), ippo=ippo)
                    (out.policy_loss - out.dist_entropy * tr.entropy_coef).backward()        # mappo_trainer.py:146
                    (out.value_loss * tr.value_loss_coef).backward()                          # mappo_trainer.py:162
                    self.opt.collect_grads((2 * a, 2 * a + 1))
                    if self.reducer is not None:
                        lo, _ = self.opt.group_slice(2 * a)
                        _, hi = self.opt.group_slice(2 * a + 1)
                        self.reducer.reduce_async(lo, hi)
                    sums[a, 0] += out.value_loss.detach().double()
                    sums[a, 1] += out.policy_loss.detach().double()
                    sums[a, 2] += out.dist_entropy.detach().double()
                    sums[a, 3] += out.imp_weights.mean().double()
                if self.reducer is not None:
                    self.reducer.wait()
                self.opt.step()
                norms = self.opt.grad_norms().view(self.A, 2)
                sums[:, 4] += norms[:, 0]
                sums[:, 5] += norms[:, 1]
        for b in self.buffers:
            b.after_update()
        n = self.ppo_epoch * self.num_mini_batch
        s = (sums / n).cpu()                                                                  # the update's only host read-back
        self.last_grad_norms = s[:, 4:6]
        return [dict(value_loss=float(s[a, 0]), policy_loss=float(s[a, 1]), dist_entropy=float(s[a, 2]), ratio=float(s[a, 3]),
                     actor_grad_norm=float(s[a, 4]), critic_grad_norm=float(s[a, 5])) for a in range(self.A)]

    def replica_checksum(self):
        """(max - min) over ranks of the sum and of the sum of squares of all parameters (fp64): (0, 0) when every replica
        holds bit-identical parameters.  Host sync + two small collectives; for tests and the bench."""
        v = torch.stack([self.opt.flat_params.double().sum(), (self.opt.flat_params.double() ** 2).sum()])
        if self.world == 1:
            return 0.0, 0.0
        import torch.distributed as dist
        hi, lo = v.clone(), v.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        d = (hi - lo).cpu()
        return float(d[0]), float(d[1])

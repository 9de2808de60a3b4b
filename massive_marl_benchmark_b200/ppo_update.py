"""Drop-in body for the reference's `PPO.update` (agents/algorithms/rl/ppo/ppo.py:243-317): the same minibatch loop with
the storage-side work and the loss on this library's kernels,

    batch = storage.mini_batch_generator(n)        device-side shuffle (`mmb_permutation`), CUDA index tensors
    storage.gather_minibatch(indices)              the nine gathers of ppo.py:253-264 in one launch (`mmb_shuffle_gather`)
    actor_critic.actor / .critic                   the reference's torch modules (autograd: their backward is torch's)
    ppo_loss(...)                                  distribution + KL + surrogate + value loss + their backward, one launch
    loss.backward(); clip_grad_norm_; optimizer.step()      unchanged

and one host read-back per minibatch (the adaptive-KL decision, which the reference also makes on the host) instead of
three.  Use:  `PPO.update = ppo_update`  (or call `ppo_update(ppo)`); `ppo` needs the attributes `PPO.__init__` sets
(ppo.py:30-97): storage, actor_critic, optimizer, num_mini_batches, num_learning_epochs, clip_param, value_loss_coef,
entropy_coef, use_clipped_value_loss, desired_kl, schedule, step_size, max_grad_norm, asymmetric.

Env-sharded data parallel (one process per GPU, each with its own envs and storage): set `ppo.grad_sync =
dist.all_reduce_grads` and `ppo.scalar_sync = dist.mean_over_ranks`.  Gradients are averaged after `backward()` and before
the clipping, the KL estimate before the step-size decision, so every rank takes the same step and the result equals the
single-process update on the concatenated minibatches (equal shard sizes; tests/test_dist_gloo.py, world 2).
"""
import torch
import torch.nn as nn

from .ppo_loss import ppo_loss


def _gather(storage, indices):
    if hasattr(storage, "gather_minibatch"):
        return storage.gather_minibatch(indices)
    flat = lambda t: t.view(-1, *t.size()[2:])              # noqa: E731  (a reference-style storage: ppo.py:253-264)
    names = ("observations", "states", "actions", "values", "returns", "actions_log_prob", "advantages", "mu", "sigma")
    return {n: flat(getattr(storage, n))[indices] for n in names if getattr(storage, n).numel()}


def ppo_update(self):
    st, ac = self.storage, self.actor_critic
    value_loss_sum = torch.zeros((), dtype=torch.float64, device=ac.log_std.device)
    surrogate_loss_sum = torch.zeros_like(value_loss_sum)
    adaptive = self.desired_kl is not None and self.schedule == "adaptive"
    grad_sync, scalar_sync = getattr(self, "grad_sync", None), getattr(self, "scalar_sync", None)

    batch = st.mini_batch_generator(self.num_mini_batches)
    for _epoch in range(self.num_learning_epochs):
        for indices in batch:
            mb = _gather(st, indices)
            obs_batch = mb["observations"]
            critic_in = mb["states"] if self.asymmetric else obs_batch
            mu_batch = ac.actor(obs_batch)                                        # module.py:93
            value_batch = ac.critic(critic_in)                                    # module.py:101-104
            out = ppo_loss(mu_batch, ac.log_std, value_batch, mb["actions"], mb["actions_log_prob"], mb["advantages"],
                           mb["values"], mb["returns"], mb["mu"] if adaptive else None, mb["sigma"] if adaptive else None,
                           clip_param=self.clip_param, value_loss_coef=self.value_loss_coef,
                           entropy_coef=self.entropy_coef, use_clipped_value_loss=self.use_clipped_value_loss)

            if adaptive:                                                          # ppo.py:270-283
                kl_mean = float(out.kl_mean if scalar_sync is None else scalar_sync(out.kl_mean))
                if kl_mean > self.desired_kl * 2.0:
                    self.step_size = max(1e-5, self.step_size / 1.5)
                elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
                    self.step_size = min(1e-2, self.step_size * 1.5)
                for param_group in self.optimizer.param_groups:
                    param_group["lr"] = self.step_size

            self.optimizer.zero_grad()                                            # ppo.py:305-308
            out.loss.backward()
            if grad_sync is not None:
                grad_sync(list(ac.parameters()))
            nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)
            self.optimizer.step()

            value_loss_sum += out.value_loss.double()                             # ppo.py:310-311, read back once at the end
            surrogate_loss_sum += out.surrogate_loss.double()

    num_updates = self.num_learning_epochs * self.num_mini_batches
    sums = torch.stack([value_loss_sum, surrogate_loss_sum])
    if scalar_sync is not None:
        sums = scalar_sync(sums)
    sums = sums.cpu()
    return float(sums[0]) / num_updates, float(sums[1]) / num_updates

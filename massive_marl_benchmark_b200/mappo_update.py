"""Drop-in body for the reference's `MAPPO.ppo_update` (agents/algorithms/marl/mappo_trainer.py:106-172) for the
configuration the benchmark ships (feed-forward policy, Box actions, cfg/mappo/config.yaml): the same sequence with the
distribution, both losses and their backward on one kernel,

    actor.base -> act.action_out.fc_mean, critic.base -> v_out      the reference's torch modules (autograd)
    value_normalizer(return_batch) x 2                              the reference's own PopArt, called as often as the
                                                                    reference calls it (each call updates the statistics)
    mappo_loss(...)                                                 log-probs, importance weights, surrogate, value loss,
                                                                    gradients: one launch (`mmb_mappo_loss`)
    (policy_loss - dist_entropy * entropy_coef).backward(); clip; actor_optimizer.step()        unchanged
    (value_loss * value_loss_coef).backward(); clip; critic_optimizer.step()                    unchanged

Use:  `MAPPO.ppo_update = mappo_ppo_update`, `HAPPO.ppo_update = happo_ppo_update` (agents/algorithms/marl/happo_trainer.py:
MAPPO's update with the factor of the previously updated agents inside the surrogate; it is positive, so it folds into the
advantage), `IPPO.ppo_update = ippo_ppo_update` (agents/algorithms/marl/ippo_trainer.py:
the same update; its value normaliser - ValueNorm in cfg/ippo/config.yaml - is updated once and both error terms are
normalised with the same moments, ippo_trainer.py:74-77).  Recurrent policies, discrete actions and `available_actions`
are outside the benchmark's configurations and raise.
"""
import torch
import torch.nn as nn

from .mappo_loss import mappo_loss


def _grad_norm(parameters):                                  # agents/utils/util.py get_gard_norm
    total = 0.0
    for p in parameters:
        if p.grad is not None:
            total += float(p.grad.norm()) ** 2
    return total ** 0.5


def _t(x, like):
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=like.device, dtype=torch.float32)


def mappo_ppo_update(self, sample, update_actor=True):
    return _ppo_update(self, sample, update_actor)


def ippo_ppo_update(self, sample, update_actor=True):
    return _ppo_update(self, sample, update_actor, ippo=True)


def happo_ppo_update(self, sample, update_actor=True):
    return _ppo_update(self, sample, update_actor, happo=True)


def _norm_input(self, return_batch):
    """What the value normaliser's running statistics are updated with.  Single process: the minibatch returns, as in the
    reference.  Env-sharded data parallel (`self.moment_sync` set, e.g. `dist.mean_over_ranks`): PopArt / ValueNorm update
    from the batch mean and mean square only (popart.py:46-57, valuenorm.py:44-55), so every rank feeds the reference's own
    update a two-element surrogate [mu + s, mu - s] carrying the moments averaged over ALL shards - the replicas' running
    statistics (and PopArt's output-layer rescaling) stay bit-identical without touching the reference's classes."""
    sync = getattr(self, "moment_sync", None)
    if sync is None:
        return return_batch
    m = sync(torch.stack([return_batch.mean(), (return_batch * return_batch).mean()]))
    sd = (m[1] - m[0] * m[0]).clamp(min=0.0).sqrt()
    return torch.stack([m[0] + sd, m[0] - sd]).reshape(2, 1)


def evaluate_losses(self, sample, ippo=False, happo=False):
    """Forward of one agent's actor and critic on a minibatch + the fused loss kernel: everything of `ppo_update` up to the
    two `backward()` calls.  Returns the `MappoLossOut`."""
    (share_obs_batch, obs_batch, _rnn_a, _rnn_c, actions_batch, value_preds_batch, return_batch, _masks_batch,
     active_masks_batch, old_action_log_probs_batch, adv_targ, available_actions_batch, factor_batch) = sample
    actor, critic = self.policy.actor, self.policy.critic
    if getattr(self, "_use_recurrent_policy", False) or getattr(self, "_use_naive_recurrent", False):
        raise NotImplementedError("recurrent policies are outside the benchmark's configurations")
    if available_actions_batch is not None:
        raise NotImplementedError("available_actions (discrete action masks) are outside the benchmark's configurations")
    head = actor.act.action_out                                                  # DiagGaussian, distributions.py:94-117
    ref = head.log_std
    obs_batch, share_obs_batch = _t(obs_batch, ref), _t(share_obs_batch, ref)
    return_batch = _t(return_batch, ref)

    mean = head.fc_mean(actor.base(obs_batch))                                   # actor_critic.py:95, distributions.py:115
    std = torch.sigmoid(head.log_std / head.std_x_coef) * head.std_y_coef        # distributions.py:116
    values = critic.v_out(critic.base(share_obs_batch))                          # actor_critic.py:163-166

    moments = [None, None, None, None]
    norm_in = _norm_input(self, return_batch)
    if ippo:
        if self._use_popart or self._use_valuenorm:                              # ippo_trainer.py:74-77: one update, one pair
            self.value_normalizer.update(norm_in)
            m, v = self.value_normalizer.running_mean_var()
            moments = [m.clone(), v.clone(), None, None]
    elif getattr(self, "_use_valuenorm", False):                                 # mappo_trainer.py:75-78: the statistics are
        self.value_normalizer.update(norm_in)                                    # updated, but the errors normalised there are
                                                                                 # overwritten by the else branch at :83-85
    if self._use_popart and not ippo:                                            # mappo_trainer.py:80-82: two training-mode calls,
        self.value_normalizer(norm_in)                                           # the first normalises the clipped error,
        m1, v1 = self.value_normalizer.running_mean_var()
        self.value_normalizer(norm_in)                                           # the second the original one
        m2, v2 = self.value_normalizer.running_mean_var()
        moments = [m1.clone(), v1.clone(), m2.clone(), v2.clone()]

    adv_targ = _t(adv_targ, ref)
    if happo:
        # happo_trainer.py:135-141: -sum_j factor[b, j] * min(surr1, surr2)[b].  The factor is a product of probability ratios,
        # hence positive, and min(w * a * f, clamp(w) * a * f) = f * min(w * a, clamp(w) * a) for f > 0: the row factor
        # folds into the advantage and the same kernel serves
        adv_targ = adv_targ * _t(factor_batch, ref).sum(dim=-1, keepdim=True)

    out = mappo_loss(mean, std, values, _t(actions_batch, ref), _t(old_action_log_probs_batch, ref), adv_targ,
                     _t(value_preds_batch, ref), return_batch,
                     None if active_masks_batch is None else _t(active_masks_batch, ref), *moments,
                     clip_param=self.clip_param, huber_delta=self.huber_delta, use_huber_loss=self._use_huber_loss,
                     use_clipped_value_loss=self._use_clipped_value_loss,
                     use_value_active_masks=self._use_value_active_masks,
                     use_policy_active_masks=self._use_policy_active_masks)
    return out


def _ppo_update(self, sample, update_actor, ippo=False, happo=False):
    out = evaluate_losses(self, sample, ippo, happo)
    actor, critic = self.policy.actor, self.policy.critic

    # actor update, mappo_trainer.py:143-153
    self.policy.actor_optimizer.zero_grad()
    if update_actor:
        (out.policy_loss - out.dist_entropy * self.entropy_coef).backward()
    if self._use_max_grad_norm:
        actor_grad_norm = nn.utils.clip_grad_norm_(actor.parameters(), self.max_grad_norm)
    else:
        actor_grad_norm = _grad_norm(actor.parameters())
    self.policy.actor_optimizer.step()

    # critic update, mappo_trainer.py:155-170
    self.policy.critic_optimizer.zero_grad()
    (out.value_loss * self.value_loss_coef).backward()
    if self._use_max_grad_norm:
        critic_grad_norm = nn.utils.clip_grad_norm_(critic.parameters(), self.max_grad_norm)
    else:
        critic_grad_norm = _grad_norm(critic.parameters())
    self.policy.critic_optimizer.step()

    return out.value_loss, critic_grad_norm, out.policy_loss, out.dist_entropy, actor_grad_norm, out.imp_weights

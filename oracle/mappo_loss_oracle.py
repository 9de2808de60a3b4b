"""TEST INFRASTRUCTURE - CPU restatement of the MAPPO minibatch losses of the reference (forward + autograd gradients).

Follows, statement by statement,
  DiagGaussian.forward / FixedNormal      agents/algorithms/utils/distributions.py:94-117, 32-43
                                          (std = sigmoid(log_std / std_x_coef) * std_y_coef, per-dimension log-probs)
  ACTLayer.evaluate_actions (Box branch)  agents/algorithms/utils/act.py:154-165 (entropy: masked mean or plain mean)
  MAPPO.ppo_update                        agents/algorithms/marl/mappo_trainer.py:127-146 (importance weights, clipped
                                          surrogate, optional active masks), :164-168 (critic loss)
  MAPPO.cal_value_loss                    mappo_trainer.py:62-103 (clipped value prediction, PopArt / ValueNorm-normalised
                                          returns, huber or mse, optional active masks)
  huber_loss / mse_loss                   agents/utils/util.py:23-29 - the reference's huber has NO branch for e < -d
                                          (`b = (e > d)`): such errors contribute zero loss and zero gradient.  Kept.
with the actor mean [B,A], log_std [A] and the critic value [B,1] as the differentiable inputs.  The reference calls its
PopArt normaliser once per error term (mappo_trainer.py:80-81) and every call first folds the batch into the running
statistics (popart.py:38-57), so the clipped error is normalised with the moments after ONE update and the original
error with the moments after TWO.  `popart_update` restates that update; the debiased moments are inputs here:
(`ret_mean`, `ret_var`) for the clipped term, (`ret_mean_orig`, `ret_var_orig`) for the original one (None = the same
pair); `ret_mean` None = returns used as they are - which is also what the reference computes with ValueNorm: its
normalised errors (mappo_trainer.py:75-78) are overwritten by the `else` branch of the PopArt test that follows (:83-85).
Pinned: tests/test_oracle_vs_reference.py runs the reference's own `MAPPO.ppo_update` with its own `ACTLayer` and `PopArt`
(oracle/ref_mappo_loss.py) on the same inputs and requires identical losses and gradients; tests/golden/mappo_loss.npz
holds outputs of that reference run.  Only tests/, smoke() and bench.py's cpu_baseline leg may import this module.
"""
import torch


def huber_loss(e, d):                                    # agents/utils/util.py:23-26
    a = (abs(e) <= d).float()
    b = (e > d).float()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


def mse_loss(e):                                         # agents/utils/util.py:28-29
    return e ** 2 / 2


def popart_update(state, batch, beta=0.99999, epsilon=1e-5):
    """One training-mode call of PopArt.forward on `batch` [B,1] (popart.py:38-57 with norm_axes=1,
    per_element_update=False): folds the batch moments into `state` = dict(running_mean, running_mean_sq, debiasing_term)
    IN PLACE and returns the debiased (mean, var) the call then normalises with (popart.py:30-34)."""
    detached = batch.detach()
    batch_mean = detached.mean(dim=(0,))
    batch_sq_mean = (detached ** 2).mean(dim=(0,))
    weight = beta
    state["running_mean"].mul_(weight).add_(batch_mean * (1.0 - weight))
    state["running_mean_sq"].mul_(weight).add_(batch_sq_mean * (1.0 - weight))
    state["debiasing_term"].mul_(weight).add_(1.0 * (1.0 - weight))
    debiased_mean = state["running_mean"] / state["debiasing_term"].clamp(min=epsilon)
    debiased_mean_sq = state["running_mean_sq"] / state["debiasing_term"].clamp(min=epsilon)
    debiased_var = (debiased_mean_sq - debiased_mean ** 2).clamp(min=1e-2)
    return debiased_mean, debiased_var


def mappo_loss_terms(mean, log_std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks,
                     ret_mean=None, ret_var=None, ret_mean_orig=None, ret_var_orig=None, clip_param=0.2, huber_delta=10.0,
                     use_huber_loss=True, use_clipped_value_loss=True, use_value_active_masks=False,
                     use_policy_active_masks=False, std_x_coef=1.0, std_y_coef=0.5, factor=None):
    """The statements themselves, on whatever graph `mean`, `log_std`, `values` belong to.  Returns
    (policy_loss, dist_entropy, value_loss, imp_weights, action_log_probs).  `factor` [B,A]: HAPPO's product of the
    previously updated agents' probability ratios, multiplied into the surrogate before the sum over the last dimension
    (happo_trainer.py:135-141); None = MAPPO / IPPO."""
    action_std = torch.sigmoid(log_std / std_x_coef) * std_y_coef                        # distributions.py:116
    dist = torch.distributions.Normal(mean, action_std)                                  # FixedNormal
    action_log_probs = dist.log_prob(actions)                                            # distributions.py:34
    if use_policy_active_masks:                                                          # act.py:157-163 (Box branch)
        dist_entropy = (dist.entropy() * active_masks).sum() / active_masks.sum()
    else:
        dist_entropy = dist.entropy().mean()

    imp_weights = torch.exp((action_log_probs - old_logp).sum(dim=-1, keepdim=True))     # mappo_trainer.py:128
    surr1 = imp_weights * adv_targ
    surr2 = torch.clamp(imp_weights, 1.0 - clip_param, 1.0 + clip_param) * adv_targ
    if factor is not None:                                                               # happo_trainer.py:135-141
        if use_policy_active_masks:
            policy_action_loss = (-torch.sum(factor * torch.min(surr1, surr2), dim=-1, keepdim=True) * active_masks).sum() / active_masks.sum()
        else:
            policy_action_loss = -torch.sum(factor * torch.min(surr1, surr2), dim=-1, keepdim=True).mean()
    elif use_policy_active_masks:
        policy_action_loss = (-torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True) * active_masks).sum() / active_masks.sum()
    else:
        policy_action_loss = -torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True).mean()
    policy_loss = policy_action_loss

    # cal_value_loss, mappo_trainer.py:73-103
    value_pred_clipped = value_preds + (values - value_preds).clamp(-clip_param, clip_param)
    if ret_mean is not None:
        ret_c = (returns - ret_mean) / torch.sqrt(ret_var)                               # popart.py:59-60, first call
        if ret_mean_orig is not None:
            ret_o = (returns - ret_mean_orig) / torch.sqrt(ret_var_orig)                 # second call, moments moved on
        else:
            ret_o = ret_c
    else:
        ret_c = ret_o = returns
    error_clipped = ret_c - value_pred_clipped
    error_original = ret_o - values
    if use_huber_loss:
        value_loss_clipped = huber_loss(error_clipped, huber_delta)
        value_loss_original = huber_loss(error_original, huber_delta)
    else:
        value_loss_clipped = mse_loss(error_clipped)
        value_loss_original = mse_loss(error_original)
    if use_clipped_value_loss:
        value_loss = torch.max(value_loss_original, value_loss_clipped)
    else:
        value_loss = value_loss_original
    if use_value_active_masks:
        value_loss = (value_loss * active_masks).sum() / active_masks.sum()
    else:
        value_loss = value_loss.mean()
    return policy_loss, dist_entropy, value_loss, imp_weights, action_log_probs


def mappo_loss_oracle(mean, log_std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks,
                      ret_mean=None, ret_var=None, ret_mean_orig=None, ret_var_orig=None, clip_param=0.2,
                      value_loss_coef=1.0, entropy_coef=0.0, huber_delta=10.0,
                      use_huber_loss=True, use_clipped_value_loss=True, use_value_active_masks=False,
                      use_policy_active_masks=False, std_x_coef=1.0, std_y_coef=0.5, popart_running_mean=None,
                      popart_running_mean_sq=None, popart_debiasing_term=None):
    """mean [B,A], log_std [A], values [B,1] are differentiated.  (The `popart_*` entries `synthetic_minibatch` carries along
    for the reference runner - the normaliser's state BEFORE the two calls - are not used here.)
    old_logp [B,A] (per dimension), adv_targ / value_preds /
    returns / active_masks [B,1].  Returns detached: policy_loss, dist_entropy, value_loss, imp_weights [B,1], logp [B,A],
    grad_mean, grad_log_std (of policy_loss - dist_entropy * entropy_coef), grad_values (of value_loss * value_loss_coef)."""
    mean = mean.detach().clone().requires_grad_(True)
    log_std = log_std.detach().clone().requires_grad_(True)
    values = values.detach().clone().requires_grad_(True)
    policy_loss, dist_entropy, value_loss, imp_weights, action_log_probs = mappo_loss_terms(
        mean, log_std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks, ret_mean, ret_var,
        ret_mean_orig, ret_var_orig, clip_param, huber_delta, use_huber_loss, use_clipped_value_loss, use_value_active_masks,
        use_policy_active_masks, std_x_coef, std_y_coef)
    g_mean, g_ls = torch.autograd.grad(policy_loss - dist_entropy * entropy_coef, (mean, log_std))   # mappo_trainer.py:146
    g_v, = torch.autograd.grad(value_loss * value_loss_coef, (values,))                  # mappo_trainer.py:168
    return {"policy_loss": policy_loss.detach(), "dist_entropy": dist_entropy.detach(), "value_loss": value_loss.detach(),
            "imp_weights": imp_weights.detach(), "logp": action_log_probs.detach(),
            "grad_mean": g_mean, "grad_log_std": g_ls, "grad_values": g_v}


def mappo_update_oracle(tr, sample, update_actor=True, ippo=False, happo=False):
    """`MAPPO.ppo_update` (mappo_trainer.py:106-172) for the feed-forward Box-action policy on an object `tr` with the
    trainer's attributes (policy.actor / .critic / .actor_optimizer / .critic_optimizer, clip_param, value_loss_coef,
    entropy_coef, max_grad_norm, huber_delta, the _use_* flags) and `tr.popart` = the PopArt state dict (see
    `popart_update`) or None.  `sample` is the generator's tuple (separated_buffer.py:225-228).  Returns
    (value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, imp_weights).
    `ippo=True`: `IPPO.ppo_update` (ippo_trainer.py:101-170) - the same update except that its normaliser (ValueNorm, whose
    update is arithmetically PopArt's, valuenorm.py:39-55) is updated ONCE and both error terms share the moments
    (ippo_trainer.py:74-77); `tr.popart` then holds the ValueNorm state and `tr._use_valuenorm` says whether it is on.
    `happo=True`: `HAPPO.ppo_update` (happo_trainer.py:93-170) - MAPPO's update with the sample's 13th entry, the factor
    [B,A], inside the surrogate."""
    (share_obs, obs, _ra, _rc, actions, value_preds, returns, _masks, active_masks, old_logp, adv_targ, _avail, factor) = sample
    actor, critic = tr.policy.actor, tr.policy.critic
    head = actor.act.action_out
    mean = head.fc_mean(actor.base(obs))                                         # actor_critic.py:95, distributions.py:115
    values = critic.v_out(critic.base(share_obs))                                # actor_critic.py:163-166
    moments = (None, None, None, None)
    if ippo:
        if tr._use_popart or tr._use_valuenorm:
            m, v = popart_update(tr.popart, returns)
            moments = (m.clone(), v.clone(), None, None)
    elif tr._use_popart:                                                         # two calls, two updates (see the header)
        m1, v1 = popart_update(tr.popart, returns)
        m1, v1 = m1.clone(), v1.clone()
        m2, v2 = popart_update(tr.popart, returns)
        moments = (m1, v1, m2.clone(), v2.clone())
    policy_loss, dist_entropy, value_loss, imp_weights, _ = mappo_loss_terms(
        mean, head.log_std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks, *moments,
        clip_param=tr.clip_param, huber_delta=tr.huber_delta, use_huber_loss=tr._use_huber_loss,
        use_clipped_value_loss=tr._use_clipped_value_loss, use_value_active_masks=tr._use_value_active_masks,
        use_policy_active_masks=tr._use_policy_active_masks, std_x_coef=head.std_x_coef, std_y_coef=head.std_y_coef,
        factor=factor if happo else None)
    tr.policy.actor_optimizer.zero_grad()                                        # mappo_trainer.py:143-153
    if update_actor:
        (policy_loss - dist_entropy * tr.entropy_coef).backward()
    actor_grad_norm = torch.nn.utils.clip_grad_norm_(actor.parameters(), tr.max_grad_norm)
    tr.policy.actor_optimizer.step()
    tr.policy.critic_optimizer.zero_grad()                                       # mappo_trainer.py:164-170
    (value_loss * tr.value_loss_coef).backward()
    critic_grad_norm = torch.nn.utils.clip_grad_norm_(critic.parameters(), tr.max_grad_norm)
    tr.policy.critic_optimizer.step()
    return value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, imp_weights


def without_popart(mb):
    """The same minibatch for a trainer without value normalisation."""
    return {k: (None if k.startswith(("ret_", "popart_")) else v) for k, v in mb.items()}


def synthetic_minibatch(B, A, seed, spread=0.2, huber_delta=10.0):
    """Seeded minibatch: importance weights around 1 on both sides of the clip range, advantages of both signs with exact
    zeros, inactive rows, values inside / outside the value clip range with exact ties, errors beyond +/- huber_delta."""
    g = torch.Generator().manual_seed(seed)
    log_std = 1.0 + 0.3 * torch.randn(A, generator=g)
    std = torch.sigmoid(log_std) * 0.5
    old_mean = 0.3 * torch.randn(B, A, generator=g)
    actions = old_mean + std * torch.randn(B, A, generator=g)
    mean = old_mean + spread * std * torch.randn(B, A, generator=g) / (A ** 0.5)
    old_std = torch.sigmoid(log_std + 0.01 * torch.randn(A, generator=g)) * 0.5
    old_logp = torch.distributions.Normal(old_mean, old_std).log_prob(actions)
    adv_targ = torch.randn(B, 1, generator=g)
    adv_targ[::7] = 0.0
    active_masks = (torch.rand(B, 1, generator=g) > 0.2).float()
    value_preds = torch.randn(B, 1, generator=g)
    values = value_preds + 0.3 * torch.randn(B, 1, generator=g)
    values[::5] = value_preds[::5]
    returns = 2.0 + 3.0 * torch.randn(B, 1, generator=g)
    scale = max(huber_delta, 1.0)
    returns[1::11] += 4.0 * scale                         # error > delta  (linear branch)
    returns[2::11] -= 4.0 * scale                         # error < -delta (the reference's zero branch)
    # a PopArt state a few hundred updates into training (debiasing term ~ 1 - beta^n): the moments still move visibly
    deb = torch.tensor(1.0 - 0.99999 ** 300)
    popart = dict(popart_running_mean=torch.tensor([1.7]) * deb, popart_running_mean_sq=torch.tensor([6.5 + 1.7 ** 2]) * deb,
                  popart_debiasing_term=deb)
    st = {k[len("popart_"):]: v.clone() for k, v in popart.items()}
    m1, v1 = popart_update(st, returns)
    m2, v2 = popart_update(st, returns)
    return dict(mean=mean, log_std=log_std, values=values, actions=actions, old_logp=old_logp, adv_targ=adv_targ,
                value_preds=value_preds, returns=returns, active_masks=active_masks, ret_mean=m1, ret_var=v1,
                ret_mean_orig=m2, ret_var_orig=v2, **popart)

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
with the actor mean [B,A], log_std [A] and the critic value [B,1] as the differentiable inputs.  The running PopArt
statistics are updated by the reference BEFORE the normalisation (popart.py:38-57); here the already-updated debiased
(mean, var) are inputs (`ret_mean`, `ret_var`), None = returns used as they are.
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


def mappo_loss_oracle(mean, log_std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks,
                      ret_mean=None, ret_var=None, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, huber_delta=10.0,
                      use_huber_loss=True, use_clipped_value_loss=True, use_value_active_masks=False,
                      use_policy_active_masks=False, std_x_coef=1.0, std_y_coef=0.5):
    """mean [B,A], log_std [A], values [B,1] are differentiated.  old_logp [B,A] (per dimension), adv_targ / value_preds /
    returns / active_masks [B,1].  Returns detached: policy_loss, dist_entropy, value_loss, imp_weights [B,1], logp [B,A],
    grad_mean, grad_log_std (of policy_loss - dist_entropy * entropy_coef), grad_values (of value_loss * value_loss_coef)."""
    mean = mean.detach().clone().requires_grad_(True)
    log_std = log_std.detach().clone().requires_grad_(True)
    values = values.detach().clone().requires_grad_(True)

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
    if use_policy_active_masks:
        policy_action_loss = (-torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True) * active_masks).sum() / active_masks.sum()
    else:
        policy_action_loss = -torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True).mean()
    policy_loss = policy_action_loss
    g_mean, g_ls = torch.autograd.grad(policy_loss - dist_entropy * entropy_coef, (mean, log_std))   # mappo_trainer.py:146

    # cal_value_loss, mappo_trainer.py:73-103
    value_pred_clipped = value_preds + (values - value_preds).clamp(-clip_param, clip_param)
    if ret_mean is not None:
        ret_n = (returns - ret_mean) / torch.sqrt(ret_var)                               # popart.py:59-60
    else:
        ret_n = returns
    error_clipped = ret_n - value_pred_clipped
    error_original = ret_n - values
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
    g_v, = torch.autograd.grad(value_loss * value_loss_coef, (values,))                  # mappo_trainer.py:168

    return {"policy_loss": policy_loss.detach(), "dist_entropy": dist_entropy.detach(), "value_loss": value_loss.detach(),
            "imp_weights": imp_weights.detach(), "logp": action_log_probs.detach(),
            "grad_mean": g_mean, "grad_log_std": g_ls, "grad_values": g_v}


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
    ret_mean = torch.tensor([1.7])
    ret_var = torch.tensor([6.5])
    return dict(mean=mean, log_std=log_std, values=values, actions=actions, old_logp=old_logp, adv_targ=adv_targ,
                value_preds=value_preds, returns=returns, active_masks=active_masks, ret_mean=ret_mean, ret_var=ret_var)

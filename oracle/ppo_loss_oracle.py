"""TEST INFRASTRUCTURE - CPU restatement of the PPO minibatch loss of the reference (forward + autograd gradients).

Follows, statement by statement,
  ActorCritic.evaluate            agents/algorithms/rl/ppo/module.py:92-107  (MultivariateNormal with
                                  scale_tril = diag(exp(log_std)^2): the effective standard deviation is sigma^2)
  PPO.update, KL / surrogate /    agents/algorithms/rl/ppo/ppo.py:268-302
  value loss / total loss
with the actor mean and the critic value as the differentiable inputs (what the two MLPs hand over).  The gradients are
torch autograd's on exactly these statements.  `ppo_update_oracle` is the whole minibatch loop of `PPO.update`
(ppo.py:243-317: gathers, evaluate, loss, adaptive-KL step size, backward, gradient clipping, optimiser step), pinned by
running the reference's own unmodified `PPO.update` on the reference's `ActorCritic` and `RolloutStorage`.
Pinned: tests/test_oracle_vs_reference.py runs the reference's own `ActorCritic.evaluate` (imported) and the reference's
own source lines of `PPO.update` (textually extracted) on the same inputs and requires identical losses and gradients;
tests/golden/ppo_loss.npz holds outputs of that reference run (tests/golden/make_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import torch
from torch.distributions import MultivariateNormal


def ppo_loss_terms(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                   clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True):
    """The statements themselves, on whatever graph `mu`, `log_std`, `value` belong to.  Returns
    (loss, surrogate_loss, value_loss, kl_mean, actions_log_prob, entropy)."""
    # module.py:95-99
    covariance = torch.diag(log_std.exp() * log_std.exp())
    distribution = MultivariateNormal(mu, scale_tril=covariance)
    actions_log_prob = distribution.log_prob(actions)
    entropy = distribution.entropy()
    sigma = log_std.repeat(mu.shape[0], 1)                                        # module.py:107

    # ppo.py:271-273
    kl = torch.sum(sigma - old_sigma + (torch.square(old_sigma.exp()) + torch.square(old_mu - mu))
                   / (2.0 * torch.square(sigma.exp())) - 0.5, axis=-1)
    kl_mean = torch.mean(kl)

    # ppo.py:284-288
    ratio = torch.exp(actions_log_prob - torch.squeeze(old_logp))
    surrogate = -torch.squeeze(advantages) * ratio
    surrogate_clipped = -torch.squeeze(advantages) * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()

    # ppo.py:291-300
    if use_clipped_value_loss:
        value_clipped = target_values + (value - target_values).clamp(-clip_param, clip_param)
        value_losses = (value - returns).pow(2)
        value_losses_clipped = (value_clipped - returns).pow(2)
        value_loss = torch.max(value_losses, value_losses_clipped).mean()
    else:
        value_loss = (returns - value).pow(2).mean()

    loss = surrogate_loss + value_loss_coef * value_loss - entropy_coef * entropy.mean()      # ppo.py:302
    return loss, surrogate_loss, value_loss, kl_mean, actions_log_prob, entropy


def ppo_loss_oracle(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                    clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True):
    """mu [B,A], log_std [A], value [B,1] are differentiated; the rest are the minibatch rows gathered from the storage
    (`old_logp`, `advantages`, `target_values`, `returns` [B,1]; `old_mu`, `old_sigma` [B,A]).  Returns a dict of
    detached tensors: loss, surrogate_loss, value_loss, kl_mean, logp [B], entropy (scalar), grad_mu, grad_log_std,
    grad_value."""
    mu = mu.detach().clone().requires_grad_(True)
    log_std = log_std.detach().clone().requires_grad_(True)
    value = value.detach().clone().requires_grad_(True)
    loss, surrogate_loss, value_loss, kl_mean, actions_log_prob, entropy = ppo_loss_terms(
        mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma, clip_param,
        value_loss_coef, entropy_coef, use_clipped_value_loss)
    g_mu, g_ls, g_v = torch.autograd.grad(loss, (mu, log_std, value))
    return {"loss": loss.detach(), "surrogate_loss": surrogate_loss.detach(), "value_loss": value_loss.detach(),
            "kl_mean": kl_mean.detach(), "logp": actions_log_prob.detach(), "entropy": entropy.detach()[0],
            "grad_mu": g_mu, "grad_log_std": g_ls, "grad_value": g_v}


def ppo_update_oracle(ppo, epoch_orders):
    """`PPO.update` (ppo.py:243-317) on an object with the reference's attributes (`storage` with the [T, N, .] tensors,
    `actor_critic` with `.actor`, `.critic`, `.log_std`, `optimizer`, `num_mini_batches`, `num_learning_epochs`,
    `clip_param`, `value_loss_coef`, `entropy_coef`, `use_clipped_value_loss`, `desired_kl`, `schedule`, `step_size`,
    `max_grad_norm`, `asymmetric`).  `epoch_orders[e]` is the flat index order of epoch e (what the storage's sampler
    yields: arange for 'sequential'), cut into `num_mini_batches` consecutive minibatches, the remainder dropped.
    Returns (mean_value_loss, mean_surrogate_loss)."""
    st, ac = ppo.storage, ppo.actor_critic
    flat = lambda t: t.view(-1, *t.size()[2:])              # noqa: E731
    mb_size = (st.num_envs * st.num_transitions_per_env) // ppo.num_mini_batches
    mean_value_loss = mean_surrogate_loss = 0.0
    for e in range(ppo.num_learning_epochs):
        order = epoch_orders[e]
        for b in range(ppo.num_mini_batches):
            idx = order[b * mb_size:(b + 1) * mb_size]
            obs = flat(st.observations)[idx]
            critic_in = flat(st.states)[idx] if ppo.asymmetric else obs
            mu, value = ac.actor(obs), ac.critic(critic_in)                       # module.py:93,101-104
            loss, s_loss, v_loss, kl_mean, _, _ = ppo_loss_terms(
                mu, ac.log_std, value, flat(st.actions)[idx], flat(st.actions_log_prob)[idx], flat(st.advantages)[idx],
                flat(st.values)[idx], flat(st.returns)[idx], flat(st.mu)[idx], flat(st.sigma)[idx], ppo.clip_param,
                ppo.value_loss_coef, ppo.entropy_coef, ppo.use_clipped_value_loss)
            if ppo.desired_kl is not None and ppo.schedule == "adaptive":         # ppo.py:270-283
                if kl_mean > ppo.desired_kl * 2.0:
                    ppo.step_size = max(1e-5, ppo.step_size / 1.5)
                elif kl_mean < ppo.desired_kl / 2.0 and kl_mean > 0.0:
                    ppo.step_size = min(1e-2, ppo.step_size * 1.5)
                for group in ppo.optimizer.param_groups:
                    group["lr"] = ppo.step_size
            ppo.optimizer.zero_grad()                                             # ppo.py:305-308
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ac.parameters(), ppo.max_grad_norm)
            ppo.optimizer.step()
            mean_value_loss += v_loss.item()
            mean_surrogate_loss += s_loss.item()
    n = ppo.num_learning_epochs * ppo.num_mini_batches
    return mean_value_loss / n, mean_surrogate_loss / n


def synthetic_minibatch(B, A, seed, ratio_spread=0.15, value_spread=0.3):
    """Seeded minibatch in the ranges a PPO update sees: new/old policy close (ratios around 1 with some rows outside
    the clip range on both sides), advantages of both signs incl. exact zeros, values inside and outside the value
    clip range."""
    g = torch.Generator().manual_seed(seed)
    log_std = torch.randn(A, generator=g) * 0.2 - 0.3
    std = log_std.exp() * log_std.exp()
    old_mu = torch.randn(B, A, generator=g) * 0.5
    actions = old_mu + std * torch.randn(B, A, generator=g)
    mu = old_mu + ratio_spread * std * torch.randn(B, A, generator=g) / (A ** 0.5)
    old_sigma = (log_std + 0.02 * torch.randn(A, generator=g)).repeat(B, 1)
    ostd = old_sigma[0].exp() * old_sigma[0].exp()
    old_logp = (-0.5 * (((actions - old_mu) / ostd) ** 2).sum(-1) - ostd.log().sum() - 0.5 * A * 1.8378770664093453).view(B, 1)
    advantages = torch.randn(B, 1, generator=g)
    advantages[::7] = 0.0
    target_values = torch.randn(B, 1, generator=g)
    value = target_values + value_spread * torch.randn(B, 1, generator=g)
    value[::5] = target_values[::5]                       # exact ties of the two value losses
    returns = target_values + 0.5 * torch.randn(B, 1, generator=g)
    return dict(mu=mu, log_std=log_std, value=value, actions=actions, old_logp=old_logp, advantages=advantages,
                target_values=target_values, returns=returns, old_mu=old_mu, old_sigma=old_sigma)

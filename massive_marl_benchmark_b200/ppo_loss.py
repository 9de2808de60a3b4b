"""The PPO minibatch loss of the reference's `PPO.update` (agents/algorithms/rl/ppo/ppo.py:266-302) as ONE kernel that
produces the loss terms and the gradients torch's autograd would (`mmb_ppo_loss`), wrapped in an `autograd.Function` so
that `loss.backward()` continues into the two MLPs exactly as in the reference:

    mu = actor(obs_batch); value = critic(obs_batch)                      # reference modules (autograd)
    out = ppo_loss(mu, actor_critic.log_std, value, actions_batch, old_actions_log_prob_batch, advantages_batch,
                   target_values_batch, returns_batch, old_mu_batch, old_sigma_batch, clip_param=..., ...)
    out.loss.backward()                                                    # ppo.py:306
    # out.kl_mean drives the adaptive learning-rate schedule (ppo.py:277-283); out.surrogate_loss / out.value_loss are
    # the logged terms (ppo.py:310-311)

Replaces ActorCritic.evaluate's distribution part (module.py:95-99,107: the `scale_tril = diag(exp(log_std)^2)` quirk is
kept) and about forty elementwise / reduction kernels of the forward + backward pass.  No CPU path: raises if the CUDA
library is missing.
"""
from collections import namedtuple

import torch

from . import _lib as L

PpoLossOut = namedtuple("PpoLossOut", "loss surrogate_loss value_loss kl_mean logp entropy")


def _rows(t, B, name):
    t = t.detach()
    if t.numel() != B:
        raise ValueError("%s: expected %d elements, got %s" % (name, B, tuple(t.shape)))
    return t.reshape(B).float().contiguous()


def ppo_loss_raw(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu=None, old_sigma=None,
                 clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True, need_grads=True):
    """One `mmb_ppo_loss` launch.  Returns (sums [4 + A] fp64: {sum surrogate, sum value loss, sum kl, entropy} followed by
    d loss / d log_std, logp [B], grad_mu [B, A] | None, grad_value [B] | None)."""
    if not mu.is_cuda:
        raise L.MmbError("ppo_loss needs CUDA tensors (there is no CPU path)")
    B, A = mu.shape
    mu_ = mu.detach().float()
    if mu_.stride(1) != 1:
        mu_ = mu_.contiguous()
    keep = [mu_]
    p = L.PpoLossParams()
    p.num_rows, p.act_dim, p.use_clipped_value_loss = B, A, int(bool(use_clipped_value_loss))
    p.mu, p.mu_stride = mu_.data_ptr(), mu_.stride(0)

    def put(field, t):
        keep.append(t)
        setattr(p, field, t.data_ptr())

    put("log_std", log_std.detach().reshape(A).float().contiguous())
    act = actions.detach().float().contiguous()
    if tuple(act.shape) != (B, A):
        raise ValueError("actions: expected %s, got %s" % ((B, A), tuple(act.shape)))
    put("actions", act)
    put("old_logp", _rows(old_logp, B, "old_logp"))
    put("advantages", _rows(advantages, B, "advantages"))
    put("value", _rows(value, B, "value"))
    put("returns", _rows(returns, B, "returns"))
    if use_clipped_value_loss:
        put("target_values", _rows(target_values, B, "target_values"))
    if (old_mu is None) != (old_sigma is None):
        raise ValueError("old_mu and old_sigma go together")
    if old_mu is not None:
        put("old_mu", old_mu.detach().float().reshape(B, A).contiguous())
        put("old_sigma", old_sigma.detach().float().reshape(B, A).contiguous())
    p.clip_param = float(clip_param)
    p.ratio_lo, p.ratio_hi = 1.0 - clip_param, 1.0 + clip_param     # rounded to fp32 like the reference's Python scalars
    p.value_loss_coef, p.entropy_coef = float(value_loss_coef), float(entropy_coef)
    dev = mu.device
    sums = torch.zeros(4 + A, dtype=torch.float64, device=dev)
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    p.sums, p.logp = sums.data_ptr(), logp.data_ptr()
    grad_mu = grad_value = None
    if need_grads:
        grad_mu = torch.empty(B, A, dtype=torch.float32, device=dev)
        grad_value = torch.empty(B, dtype=torch.float32, device=dev)
        p.grad_mu, p.grad_value = grad_mu.data_ptr(), grad_value.data_ptr()
    L.check(L.lib().mmb_ppo_loss(p, L.stream_ptr()), "mmb_ppo_loss")
    return sums, logp, grad_mu, grad_value


class _FusedPPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss):
        B, A = mu.shape
        sums, logp, grad_mu, grad_value = ppo_loss_raw(mu, log_std, value, actions, old_logp, advantages, target_values,
                                                       returns, old_mu, old_sigma, clip_param, value_loss_coef,
                                                       entropy_coef, use_clipped_value_loss)
        # three small launches instead of a dozen (the wrapper, not the 26 us kernel, was the cost of this op):
        # fp32 copy of the sums; loss = <sums[:4], (1/B, c_v/B, 0, -c_e)> (ppo.py:302); the three reported means
        s32 = sums.float()
        w = _loss_weights(mu.device, B, float(value_loss_coef), float(entropy_coef))
        loss = torch.dot(s32[:4], w)
        means = s32[:3] / B
        ctx.save_for_backward(grad_mu, s32[4:].reshape(log_std.shape), grad_value.reshape(value.shape))
        ctx.mark_non_differentiable(logp)
        return loss, means[0], means[1], means[2], logp, s32[3]

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        grads = torch._foreach_mul(list(ctx.saved_tensors), g_loss)       # one multi-tensor launch
        return tuple(grads) + (None,) * 11


_WEIGHTS = {}


def _loss_weights(device, B, value_loss_coef, entropy_coef):
    key = (str(device), B, value_loss_coef, entropy_coef)
    w = _WEIGHTS.get(key)
    if w is None:
        if len(_WEIGHTS) > 64:
            _WEIGHTS.clear()
        w = _WEIGHTS[key] = torch.tensor([1.0 / B, value_loss_coef / B, 0.0, -entropy_coef], dtype=torch.float32, device=device)
    return w


def ppo_loss(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu=None, old_sigma=None,
             clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True):
    """Differentiable with respect to `mu` [B, A], `log_std` [A] and `value` [B, 1] through `.loss` (gradients of the other
    returned terms are not provided: the reference only differentiates the total).  The remaining arguments are the
    minibatch rows as `PPO.update` gathers them ([B, 1] or [B]; `old_mu`, `old_sigma` [B, A], both None to skip the KL
    estimate).  Returns PpoLossOut(loss, surrogate_loss, value_loss, kl_mean, logp [B], entropy)."""
    out = _FusedPPOLoss.apply(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                              clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss)
    return PpoLossOut(*out)

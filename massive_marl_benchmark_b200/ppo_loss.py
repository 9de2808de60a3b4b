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


def _f32(t):
    return t if (t.dtype is torch.float32 and t.is_contiguous()) else t.detach().float().contiguous()


def _rows(t, B, name):
    if t.numel() != B:
        raise ValueError("%s: expected %d elements, got %s" % (name, B, tuple(t.shape)))
    return _f32(t)


def _fill(p, keep, mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
          clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss):
    """Inputs of one `mmb_ppo_loss` launch into the params struct `p`; tensors that had to be converted go to `keep`."""
    if not mu.is_cuda:
        raise L.MmbError("ppo_loss needs CUDA tensors (there is no CPU path)")
    B, A = mu.shape
    mu_ = mu if (mu.dtype is torch.float32 and mu.stride(1) == 1) else mu.detach().float().contiguous()
    p.num_rows, p.act_dim, p.use_clipped_value_loss = B, A, int(bool(use_clipped_value_loss))
    p.mu, p.mu_stride = mu_.data_ptr(), mu_.stride(0)
    keep.append(mu_)

    def put(field, t):
        keep.append(t)
        setattr(p, field, t.data_ptr())

    put("log_std", _rows(log_std, A, "log_std"))
    if tuple(actions.shape) != (B, A):
        raise ValueError("actions: expected %s, got %s" % ((B, A), tuple(actions.shape)))
    put("actions", _f32(actions))
    put("old_logp", _rows(old_logp, B, "old_logp"))
    put("advantages", _rows(advantages, B, "advantages"))
    put("value", _rows(value, B, "value"))
    put("returns", _rows(returns, B, "returns"))
    if use_clipped_value_loss:
        put("target_values", _rows(target_values, B, "target_values"))
    else:
        p.target_values = None
    if (old_mu is None) != (old_sigma is None):
        raise ValueError("old_mu and old_sigma go together")
    if old_mu is not None:
        put("old_mu", _rows(old_mu, B * A, "old_mu"))
        put("old_sigma", _rows(old_sigma, B * A, "old_sigma"))
    else:
        p.old_mu = p.old_sigma = None
    p.clip_param = float(clip_param)
    p.ratio_lo, p.ratio_hi = 1.0 - clip_param, 1.0 + clip_param     # rounded to fp32 like the reference's Python scalars
    p.value_loss_coef, p.entropy_coef = float(value_loss_coef), float(entropy_coef)
    return B, A


def ppo_loss_raw(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu=None, old_sigma=None,
                 clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True, need_grads=True):
    """One `mmb_ppo_loss` launch.  Returns (sums [4 + A] fp64: {sum surrogate, sum value loss, sum kl, entropy} followed by
    d loss / d log_std, logp [B], grad_mu [B, A] | None, grad_value [B] | None)."""
    p, keep = L.PpoLossParams(), []
    B, A = _fill(p, keep, mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                 clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss)
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


_SCRATCH = {}     # (device, stream, A) -> (params struct, fp64 sums + ticket word): the kernel hands the scratch back zeroed


def _scratch(dev, stream, A):
    key = (dev.index, stream, A)
    e = _SCRATCH.get(key)
    if e is None:
        if len(_SCRATCH) > 64:
            _SCRATCH.clear()
        buf = torch.zeros(4 + A + 1, dtype=torch.float64, device=dev)
        p = L.PpoLossParams()
        p.sums, p.ticket = buf.data_ptr(), buf.data_ptr() + 8 * (4 + A)
        e = _SCRATCH[key] = (p, buf)
    return e[0]


class _FusedPPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss):
        # ONE launch: the kernel's last block finalises the fp32 terms (loss = mean surrogate + c_v mean value loss -
        # c_e entropy, ppo.py:302; the three reported means; d loss / d log_std) into `out`; everything below is views.
        st = L.stream_ptr()
        A = mu.shape[1]
        p = _scratch(mu.device, st, A)
        keep = []
        B, A = _fill(p, keep, mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                     clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss)
        dev = mu.device
        # gradients and finalised terms share one buffer, so that backward scales all of them with ONE multiply
        o_val = (B * A + 3) & ~3
        o_out = (o_val + B + 3) & ~3
        flat = torch.empty(o_out + 5 + A, dtype=torch.float32, device=dev)
        logp = torch.empty(B, dtype=torch.float32, device=dev)
        base = flat.data_ptr()
        p.grad_mu, p.grad_value, p.out, p.logp = base, base + 4 * o_val, base + 4 * o_out, logp.data_ptr()
        L.check(L.lib().mmb_ppo_loss(p, st), "mmb_ppo_loss")
        loss, surrogate, value_loss, kl_mean, entropy = flat[o_out:o_out + 5].unbind(0)
        ctx.save_for_backward(flat)
        ctx.geom = (B, A, o_val, o_out, log_std.shape, value.shape)
        ctx.mark_non_differentiable(logp, surrogate, value_loss, kl_mean, entropy)   # only the total is differentiated (ppo.py:306)
        return loss, surrogate, value_loss, kl_mean, logp, entropy

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        B, A, o_val, o_out, ls_shape, v_shape = ctx.geom
        g = ctx.saved_tensors[0] * g_loss
        return (g[:B * A].view(B, A), g[o_out + 5:].view(ls_shape), g[o_val:o_val + B].view(v_shape)) + (None,) * 11


def ppo_loss(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu=None, old_sigma=None,
             clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True):
    """Differentiable with respect to `mu` [B, A], `log_std` [A] and `value` [B, 1] through `.loss` (gradients of the other
    returned terms are not provided: the reference only differentiates the total).  The remaining arguments are the
    minibatch rows as `PPO.update` gathers them ([B, 1] or [B]; `old_mu`, `old_sigma` [B, A], both None to skip the KL
    estimate).  Returns PpoLossOut(loss, surrogate_loss, value_loss, kl_mean, logp [B], entropy)."""
    out = _FusedPPOLoss.apply(mu, log_std, value, actions, old_logp, advantages, target_values, returns, old_mu, old_sigma,
                              clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss)
    return PpoLossOut(*out)

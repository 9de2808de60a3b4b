"""The MAPPO minibatch losses of the reference's `MAPPO.ppo_update` / `cal_value_loss`
(agents/algorithms/marl/mappo_trainer.py:62-103,127-168) for one agent as ONE kernel that produces both loss values and
the gradients autograd would (`mmb_mappo_loss`), attached to the graph so that the reference's two backward calls keep
working unchanged:

    mean = actor_trunk_and_head(obs_batch); values = critic(share_obs_batch)          # reference modules (autograd)
    std = torch.sigmoid(act.log_std / std_x_coef) * std_y_coef                         # distributions.py:116
    out = mappo_loss(mean, std, values, actions_batch, old_action_log_probs_batch, adv_targ, value_preds_batch,
                     return_batch, active_masks_batch, ret_mean, ret_var, clip_param=..., huber_delta=..., ...)
    (out.policy_loss - out.dist_entropy * entropy_coef).backward()                     # mappo_trainer.py:146
    (out.value_loss * value_loss_coef).backward()                                      # mappo_trainer.py:168

`ret_mean`, `ret_var` are the PopArt / ValueNorm moments the CLIPPED error term is normalised with, `ret_mean_orig`,
`ret_var_orig` those of the unclipped one (None = the same pair; all None = raw returns): the reference calls PopArt once
per error term and each call first updates the running statistics (mappo_trainer.py:80-81, popart.py:38-60), so with
PopArt pass `running_mean_var()` after the first and after the second call (`mappo_update.mappo_ppo_update` does).
The entropy depends on `std` only and is computed here with plain torch ops (so its gradient reaches `log_std` through
autograd).  No CPU path: raises if the CUDA library is missing.
"""
import math
from collections import namedtuple

import torch

from . import _lib as L

MappoLossOut = namedtuple("MappoLossOut", "policy_loss value_loss dist_entropy imp_weights logp")


class _Attach(torch.autograd.Function):
    """Returns `value` (a scalar computed by the kernel) as a function of `inputs` whose gradients are the precomputed
    `grads`: backward multiplies them by the incoming gradient."""

    @staticmethod
    def forward(ctx, value, n, *args):
        ctx.n = n
        ctx.save_for_backward(*args[n:])
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        return (None, None) + tuple(g * t for t in ctx.saved_tensors) + (None,) * ctx.n


def _rows(t, B, name):
    t = t.detach()
    if t.numel() != B:
        raise ValueError("%s: expected %d elements, got %s" % (name, B, tuple(t.shape)))
    return t.reshape(B).float().contiguous()


_SCRATCH = {}     # (device, stream, A) -> fp64 sums + ticket word: the finalising kernel hands the scratch back zeroed


def _scratch(dev, stream, A):
    key = (dev.index, stream, A)
    buf = _SCRATCH.get(key)
    if buf is None:
        if len(_SCRATCH) > 64:
            _SCRATCH.clear()
        buf = _SCRATCH[key] = torch.zeros(2 + A + 1, dtype=torch.float64, device=dev)
    return buf


def mappo_loss_raw(mean, std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks=None, ret_mean=None,
                   ret_var=None, ret_mean_orig=None, ret_var_orig=None, clip_param=0.2, huber_delta=10.0, use_huber_loss=True, use_clipped_value_loss=True,
                   use_value_active_masks=False, use_policy_active_masks=False, finalise=True):
    """One `mmb_mappo_loss` launch.  Returns (policy_loss, value_loss [0-dim fp32 tensors], imp_weights [B, 1], logp [B, A],
    grad_mean [B, A], grad_std [A], grad_values [B]).  finalise=True (default): the kernel's last block computes the two
    fp32 loss values and the fp32 std gradient itself (include/mmb.h, `out`) - one launch, everything returned is a view of
    its outputs; False: the fp64 sums are reduced by torch ops afterwards (the first version, kept for the parity test)."""
    if not mean.is_cuda:
        raise L.MmbError("mappo_loss needs CUDA tensors (there is no CPU path)")
    B, A = mean.shape
    dev = mean.device
    keep = []
    p = L.MappoLossParams()

    def put(field, t):
        keep.append(t)
        setattr(p, field, t.data_ptr())

    mean_ = mean.detach().float()
    if mean_.stride(1) != 1:
        mean_ = mean_.contiguous()
    put("mean", mean_)
    p.num_rows, p.act_dim, p.mean_stride = B, A, mean_.stride(0)
    p.use_huber_loss, p.use_clipped_value_loss = int(bool(use_huber_loss)), int(bool(use_clipped_value_loss))
    p.use_value_active_masks, p.use_policy_active_masks = int(bool(use_value_active_masks)), int(bool(use_policy_active_masks))
    put("std", std.detach().reshape(A).float().contiguous())
    for name, t in (("actions", actions), ("old_logp", old_logp)):
        t = t.detach().float().contiguous()
        if tuple(t.shape) != (B, A):
            raise ValueError("%s: expected %s, got %s" % (name, (B, A), tuple(t.shape)))
        put(name, t)
    put("adv_targ", _rows(adv_targ, B, "adv_targ"))
    put("values", _rows(values, B, "values"))
    put("value_preds", _rows(value_preds, B, "value_preds"))
    put("returns", _rows(returns, B, "returns"))
    den = float(B)
    mask_sum = None
    if use_value_active_masks or use_policy_active_masks:
        if active_masks is None:
            raise ValueError("active_masks is required with use_*_active_masks")
        am = _rows(active_masks, B, "active_masks")
        mask_sum = am.sum().reshape(1)                            # the reference's own denominator, active_masks.sum()
        put("active_masks", am)
        put("mask_sum", mask_sum)
    if (ret_mean is None) != (ret_var is None):
        raise ValueError("ret_mean and ret_var go together")
    if (ret_mean_orig is None) != (ret_var_orig is None) or (ret_mean_orig is not None and ret_mean is None):
        raise ValueError("ret_mean_orig and ret_var_orig go together, and with ret_mean / ret_var")
    for name, t in (("ret_mean", ret_mean), ("ret_var", ret_var), ("ret_mean_orig", ret_mean_orig), ("ret_var_orig", ret_var_orig)):
        if t is not None:
            put(name, t.detach().reshape(-1)[:1].float().contiguous().to(dev))
    p.clip_param, p.huber_delta = float(clip_param), float(huber_delta)
    p.ratio_lo, p.ratio_hi = 1.0 - clip_param, 1.0 + clip_param
    imp = torch.empty(B, 1, dtype=torch.float32, device=dev)
    logp = torch.empty(B, A, dtype=torch.float32, device=dev)
    p.imp_weights, p.logp = imp.data_ptr(), logp.data_ptr()
    st = L.stream_ptr()
    if finalise:
        o_val = (B * A + 3) & ~3
        o_out = (o_val + B + 3) & ~3
        flat = torch.empty(o_out + 2 + A, dtype=torch.float32, device=dev)     # gradients and finalised terms: one allocation
        base = flat.data_ptr()
        scratch = _scratch(dev, st, A)
        p.sums, p.ticket = scratch.data_ptr(), scratch.data_ptr() + 8 * (2 + A)
        p.grad_mean, p.grad_values, p.out = base, base + 4 * o_val, base + 4 * o_out
        L.check(L.lib().mmb_mappo_loss(p, st), "mmb_mappo_loss")
        return (flat[o_out], flat[o_out + 1], imp, logp, flat[:B * A].view(B, A), flat[o_out + 2:], flat[o_val:o_val + B])
    sums = torch.zeros(2 + A, dtype=torch.float64, device=dev)
    grad_mean = torch.empty(B, A, dtype=torch.float32, device=dev)
    grad_values = torch.empty(B, dtype=torch.float32, device=dev)
    p.sums = sums.data_ptr()
    p.grad_mean, p.grad_values = grad_mean.data_ptr(), grad_values.data_ptr()
    L.check(L.lib().mmb_mappo_loss(p, st), "mmb_mappo_loss")
    den_p = mask_sum[0].double() if use_policy_active_masks else den
    den_v = mask_sum[0].double() if use_value_active_masks else den
    return ((sums[0] / den_p).float(), (sums[1] / den_v).float(), imp, logp, grad_mean, sums[2:].float(), grad_values)


def mappo_loss(mean, std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks=None, ret_mean=None,
               ret_var=None, ret_mean_orig=None, ret_var_orig=None, clip_param=0.2, huber_delta=10.0, use_huber_loss=True, use_clipped_value_loss=True,
               use_value_active_masks=False, use_policy_active_masks=False):
    """`policy_loss` is differentiable with respect to `mean` [B, A] and `std` [A], `value_loss` with respect to `values`
    [B, 1], `dist_entropy` with respect to `std`; they can be back-propagated separately, in any order, as the reference
    does.  Returns MappoLossOut(policy_loss, value_loss, dist_entropy, imp_weights [B, 1], logp [B, A])."""
    pl, vl, imp, logp, g_mean, g_std, g_values = mappo_loss_raw(
        mean, std, values, actions, old_logp, adv_targ, value_preds, returns, active_masks, ret_mean, ret_var, ret_mean_orig,
        ret_var_orig, clip_param, huber_delta, use_huber_loss, use_clipped_value_loss, use_value_active_masks, use_policy_active_masks)
    policy_loss = _Attach.apply(pl, 2, mean, std, g_mean, g_std.reshape(std.shape))
    value_loss = _Attach.apply(vl, 1, values, g_values.reshape(values.shape))
    # Normal.entropy = 0.5 + 0.5 log(2 pi) + log(scale) per dimension, the same on every row; act.py:157-163: masked mean
    # over rows (= sum over dimensions) with the policy active masks, else the mean over all B x A elements
    ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std.reshape(-1).float())
    dist_entropy = ent.sum() if use_policy_active_masks else ent.mean()
    return MappoLossOut(policy_loss, value_loss, dist_entropy, imp, logp)

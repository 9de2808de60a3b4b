"""GPU: the fused MAPPO minibatch losses (`mmb_mappo_loss` through massive_marl_benchmark_b200.mappo_loss) against the
oracle (oracle/mappo_loss_oracle.py: the reference's statements + torch autograd on the CPU, pinned bit for bit against
the reference's own `MAPPO.ppo_update`) and the golden fixture generated from the reference itself.

Tolerance: 1e-5 relative to the tensor's scale for the value side and the per-dimension log-probs.  The importance weight
is exp of a sum of A log-prob differences, each carrying up to an ulp (~2.4e-7) of rounding that becomes RELATIVE error
of the weight: quantities downstream of it are compared at max(1e-5, 2.5e-7 * A)."""
import pytest
import torch

from conftest import load_golden
from test_oracle_golden import MAPPO_BOOLS, _loss_case

pytestmark = pytest.mark.gpu

BASE = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, huber_delta=10.0, use_huber_loss=True,
            use_clipped_value_loss=True, use_value_active_masks=False, use_policy_active_masks=False, std_x_coef=1.0,
            std_y_coef=0.5)


def _close(a, b, rel, what, floor=0.0):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    scale = float(b.abs().max()) + 1e-12
    err = float((a - b).abs().max())
    assert err <= rel * scale + floor, "%s: max err %.3e at scale %.3e (rel %.1e)" % (what, err, scale, rel)


def _check(dev, mb, cfg, want):
    from massive_marl_benchmark_b200.mappo_loss import mappo_loss
    d = {k: (v.to(dev) if v is not None else None) for k, v in mb.items()}
    mean = d["mean"].clone().requires_grad_(True)
    log_std = d["log_std"].clone().requires_grad_(True)
    values = d["values"].clone().requires_grad_(True)
    std = torch.sigmoid(log_std / cfg["std_x_coef"]) * cfg["std_y_coef"]                 # distributions.py:116
    out = mappo_loss(mean, std, values, d["actions"], d["old_logp"], d["adv_targ"], d["value_preds"], d["returns"],
                     d["active_masks"], d["ret_mean"], d["ret_var"], d.get("ret_mean_orig"), d.get("ret_var_orig"),
                     clip_param=cfg["clip_param"],
                     huber_delta=cfg["huber_delta"], use_huber_loss=cfg["use_huber_loss"],
                     use_clipped_value_loss=cfg["use_clipped_value_loss"],
                     use_value_active_masks=cfg["use_value_active_masks"],
                     use_policy_active_masks=cfg["use_policy_active_masks"])
    (out.policy_loss - out.dist_entropy * cfg["entropy_coef"]).backward()                # mappo_trainer.py:146
    (out.value_loss * cfg["value_loss_coef"]).backward()                                 # mappo_trainer.py:168
    A = mean.shape[1]
    rr = max(1e-5, 2.5e-7 * A)
    _close(out.logp, want["logp"], 1e-5, "logp")
    _close(out.dist_entropy, want["dist_entropy"], 1e-5, "dist_entropy")
    _close(out.value_loss, want["value_loss"], 1e-5, "value_loss")
    _close(values.grad, want["grad_values"], 1e-5, "grad_values")
    _close(out.imp_weights, want["imp_weights"], rr, "imp_weights")
    # the policy loss is a mean of signed terms that largely cancel: the error scales with the terms, not with the mean
    term_scale = float((want["imp_weights"] * mb["adv_targ"]).abs().mean())
    _close(out.policy_loss, want["policy_loss"], rr, "policy_loss", floor=rr * term_scale)
    _close(mean.grad, want["grad_mean"], rr, "grad_mean")
    _close(log_std.grad, want["grad_log_std"], rr, "grad_log_std")
    assert out.imp_weights.shape == want["imp_weights"].shape and values.grad.shape == values.shape
    return out


def test_mappo_loss_against_the_reference_fixture(cuda_device):
    g = load_golden("mappo_loss")
    for tag in ("cfg", "mask", "mse"):
        mb, cfg, want = _loss_case(g, tag, MAPPO_BOOLS)
        _check(cuda_device, mb, cfg, want)


@pytest.mark.parametrize("B,A", [(1, 8), (33, 6), (1000, 8), (777, 16), (513, 80), (300, 200), (16384, 8)])
def test_mappo_loss_against_the_oracle(cuda_device, B, A):
    """Ragged row counts, every lane-group width / columns-per-lane case, the trainer's flag combinations."""
    from oracle.mappo_loss_oracle import mappo_loss_oracle, synthetic_minibatch, without_popart
    for i, over in enumerate((dict(), dict(use_value_active_masks=True, use_policy_active_masks=True, entropy_coef=0.01,
                                           huber_delta=1.0, clip_param=0.1),
                              dict(use_huber_loss=False, use_clipped_value_loss=False, value_loss_coef=0.5, popart=False),
                              dict(use_policy_active_masks=True, use_huber_loss=False))):
        cfg = dict(BASE, **over)
        popart = cfg.pop("popart", True)
        mb = synthetic_minibatch(B, A, seed=B + A + i, huber_delta=cfg["huber_delta"])
        mb["active_masks"][0] = 1.0                              # (B = 1: keep the mask sum non-zero)
        if not popart:
            mb = without_popart(mb)
        want = mappo_loss_oracle(**mb, **cfg)
        _check(cuda_device, mb, cfg, want)


def test_mappo_loss_gradient_rules_on_ties_and_the_huber_branches(cuda_device):
    """Hand-built rows: advantage 0, importance weight exactly 1 (tie inside the clip range), weights beyond either clip
    edge with either advantage sign, value exactly at the old prediction (tie of the two value losses), value outside the
    value clip range, errors above +delta (linear branch) and below -delta (the reference's zero branch)."""
    from oracle.mappo_loss_oracle import mappo_loss_oracle
    A = 8
    log_std = torch.full((A,), 0.7)
    std = torch.sigmoid(log_std) * 0.5
    shifts = torch.tensor([0.0, 0.0, 1.5, 1.5, -1.5, -1.5, 0.1, -0.1, 0.0, 0.0])
    B = shifts.numel()
    old_mean = torch.zeros(B, A)
    actions = old_mean + 0.5 * std
    mean = old_mean + shifts[:, None] * std / A
    old_logp = torch.distributions.Normal(old_mean, std).log_prob(actions)
    adv = torch.tensor([0.0, 1.0, 1.0, -1.0, 1.0, -1.0, 0.0, 2.0, 1.0, -1.0]).view(B, 1)
    value_preds = torch.zeros(B, 1)
    values = torch.tensor([0.0, 0.1, 0.5, 0.5, -0.5, -0.5, 0.2, -0.2, 0.0, 0.0]).view(B, 1)
    returns = torch.tensor([1.0, 1.0, 1.0, -1.0, 1.0, -1.0, 0.0, 0.0, 5.0, -5.0]).view(B, 1)
    active = torch.tensor([1.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0]).view(B, 1)
    mb = dict(mean=mean, log_std=log_std, values=values, actions=actions, old_logp=old_logp, adv_targ=adv,
              value_preds=value_preds, returns=returns, active_masks=active, ret_mean=None, ret_var=None)
    for over in (dict(huber_delta=2.0), dict(huber_delta=2.0, use_value_active_masks=True, use_policy_active_masks=True)):
        cfg = dict(BASE, **over)
        want = mappo_loss_oracle(**mb, **cfg)
        assert (want["grad_mean"][0] == 0).all() and (want["grad_mean"][1] != 0).any()
        assert want["grad_values"][9] == 0 and want["grad_values"][8] != 0            # e < -delta: zero; e > delta: linear
        _check(cuda_device, mb, cfg, want)


def test_mappo_loss_argument_errors(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.mappo_loss import mappo_loss, mappo_loss_raw
    from oracle.mappo_loss_oracle import synthetic_minibatch
    dev = cuda_device
    mb = synthetic_minibatch(64, 8, seed=3)
    d = {k: v.to(dev) for k, v in mb.items() if not k.startswith("popart_")}
    std = torch.sigmoid(d["log_std"]) * 0.5
    with pytest.raises(ValueError):                                # masks requested, none given
        mappo_loss(d["mean"], std, d["values"], d["actions"], d["old_logp"], d["adv_targ"], d["value_preds"], d["returns"],
                   use_policy_active_masks=True)
    with pytest.raises(ValueError):                                # per-dimension old log-probs are required
        mappo_loss(d["mean"], std, d["values"], d["actions"], d["old_logp"][:, :1], d["adv_targ"], d["value_preds"], d["returns"])
    with pytest.raises(L.MmbError):                                # CPU tensors: no CPU path
        mappo_loss_raw(mb["mean"], torch.sigmoid(mb["log_std"]) * 0.5, mb["values"], mb["actions"], mb["old_logp"],
                       mb["adv_targ"], mb["value_preds"], mb["returns"])
    # strided mean (column slice of a padded grouped head) gives the same result as the contiguous one
    wide = torch.zeros(64, 16, device=dev)
    wide[:, :8] = d["mean"]
    a = mappo_loss_raw(wide[:, :8], std, d["values"], d["actions"], d["old_logp"], d["adv_targ"], d["value_preds"], d["returns"])
    b = mappo_loss_raw(d["mean"], std, d["values"], d["actions"], d["old_logp"], d["adv_targ"], d["value_preds"], d["returns"])
    assert torch.equal(a[4], b[4]) and torch.equal(a[2], b[2]) and torch.equal(a[6], b[6])


@pytest.mark.parametrize("masks", [False, True])
def test_mappo_loss_in_kernel_finalisation(cuda_device, masks):
    """finalise=True (the default: the last block computes the two fp32 losses and the fp32 std gradient, and hands the
    fp64 scratch back zeroed) against the first version (fp64 sums reduced by torch ops): the same numbers, repeatedly (a
    scratch that did not come back zeroed would show up on the second call), for several batch sizes sharing the scratch."""
    from massive_marl_benchmark_b200.mappo_loss import mappo_loss_raw
    from oracle.mappo_loss_oracle import synthetic_minibatch
    dev = cuda_device
    for B in (64, 1000, 20000):
        mb = synthetic_minibatch(B, 8, seed=B)
        d = {k: v.to(dev) for k, v in mb.items() if not k.startswith("popart_")}
        std = torch.sigmoid(d["log_std"]) * 0.5
        args = (d["mean"], std, d["values"], d["actions"], d["old_logp"], d["adv_targ"], d["value_preds"], d["returns"])
        kw = dict(active_masks=d["active_masks"], use_value_active_masks=masks, use_policy_active_masks=masks)
        want = mappo_loss_raw(*args, finalise=False, **kw)
        for rep in range(3):
            got = mappo_loss_raw(*args, finalise=True, **kw)
            for i, (w, g) in enumerate(zip(want, got)):
                if i in (2, 3, 4, 6):                            # imp_weights, logp, grad_mean, grad_values: per-row, no reduction
                    assert torch.equal(w, g), (B, rep, i)
                else:                                            # sums of fp64 atomics in launch order: equal after the fp32 rounding
                    assert torch.allclose(w, g, rtol=2e-6, atol=1e-9), (B, rep, i, w, g)

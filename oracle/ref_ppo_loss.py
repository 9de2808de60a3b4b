"""TEST INFRASTRUCTURE, build container only - runs the REFERENCE's PPO minibatch loss on a given minibatch.

`ActorCritic.evaluate` is imported from the reference (agents/algorithms/rl/ppo/module.py:92-107) with its actor / critic
replaced by modules that return the supplied mean / value (so the gradients with respect to them can be read), and the
loss block of `PPO.update` (ppo.py: "# KL" ... "loss = ...") is extracted textually from the reference's source file and
executed.  Used by tests/golden/make_golden.py and tests/test_oracle_vs_reference.py; needs oracle.refshim.install().
"""
import importlib.util
import os
import textwrap
import types

import torch
import torch.nn as nn


class _Const(nn.Module):
    def __init__(self, t):
        super().__init__()
        self.t = nn.Parameter(t.clone())

    def forward(self, _):
        return self.t


def _load_module(reference_root):
    path = os.path.join(reference_root, "agents/algorithms/rl/ppo/module.py")
    spec = importlib.util.spec_from_file_location("_ref_ppo_module", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_ppo_loss(reference_root, mb, clip_param, value_loss_coef, entropy_coef, use_clipped_value_loss):
    mod = _load_module(reference_root)
    B, A = mb["mu"].shape
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):                   # the ctor prints both networks
        ac = mod.ActorCritic((4,), (4,), (A,), 1.0, {"pi_hid_sizes": [4], "vf_hid_sizes": [4], "activation": "elu"})
    ac.actor, ac.critic = _Const(mb["mu"]), _Const(mb["value"])
    with torch.no_grad():
        ac.log_std.copy_(mb["log_std"])

    src = open(os.path.join(reference_root, "agents/algorithms/rl/ppo/ppo.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.strip() == "# KL")
    end = next(i for i in range(start, len(src)) if src[i].strip().startswith("loss = surrogate_loss"))
    body = compile(textwrap.dedent("\n".join(src[start:end + 1])), "ppo.py:%d-%d" % (start + 1, end + 1), "exec")

    opt = types.SimpleNamespace(param_groups=[{"lr": 0.0}])
    self = types.SimpleNamespace(desired_kl=0.016, schedule="adaptive", step_size=3e-4, optimizer=opt, clip_param=clip_param,
                                 use_clipped_value_loss=use_clipped_value_loss, value_loss_coef=value_loss_coef,
                                 entropy_coef=entropy_coef)
    logp, entropy, value, mu, sigma = ac.evaluate(torch.zeros(B, 4), None, mb["actions"])        # ppo.py:266-268
    ns = {"self": self, "torch": torch, "actions_log_prob_batch": logp, "entropy_batch": entropy, "value_batch": value,
          "mu_batch": mu, "sigma_batch": sigma, "old_actions_log_prob_batch": mb["old_logp"],
          "advantages_batch": mb["advantages"], "target_values_batch": mb["target_values"], "returns_batch": mb["returns"],
          "old_mu_batch": mb["old_mu"], "old_sigma_batch": mb["old_sigma"], "max": max, "min": min}
    exec(body, ns)
    ns["loss"].backward()                                                                         # ppo.py:306
    return {"loss": ns["loss"].detach(), "surrogate_loss": ns["surrogate_loss"].detach(), "value_loss": ns["value_loss"].detach(),
            "kl_mean": ns["kl_mean"].detach(), "logp": logp.detach(), "entropy": entropy.detach()[0],
            "grad_mu": ac.actor.t.grad.clone(), "grad_log_std": ac.log_std.grad.clone(), "grad_value": ac.critic.t.grad.clone()}

"""TEST INFRASTRUCTURE, build container only - runs the REFERENCE's `MAPPO.ppo_update` on a given minibatch.

The reference's own trainer (agents/algorithms/marl/mappo_trainer.py), its own `ACTLayer` / `DiagGaussian` / `FixedNormal`
(agents/algorithms/utils/act.py, distributions.py) and its own `PopArt` (agents/algorithms/marl/utils/popart.py) are
imported under oracle.refshim; only the two MLP trunks are replaced by modules returning the supplied mean / values so
that the gradients with respect to them can be read, and the optimisers are no-ops.  `max_grad_norm` is set so large that
`clip_grad_norm_` multiplies by exactly 1.  Used by tests/golden/make_golden.py and tests/test_oracle_vs_reference.py.
"""
import types

import torch
import torch.nn as nn


class _Const(nn.Module):
    def __init__(self, t):
        super().__init__()
        self.t = nn.Parameter(t.clone())

    def forward(self, _):
        return self.t


class _NoOpt:
    def zero_grad(self):
        pass

    def step(self):
        pass


def reference_mappo_loss(mb, cfg):
    from agents.algorithms.marl.mappo_trainer import MAPPO
    from agents.algorithms.utils.act import ACTLayer
    from gym import spaces
    B, A = mb["mean"].shape
    config = {"clip_param": cfg["clip_param"], "ppo_epoch": 1, "num_mini_batch": 1, "data_chunk_length": None,
              "value_loss_coef": cfg["value_loss_coef"], "entropy_coef": cfg["entropy_coef"], "max_grad_norm": 1e30,
              "huber_delta": cfg["huber_delta"], "use_valuenorm": False, "use_recurrent_policy": False,
              "use_naive_recurrent_policy": False, "use_max_grad_norm": True,
              "use_clipped_value_loss": cfg["use_clipped_value_loss"], "use_huber_loss": cfg["use_huber_loss"],
              "use_popart": mb.get("popart_running_mean") is not None, "use_value_active_masks": cfg["use_value_active_masks"],
              "use_policy_active_masks": cfg["use_policy_active_masks"],
              "std_x_coef": cfg["std_x_coef"], "std_y_coef": cfg["std_y_coef"], "actor_gain": 0.01}
    act = ACTLayer(spaces.Box(low=-1.0, high=1.0, shape=(A,)), 4, True, 0.01, config)
    act.action_out.fc_mean = _Const(mb["mean"])
    with torch.no_grad():
        act.action_out.log_std.copy_(mb["log_std"])
    critic = _Const(mb["values"])

    def evaluate_actions(share_obs, obs, rnn_a, rnn_c, action, masks, available_actions=None, active_masks=None):
        # R_MAPPOPolicy.evaluate_actions -> R_Actor.evaluate_actions (actor_critic.py:109-114): the active masks reach the
        # distribution only with use_policy_active_masks
        logp, ent = act.evaluate_actions(torch.zeros(B, 4), action, available_actions,
                                         active_masks=active_masks if cfg["use_policy_active_masks"] else None)
        return critic(None), logp, ent

    policy = types.SimpleNamespace(evaluate_actions=evaluate_actions, actor_optimizer=_NoOpt(), critic_optimizer=_NoOpt(),
                                   actor=act, critic=critic)
    trainer = MAPPO(config, policy)
    if config["use_popart"]:
        # start the reference's normaliser from the given running statistics; its two training-mode calls inside
        # cal_value_loss then update it for real, and the oracle's popart_update must reproduce both sets of moments
        from oracle.mappo_loss_oracle import popart_update
        pa = trainer.value_normalizer
        state = {k: mb["popart_" + k].clone() for k in ("running_mean", "running_mean_sq", "debiasing_term")}
        pa.running_mean.copy_(state["running_mean"]); pa.running_mean_sq.copy_(state["running_mean_sq"])
        pa.debiasing_term.copy_(state["debiasing_term"])
        m1, v1 = popart_update(state, mb["returns"])
        m2, v2 = popart_update(state, mb["returns"])
        mb = dict(mb, ret_mean=m1.clone(), ret_var=v1.clone(), ret_mean_orig=m2.clone(), ret_var_orig=v2.clone())
    sample = (None, None, None, None, mb["actions"], mb["value_preds"], mb["returns"], None, mb["active_masks"],
              mb["old_logp"], mb["adv_targ"], None, None)
    value_loss, _, policy_loss, dist_entropy, _, imp_weights = trainer.ppo_update(sample)
    out = {"policy_loss": policy_loss.detach(), "dist_entropy": dist_entropy.detach(), "value_loss": value_loss.detach(),
           "imp_weights": imp_weights.detach(),
           "grad_mean": act.action_out.fc_mean.t.grad.clone(), "grad_log_std": act.action_out.log_std.grad.clone(),
           "grad_values": critic.t.grad.clone()}
    if config["use_popart"]:
        pa = trainer.value_normalizer
        for k, t in (("running_mean", pa.running_mean), ("running_mean_sq", pa.running_mean_sq), ("debiasing_term", pa.debiasing_term)):
            if not torch.equal(t, state[k]):
                raise AssertionError("popart_update != the reference's PopArt after two calls: " + k)
    return {k: v for k, v in out.items()}, {k: v for k, v in mb.items() if not k.startswith("popart_")}

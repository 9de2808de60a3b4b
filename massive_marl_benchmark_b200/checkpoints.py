"""Checkpoint formats of the reference, read and written unchanged, so that files saved by the
reference load here and files saved here load in the reference:

  PPO    `model_{it}.pt`  = `actor_critic.state_dict()` (agents/algorithms/rl/ppo/ppo.py:90-97): keys `log_std`,
         `actor.{0,2,4,...}.{weight,bias}`, `critic.{0,2,4,...}.{weight,bias}` - Linear layers at the even indices of the two
         `nn.Sequential`s, activations in between (module.py:25-49)
  MARL   `actor_agent{i}.pt`, `critic_agent{i}.pt` = the per-agent `policy.actor.state_dict()` / `policy.critic.state_dict()`
         (agents/algorithms/marl/runner.py:319-339); consumed by `mlp.FusedMLP.from_marl_state_dict`

The module built by `ppo_modules_from_state_dict` has the attribute surface the rest of this package and the reference's
`PPO.update` use (`.actor`, `.critic`, `.log_std`, `.asymmetric`), infers the layer sizes from the tensors, and its own
`state_dict()` has exactly the reference's keys.
"""
import os
import re

import torch
import torch.nn as nn


class PPOModules(nn.Module):
    def __init__(self, actor_sizes, critic_sizes, act_dim, asymmetric=False):
        super().__init__()
        self.asymmetric = bool(asymmetric)

        def seq(sizes):
            mods = []
            for i in range(len(sizes) - 1):
                mods.append(nn.Linear(sizes[i], sizes[i + 1]))
                if i + 2 < len(sizes):
                    mods.append(nn.ELU())                          # cfg/ppo/config.yaml:9 (the only activation on the path)
            return nn.Sequential(*mods)

        self.actor, self.critic = seq(actor_sizes), seq(critic_sizes)
        self.log_std = nn.Parameter(torch.zeros(act_dim))


def _layer_sizes(sd, prefix):
    idx = sorted(int(m.group(1)) for k in sd for m in [re.fullmatch(re.escape(prefix) + r"\.(\d+)\.weight", k)] if m)
    if not idx or idx != list(range(0, 2 * len(idx), 2)):
        raise ValueError("%s: expected Linear layers at the even indices of a Sequential, found %r" % (prefix, idx))
    sizes = [sd["%s.%d.weight" % (prefix, idx[0])].shape[1]]
    for i in idx:
        w = sd["%s.%d.weight" % (prefix, i)]
        if w.shape[1] != sizes[-1]:
            raise ValueError("%s.%d.weight: input width %d does not follow %d" % (prefix, i, w.shape[1], sizes[-1]))
        sizes.append(w.shape[0])
    return sizes


def ppo_modules_from_state_dict(sd, asymmetric=None):
    """The networks of a reference PPO checkpoint (`torch.load('model_1000.pt')`).  `asymmetric`: whether the critic reads
    the states instead of the observations (module.py:38-41); inferred as `critic input width != actor input width` when
    None."""
    a, c = _layer_sizes(sd, "actor"), _layer_sizes(sd, "critic")
    if c[-1] != 1:
        raise ValueError("critic head width %d, expected 1" % c[-1])
    if tuple(sd["log_std"].shape) != (a[-1],):
        raise ValueError("log_std %s does not match the actor head width %d" % (tuple(sd["log_std"].shape), a[-1]))
    m = PPOModules(a, c, a[-1], asymmetric=(a[0] != c[0]) if asymmetric is None else asymmetric)
    m.load_state_dict(sd, strict=True)
    return m


def load_ppo(path, map_location="cpu"):
    """`PPO.load` (ppo.py:90-94): (modules, learning iteration parsed from `model_{it}.pt`)."""
    sd = torch.load(path, map_location=map_location)
    return ppo_modules_from_state_dict(sd), int(path.split("_")[-1].split(".")[0])


def save_ppo(actor_critic, log_dir, it):
    """`PPO.save` (ppo.py:96-97, called as `model_{it}.pt` at ppo.py:172-175)."""
    path = os.path.join(log_dir, "model_{}.pt".format(it))
    torch.save(actor_critic.state_dict(), path)
    return path


def load_marl_agents(model_dir, num_agents, map_location="cpu"):
    """`Runner.restore` (runner.py:331-339): ([actor state dicts], [critic state dicts]) of `actor_agent{i}.pt` /
    `critic_agent{i}.pt`, ready for `mlp.MarlTeamForward(actor_sds, critic_sds)`."""
    actors = [torch.load(os.path.join(model_dir, "actor_agent%d.pt" % i), map_location=map_location) for i in range(num_agents)]
    critics = [torch.load(os.path.join(model_dir, "critic_agent%d.pt" % i), map_location=map_location) for i in range(num_agents)]
    return actors, critics


def save_marl_agents(policies, save_dir):
    """`Runner.save` (runner.py:319-329): one actor and one critic file per agent."""
    for i, p in enumerate(policies):
        torch.save(p.actor.state_dict(), os.path.join(save_dir, "actor_agent%d.pt" % i))
        torch.save(p.critic.state_dict(), os.path.join(save_dir, "critic_agent%d.pt" % i))

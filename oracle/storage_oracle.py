"""Torch restatement of the rollout-storage arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

PPO:   reference agents/algorithms/rl/ppo/storage.py (RolloutStorage)
MARL:  reference agents/algorithms/marl/utils/separated_buffer.py (SeparatedReplayBuffer),
       agents/algorithms/marl/utils/popart.py / valuenorm.py (running_mean_var, denormalize),
       agents/algorithms/marl/mappo_trainer.py:189-199 (advantage prologue),
       agents/algorithms/marl/runner.py:229-255 (mask logic of Runner.insert).
"""
from typing import Optional, Tuple

import torch


# ---------------------------------------------------------------------------------------------
# PPO RolloutStorage
# ---------------------------------------------------------------------------------------------

def ppo_compute_returns(rewards, values, dones, last_values, gamma: float, lam: float):
    """storage.py:51-65.  rewards/values [T,N,1] fp32, dones [T,N,1] uint8, last_values [N,1].
    Returns (returns, normalised advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def ppo_get_statistics(dones, rewards):
    """storage.py:67-73: mean trajectory length (last row forced done, env-major flatten), mean reward."""
    done = dones.cpu().clone()
    done[-1] = 1
    flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
    done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64),
                              flat_dones.nonzero(as_tuple=False)[:, 0]))
    trajectory_lengths = (done_indices[1:] - done_indices[:-1])
    return trajectory_lengths.float().mean(), rewards.mean()


def ppo_minibatch_partition(num_envs: int, T: int, num_mini_batches: int, perm=None):
    """storage.py:75-87: BatchSampler(drop_last=True) over range(T*N) ('sequential') or over a
    permutation ('random'; the permutation is an input because torch's CPU randperm stream is the
    reference's only source of it).  Returns a list of index lists."""
    batch_size = num_envs * T
    mb = batch_size // num_mini_batches
    order = list(range(batch_size)) if perm is None else [int(i) for i in perm]
    out = []
    for i in range(0, batch_size - mb + 1, mb):
        out.append(order[i:i + mb])
    return out


# ---------------------------------------------------------------------------------------------
# MARL SeparatedReplayBuffer
# ---------------------------------------------------------------------------------------------

def popart_running_mean_var(running_mean, running_mean_sq, debiasing_term, epsilon: float = 1e-5):
    """popart.py:30-34 / valuenorm.py:32-37"""
    debiased_mean = running_mean / debiasing_term.clamp(min=epsilon)
    debiased_mean_sq = running_mean_sq / debiasing_term.clamp(min=epsilon)
    debiased_var = (debiased_mean_sq - debiased_mean ** 2).clamp(min=1e-2)
    return debiased_mean, debiased_var


def popart_denormalize(x, mean, var):
    """popart.py:64-75 with norm_axes=1: x * sqrt(var) + mean"""
    return x * torch.sqrt(var)[(None,) * 1] + mean[(None,) * 1]


def marl_compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma: float, gae_lambda: float,
                         denorm: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, use_gae: bool = True,
                         use_proper_time_limits: bool = False, use_popart: bool = True):
    """separated_buffer.py:124-168, all four branches.  rewards [T,N,1]; value_preds/masks/bad_masks
    [T+1,N,1]; ``denorm`` = (mean, var) of PopArt/ValueNorm or None.  Returns (returns [T+1,N,1],
    value_preds with slot T overwritten)."""
    T = rewards.shape[0]
    value_preds = value_preds.clone()
    returns = torch.zeros_like(value_preds)

    def D(x):
        return popart_denormalize(x, *denorm) if denorm is not None else x

    if use_proper_time_limits:
        if use_gae:
            value_preds[-1] = next_value
            gae = 0
            for step in reversed(range(T)):
                delta = rewards[step] + gamma * D(value_preds[step + 1]) * masks[step + 1] - D(value_preds[step])
                gae = delta + gamma * gae_lambda * masks[step + 1] * gae
                gae = gae * bad_masks[step + 1]
                returns[step] = gae + D(value_preds[step])
        else:
            returns[-1] = next_value
            for step in reversed(range(T)):
                vp = D(value_preds[step]) if (use_popart and denorm is not None) else value_preds[step]
                returns[step] = (returns[step + 1] * gamma * masks[step + 1] + rewards[step]) * bad_masks[step + 1] \
                    + (1 - bad_masks[step + 1]) * vp
    else:
        if use_gae:
            value_preds[-1] = next_value
            gae = 0
            for step in reversed(range(T)):
                delta = rewards[step] + gamma * D(value_preds[step + 1]) * masks[step + 1] - D(value_preds[step])
                gae = delta + gamma * gae_lambda * masks[step + 1] * gae
                returns[step] = gae + D(value_preds[step])
        else:
            returns[-1] = next_value
            for step in reversed(range(T)):
                returns[step] = returns[step + 1] * gamma * masks[step + 1] + rewards[step]
    return returns, value_preds


def marl_advantages(returns, value_preds, denorm=None, eps: float = 1e-5):
    """mappo_trainer.py:189-199 / happo_trainer.py:180-189"""
    if denorm is not None:
        advantages = returns[:-1] - popart_denormalize(value_preds[:-1], *denorm)
    else:
        advantages = returns[:-1] - value_preds[:-1]
    advantages_copy = advantages.clone()
    mean_advantages = torch.mean(advantages_copy)
    std_advantages = torch.std(advantages_copy)
    return (advantages - mean_advantages) / (std_advantages + eps)


def runner_insert_masks(dones):
    """runner.py:229-255: dones (N,A) int64 -> masks (N,A,1), active_masks (N,A,1) fp32."""
    N, A = dones.shape
    dones_env = torch.all(dones.bool(), dim=1)
    masks = torch.ones(N, A, 1)
    masks[dones_env == True] = torch.zeros(int((dones_env == True).sum()), A, 1)
    active_masks = torch.ones(N, A, 1)
    active_masks[dones.bool() == True] = torch.zeros(int((dones.bool() == True).sum()), 1)
    active_masks[dones_env == True] = torch.ones(int((dones_env == True).sum()), A, 1)
    return masks, active_masks

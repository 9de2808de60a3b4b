"""TEST INFRASTRUCTURE - CPU restatement of the episode bookkeeping of the reference's PPO runner.

Follows agents/algorithms/rl/ppo/ppo.py:113-157 statement by statement (torch tensors for the running sums, Python
lists + deque(maxlen=100) for the finished episodes, `statistics.mean` for the logged value, ppo.py:198-200).
Pinned: the block is inline in PPO.run() and cannot be imported on its own, so tests/test_oracle_vs_reference.py
extracts the reference's own source lines (ppo.py:144-151) and executes them against this class on the same inputs
(deques, running sums and means identical).
"""
import statistics
from collections import deque

import torch


class EpisodeOracle:
    def __init__(self, num_envs, window=100):
        self.rewbuffer = deque(maxlen=window)                    # ppo.py:114
        self.lenbuffer = deque(maxlen=window)                    # ppo.py:115
        self.cur_reward_sum = torch.zeros(num_envs, dtype=torch.float)      # ppo.py:116
        self.cur_episode_length = torch.zeros(num_envs, dtype=torch.float)  # ppo.py:117
        self.finished = 0

    def step(self, rews, dones):
        """One env step (ppo.py:143-151)."""
        reward_sum, episode_length = [], []
        self.cur_reward_sum[:] += rews
        self.cur_episode_length[:] += 1
        new_ids = (dones > 0).nonzero(as_tuple=False)
        reward_sum.extend(self.cur_reward_sum[new_ids][:, 0].cpu().numpy().tolist())
        episode_length.extend(self.cur_episode_length[new_ids][:, 0].cpu().numpy().tolist())
        self.cur_reward_sum[new_ids] = 0
        self.cur_episode_length[new_ids] = 0
        self.rewbuffer.extend(reward_sum)                        # ppo.py:156 (per iteration there; order is identical)
        self.lenbuffer.extend(episode_length)
        self.finished += len(reward_sum)

    def update(self, rewards, dones):
        for t in range(rewards.shape[0]):
            self.step(rewards[t].reshape(-1), dones[t].reshape(-1))

    def means(self):
        if not self.rewbuffer:
            return float("nan"), float("nan")
        return statistics.mean(self.rewbuffer), statistics.mean(self.lenbuffer)    # ppo.py:199-200

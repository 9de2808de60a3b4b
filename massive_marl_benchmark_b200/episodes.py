"""Device-side episode bookkeeping of the PPO runner (agents/algorithms/rl/ppo/ppo.py:113-157,198-220).

The reference keeps `cur_reward_sum` / `cur_episode_length` on the device but moves every finished episode to the
host each step (`nonzero` + `.cpu().numpy().tolist()` = one sync per env step) into two `deque(maxlen=100)` whose
`statistics.mean` is logged.  `EpisodeTracker` does the same accounting for T steps per call in two launches
(`mmb_episode_update`) and keeps the two deques as device rings; nothing is read back unless asked.

    tracker = EpisodeTracker(num_envs, device)
    tracker.update(storage.rewards, storage.dones)     # [T, N(,1)] planes of a rollout, or [N] of one step
    mean_reward, mean_length = tracker.means()          # 0-dim device tensors (= statistics.mean of the deques)
"""
import torch

from . import _lib as L


class EpisodeTracker:
    def __init__(self, num_envs, device="cuda", window=100):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.MmbError("EpisodeTracker needs a CUDA device (there is no CPU path); got %r" % (device,))
        L.lib()
        self.num_envs, self.window, self.device = num_envs, window, dev
        self.cur_reward_sum = torch.zeros(num_envs, device=dev)          # ppo.py:116
        self.cur_episode_length = torch.zeros(num_envs, device=dev)      # ppo.py:117
        self.reward_ring = torch.zeros(window, device=dev)               # rewbuffer = deque(maxlen=100)
        self.length_ring = torch.zeros(window, device=dev)               # lenbuffer
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)       # [0] = episodes finished so far
        self._scratch = None
        self._slots = torch.arange(window, device=dev)

    def update(self, rewards, dones):
        """rewards [T, N] (or [T, N, 1], or [N] for one step) fp32; dones same shape, uint8 or int64 (> 0 = done)."""
        if rewards.dim() == 1:
            rewards, dones = rewards.unsqueeze(0), dones.unsqueeze(0)
        T = rewards.shape[0]
        rewards = rewards.reshape(T, -1)
        dones = dones.reshape(T, -1)
        if rewards.shape[1] != self.num_envs or dones.shape != rewards.shape:
            raise ValueError("expected [T, %d] planes, got %r / %r" % (self.num_envs, tuple(rewards.shape), tuple(dones.shape)))
        if rewards.dtype != torch.float32 or rewards.stride(1) != 1:
            rewards = rewards.float().contiguous()
        if dones.dtype not in (torch.uint8, torch.int64) or dones.stride(1) != 1:
            dones = dones.to(torch.int64).contiguous()
        if self._scratch is None or self._scratch.shape[1] != T:
            self._scratch = torch.empty(2, T, self.num_envs, device=self.device)
        p = L.EpisodeParams()
        p.num_envs, p.num_steps, p.window = self.num_envs, T, self.window
        p.rewards, p.rewards_row_stride = L.ptr(rewards), rewards.stride(0)
        if dones.dtype == torch.uint8:
            p.dones_u8, p.dones_u8_row_stride = L.ptr(dones), dones.stride(0)
        else:
            p.dones_i64, p.dones_i64_row_stride = L.ptr(dones), dones.stride(0)
        p.cur_reward_sum, p.cur_episode_length = L.ptr(self.cur_reward_sum), L.ptr(self.cur_episode_length)
        p.ep_reward, p.ep_length = L.ptr(self._scratch[0]), L.ptr(self._scratch[1])
        p.reward_ring, p.length_ring, p.state = L.ptr(self.reward_ring), L.ptr(self.length_ring), L.ptr(self.state)
        self._keep = (rewards, dones)
        L.check(L.lib().mmb_episode_update(p, L.stream_ptr()), "mmb_episode_update")

    def update_marl(self, rewards, dones):
        """The MARL runner's variant (agents/algorithms/marl/runner.py:135-144): per env step `dones_env = all(dones, 1)`,
        `reward_env = mean(rewards, 1).flatten()`, `train_episode_rewards += reward_env`, finished episodes appended at
        `dones_env`.  rewards [T, N, A, 1] (or [N, A, 1]), dones [T, N, A] (or [N, A]); the two reductions over the agent
        axis are the reference's own torch ops, the bookkeeping is `update`."""
        if rewards.dim() == 3:
            rewards, dones = rewards.unsqueeze(0), dones.unsqueeze(0)
        T, N = rewards.shape[0], rewards.shape[1]
        reward_env = torch.mean(rewards.reshape(T, N, -1), dim=2)
        dones_env = torch.all(dones.reshape(T, N, -1) != 0, dim=2).to(torch.uint8)
        self.update(reward_env, dones_env)

    @property
    def finished(self):
        """Episodes finished so far: 0-dim int64 device tensor (no sync)."""
        return self.state[0]

    def means(self):
        """(mean reward, mean length) over the last min(finished, window) episodes, as `statistics.mean(rewbuffer)` /
        `statistics.mean(lenbuffer)` (ppo.py:199-200); NaN while no episode has finished.  Device tensors, no sync."""
        n = torch.clamp(self.state[0], max=self.window)
        mask = (self._slots < n).to(torch.float64)
        nf = n.to(torch.float64)
        return ((self.reward_ring.double() * mask).sum() / nf).float(), ((self.length_ring.double() * mask).sum() / nf).float()

    def deques(self):
        """The two deques in the reference's order (oldest first) as Python lists - host sync; for tests and logging."""
        n = int(self.state[0].item())
        k = min(n, self.window)
        order = [(n - k + i) % self.window for i in range(k)]
        rr, ll = self.reward_ring.cpu(), self.length_ring.cpu()
        return [float(rr[i]) for i in order], [float(ll[i]) for i in order]

"""Synthetic Isaac-Gym-layout state frames (the stand-in for ``gym.simulate`` + ``gym.refresh_*``).

PhysX is proprietary and out of scope (BASELINE.json north_star), so the hot path is fed with
synthetic or replayed root / DOF / sensor state tensors in Isaac Gym's exact memory layout
(SURVEY.md Appendix C).  Frames are generated on the CPU from a seeded ``torch.Generator`` and
uploaded, so the CPU oracle and the GPU kernels see identical bits.

Layouts (all fp32, contiguous, env-major):
  TenAnt          root (11N,13): rows 11e+k ant k, 11e+10 box;  dof (80N,2): row 80e+8k+j
  OneAnt          root (2N,13):  rows 2e ant, 2e+1 box;         dof (8N,2);  sensor (4N,6)
  MultiIngenuity  root (4N,13):  rows 4e+h helicopter h
Root row = [px py pz | qx qy qz qw | vx vy vz | wx wy wz].
"""
import math
from typing import Dict

import torch

# nv_ant.xml:48-75 joint ranges in degrees, MJCF body-tree order
ANT_DOF_RANGE_DEG = ((-40, 40), (30, 100), (-40, 40), (-100, -30), (-40, 40), (-100, -30), (-40, 40), (30, 100))


def ant_dof_limits():
    lo = torch.tensor([math.radians(a) for a, _ in ANT_DOF_RANGE_DEG], dtype=torch.float32)
    hi = torch.tensor([math.radians(b) for _, b in ANT_DOF_RANGE_DEG], dtype=torch.float32)
    return lo, hi


def ant_initial_dof_pos():
    """lower where lower>0, upper where upper<0, else 0 (reference ten_ant.py:133-137)."""
    lo, hi = ant_dof_limits()
    z = torch.zeros(8)
    return torch.where(lo > 0, lo, torch.where(hi < 0, hi, z))


def _small_quats(n, gen, sigma=0.3):
    ax = torch.randn(n, 3, generator=gen)
    ax = ax / ax.norm(dim=-1, keepdim=True).clamp(min=1e-6)
    ang = sigma * torch.randn(n, generator=gen)
    q = torch.empty(n, 4)
    q[:, :3] = ax * torch.sin(ang / 2)[:, None]
    q[:, 3] = torch.cos(ang / 2)
    return q / q.norm(dim=-1, keepdim=True)


def _ant_dofs(n_rows, gen, edge_frac=0.02):
    lo, hi = ant_dof_limits()
    reps = n_rows // 8
    lo_r, hi_r = lo.repeat(reps), hi.repeat(reps)
    pos = lo_r + (hi_r - lo_r) * torch.rand(n_rows, generator=gen)
    edge = torch.rand(n_rows, generator=gen) < edge_frac          # exercise the `> 0.99` limit test
    jitter = 0.02 * (2 * torch.rand(n_rows, generator=gen) - 1)
    pos = torch.where(edge, hi_r + jitter, pos)
    dof = torch.empty(n_rows, 2)
    dof[:, 0] = pos
    dof[:, 1] = 2.0 * torch.randn(n_rows, generator=gen)
    return dof


def ten_ant_initial_root(num_envs):
    """Start poses (reference ten_ant.py:339-358, box :494-495), identity rotation, zero velocity."""
    rows = []
    for k in range(10):
        y = (1.5 + 3.0 * (k // 2)) * (-1.0 if k % 2 == 0 else 1.0)
        rows.append([6.0, y, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    rows.append([4.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    return torch.tensor(rows, dtype=torch.float32).repeat(num_envs, 1)


def one_ant_initial_root(num_envs):
    rows = [[-6.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], [-4.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]]
    return torch.tensor(rows, dtype=torch.float32).repeat(num_envs, 1)


def ingenuity_initial_root(num_envs):
    rows = [[0.0, y, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0] for y in (2.0, -2.0, 6.0, -6.0)]
    return torch.tensor(rows, dtype=torch.float32).repeat(num_envs, 1)


def ten_ant_frames(num_envs: int, num_frames: int, seed: int = 1234, fall_prob: float = 0.001,
                   pos_sigma: float = 0.5) -> Dict[str, torch.Tensor]:
    """Frames for TenAnt: root [F,11N,13], dof [F,80N,2], actions [F,N,80] (SURVEY.md section 8d cfg 2)."""
    gen = torch.Generator().manual_seed(seed)
    N, F = num_envs, num_frames
    R = 11 * N
    init = ten_ant_initial_root(N)
    root = init.unsqueeze(0).repeat(F, 1, 1)
    root[:, :, :2] += pos_sigma * torch.randn(F, R, 2, generator=gen)
    up = 0.32 + 0.68 * torch.rand(F, R, generator=gen)
    down = 0.10 + 0.21 * torch.rand(F, R, generator=gen)
    root[:, :, 2] = torch.where(torch.rand(F, R, generator=gen) < fall_prob, down, up)
    root[:, :, 3:7] = _small_quats(F * R, gen).view(F, R, 4)
    root[:, :, 7:13] = torch.randn(F, R, 6, generator=gen)
    # box rows: xy = (4,0)+N(0,0.3^2), z = 1, yaw-only quaternion
    th = 0.2 * torch.randn(F, N, generator=gen)
    box = root[:, 10::11, :]
    box[:, :, 0] = 4.0 + 0.3 * torch.randn(F, N, generator=gen)
    box[:, :, 1] = 0.0 + 0.3 * torch.randn(F, N, generator=gen)
    box[:, :, 2] = 1.0
    box[:, :, 3] = 0.0
    box[:, :, 4] = 0.0
    box[:, :, 5] = torch.sin(th / 2)
    box[:, :, 6] = torch.cos(th / 2)
    dof = _ant_dofs(F * 80 * N, gen).view(F, 80 * N, 2)
    actions = 2 * torch.rand(F, N, 80, generator=gen) - 1
    return dict(root=root.contiguous(), dof=dof.contiguous(), actions=actions.contiguous())


def one_ant_frames(num_envs: int, num_frames: int, seed: int = 1234, fall_prob: float = 0.01) -> Dict[str, torch.Tensor]:
    """Frames for OneAnt: root [F,2N,13], dof [F,8N,2], sensor [F,4N,6], actions [F,N,8]."""
    gen = torch.Generator().manual_seed(seed)
    N, F = num_envs, num_frames
    R = 2 * N
    root = one_ant_initial_root(N).unsqueeze(0).repeat(F, 1, 1)
    root[:, :, :2] += 1.0 * torch.randn(F, R, 2, generator=gen)
    up = 0.32 + 0.68 * torch.rand(F, R, generator=gen)
    down = 0.10 + 0.21 * torch.rand(F, R, generator=gen)
    root[:, :, 2] = torch.where(torch.rand(F, R, generator=gen) < fall_prob, down, up)
    root[:, :, 3:7] = _small_quats(F * R, gen).view(F, R, 4)
    root[:, :, 7:13] = torch.randn(F, R, 6, generator=gen)
    # a few boxes close to their target (0,0) so that goal_arrive / success fire
    near = torch.rand(F, N, generator=gen) < 0.1
    box = root[:, 1::2, :]
    box[:, :, 0] = torch.where(near, 0.3 * torch.randn(F, N, generator=gen), box[:, :, 0])
    box[:, :, 1] = torch.where(near, 0.3 * torch.randn(F, N, generator=gen), box[:, :, 1])
    dof = _ant_dofs(F * 8 * N, gen).view(F, 8 * N, 2)
    sensor = 5.0 * torch.randn(F, 4 * N, 6, generator=gen)
    actions = 2 * torch.rand(F, N, 8, generator=gen) - 1
    return dict(root=root.contiguous(), dof=dof.contiguous(), sensor=sensor.contiguous(), actions=actions.contiguous())


def ingenuity_frames(num_envs: int, num_frames: int, seed: int = 1234, pos_sigma: float = 1.5) -> Dict[str, torch.Tensor]:
    """Frames for MultiIngenuity: root [F,4N,13], actions [F,N,24].  Positions are spread so that some
    helicopters are farther than 8 from their goal and some are below z = 0.5."""
    gen = torch.Generator().manual_seed(seed)
    N, F = num_envs, num_frames
    R = 4 * N
    root = ingenuity_initial_root(N).unsqueeze(0).repeat(F, 1, 1)
    root[:, :, :3] += pos_sigma * torch.randn(F, R, 3, generator=gen)
    root[:, :, 2] += 1.5
    root[:, :, 3:7] = _small_quats(F * R, gen).view(F, R, 4)
    root[:, :, 7:13] = torch.randn(F, R, 6, generator=gen)
    actions = 2 * torch.rand(F, N, 24, generator=gen) - 1
    return dict(root=root.contiguous(), actions=actions.contiguous())


def reset_noise(num_envs: int, num_frames: int, seed: int = 99):
    """Per-frame reset noise in the reference's distributions (ten_ant.py:822-823):
    positions U(-0.2,0.2), velocities U(-0.1,0.1), each [F,N,8]; row i of a frame feeds the i-th reset env."""
    gen = torch.Generator().manual_seed(seed)
    pos = (0.2 - -0.2) * torch.rand(num_frames, num_envs, 8, generator=gen) + -0.2
    vel = (0.1 - -0.1) * torch.rand(num_frames, num_envs, 8, generator=gen) + -0.1
    return pos, vel

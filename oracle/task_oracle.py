"""Step-level oracles: the reference task classes' per-step ordering, restated (TEST INFRASTRUCTURE).

``TenAntOracle`` / ``OneAntOracle`` / ``IngenuityOracle`` keep the same buffers as the reference
task classes and run one ``BaseTask.step`` (reference agents/tasks/agent_base/base_task.py:129-149)
against a *frame provider*: the PhysX step is out of scope, so "simulate + refresh" is replaced by
copying a supplied Isaac-Gym-layout state frame into ``root_states`` / ``dof_state`` / ``sensor``
at the point where the reference calls ``gym.refresh_*``.

Side effects the reference hands to the simulator are recorded in ``self.last``:
forces (set_dof_actuation_force_tensor / apply_rigid_body_force_tensors), the int32 index lists
of set_actor_root_state_tensor_indexed / set_dof_state_tensor_indexed, ``env_ids`` and the
dof_state rows pushed at reset.
"""
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import isaac_torch_utils as itu
from . import task_math as tm

# nv_ant.xml:48-75 joint ranges in degrees (hip_1, ankle_1, hip_2, ankle_2, hip_3, ankle_3, hip_4, ankle_4)
ANT_DOF_RANGE_DEG = ((-40, 40), (30, 100), (-40, 40), (-100, -30), (-40, 40), (-100, -30), (-40, 40), (30, 100))
ANT_GEAR = 15.0  # nv_ant.xml:83-90


def ant_dof_limits(device="cpu"):
    lo = torch.tensor([math.radians(a) for a, _ in ANT_DOF_RANGE_DEG], dtype=torch.float32, device=device)
    hi = torch.tensor([math.radians(b) for _, b in ANT_DOF_RANGE_DEG], dtype=torch.float32, device=device)
    return lo, hi


DEFAULT_ENV_CFG = dict(  # cfg/TenAnt.yaml:6,39-52,64 (OneAnt/MultiIngenuity yaml identical on these keys)
    episodeLength=1000, powerScale=1.0, headingWeight=0.5, upWeight=0.1, actionsCost=0.005,
    energyCost=0.05, dofVelocityScale=0.2, contactForceScale=0.1, jointsAtLimitCost=0.1,
    deathCost=-2.0, terminationHeight=0.31, dt=0.0166)


def ten_ant_initial_root(num_envs, device="cpu"):
    """FakeGym prepare_sim: start poses of ten_ant.py:339-358 (ants) and :494-495 (box), identity quat."""
    rows = []
    for k in range(10):
        y = (1.5 + 3.0 * (k // 2)) * (-1.0 if k % 2 == 0 else 1.0)
        rows.append([6.0, y, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    rows.append([4.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0])
    return torch.tensor(rows, dtype=torch.float32, device=device).repeat(num_envs, 1)


def one_ant_initial_root(num_envs, device="cpu"):
    """one_ant.py:233-234 (ant at (-6,0,1)) and the box pose (one_ant.py:262-266)."""
    rows = [[-6.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], [-4.0, 0.0, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]]
    return torch.tensor(rows, dtype=torch.float32, device=device).repeat(num_envs, 1)


def ingenuity_initial_root(num_envs, device="cpu"):
    """multi_ingenuity.py:158-165"""
    rows = [[0.0, y, 1.0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0] for y in (2.0, -2.0, 6.0, -6.0)]
    return torch.tensor(rows, dtype=torch.float32, device=device).repeat(num_envs, 1)


class _TaskOracleBase:
    def _alloc_base(self, num_envs, num_obs, device):
        # reference base_task.py:56-68
        self.num_envs = num_envs
        self.device = device
        self.obs_buf = torch.zeros((num_envs, num_obs), device=device, dtype=torch.float)
        self.rew_buf = torch.zeros(num_envs, device=device, dtype=torch.float)
        self.reset_buf = torch.ones(num_envs, device=device, dtype=torch.long)
        self.progress_buf = torch.zeros(num_envs, device=device, dtype=torch.long)
        self.last: Dict[str, Optional[torch.Tensor]] = {}

    def _draw_noise(self, n, noise):
        """reference ten_ant.py:822-823 / one_ant.py:371-372: positions first, then velocities."""
        if noise is not None:
            positions, velocities = noise
            positions, velocities = positions[:n], velocities[:n]
        else:
            positions = itu.torch_rand_float(-0.2, 0.2, (n, 8), device=self.device)
            velocities = itu.torch_rand_float(-0.1, 0.1, (n, 8), device=self.device)
        return positions, velocities


class TenAntOracle(_TaskOracleBase):
    """reference agents/tasks/ten_ant.py (TenAnt), 10-agent internal layout, flat (N,80) actions."""

    A = 10

    def __init__(self, num_envs, cfg=None, device="cpu", initial_root=None):
        cfg = dict(DEFAULT_ENV_CFG, **(cfg or {}))
        self.cfg = cfg
        self._alloc_base(num_envs, 388, device)
        N = num_envs
        self.root_states = (ten_ant_initial_root(N, device) if initial_root is None else initial_root.clone())
        self.initial_root_states = self.root_states.clone()
        self.initial_root_states[:, 7:13] = 0                       # ten_ant.py:99-100
        self.dof_state = torch.zeros(N * 80, 2, device=device)
        v = self.dof_state.view(N, -1, 2)
        self.dof_pos = [v[:, 8 * k:8 * k + 8, 0] for k in range(10)]  # ten_ant.py:107-127
        self.dof_vel = [v[:, 8 * k:8 * k + 8, 1] for k in range(10)]
        self.dof_limits_lower, self.dof_limits_upper = ant_dof_limits(device)
        zero = torch.tensor([0.0], device=device)
        init = torch.zeros_like(self.dof_pos[0])
        self.initial_dof_pos = torch.where(self.dof_limits_lower > zero, self.dof_limits_lower,
                                           torch.where(self.dof_limits_upper < zero, self.dof_limits_upper, init))
        self.joint_gears = torch.full((80,), ANT_GEAR, device=device)
        self.up_vec = torch.tensor([0.0, 0.0, 1.0], device=device).repeat(N, 1)
        self.heading_vec = torch.tensor([1.0, 0.0, 0.0], device=device).repeat(N, 1)
        self.inv_start_rot = itu.quat_conjugate(torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)).repeat(N, 1)
        self.targets = torch.zeros(N, 3, device=device)
        self.box_targets = torch.zeros(N, 2, device=device)
        self.box_targets_k = []
        for c in tm.TEN_ANT_GOAL_OFFSETS:                             # ten_ant.py:172-181
            self.box_targets_k.append(torch.tensor([0.0, -c], device=device).repeat(N, 1))
            self.box_targets_k.append(torch.tensor([0.0, c], device=device).repeat(N, 1))
        self.pos_before = [torch.zeros(2, device=device) for _ in range(10)]
        self.goal_before = [torch.zeros(2, device=device) for _ in range(10)]
        self.box_before = torch.zeros(2, device=device)
        self.obs_k = [torch.zeros(N, 38, device=device) for _ in range(10)]
        self.box_pos = torch.zeros(N, 2, device=device)
        self.box_quat = torch.zeros(N, 4, device=device)
        self.goals = [torch.zeros(N, 2, device=device) for _ in range(10)]
        self.actions = torch.zeros(N, 80, device=device)
        ar = torch.arange(N, device=device, dtype=torch.long)
        self.ant_indices_k = [11 * ar + k for k in range(10)]
        self.box_indices = 11 * ar + 10

    # ten_ant.py:810-884
    def reset_idx(self, env_ids, noise=None):
        ant_box_indices = torch.unique(torch.cat([idx[env_ids] for idx in self.ant_indices_k] +
                                                 [self.box_indices[env_ids]]).to(torch.int32))
        positions, velocities = self._draw_noise(len(env_ids), noise)
        for k in range(10):
            self.dof_pos[k][env_ids] = itu.tensor_clamp(self.initial_dof_pos[env_ids] + positions,
                                                        self.dof_limits_lower, self.dof_limits_upper)
            self.dof_vel[k][env_ids] = velocities
        ant_indices = torch.unique(torch.cat([idx[env_ids] for idx in self.ant_indices_k]).to(torch.int32))
        self.last.update(ant_box_indices=ant_box_indices, ant_indices=ant_indices,
                         noise_positions=positions.clone(), noise_velocities=velocities.clone(),
                         dof_pushed=self.dof_state.clone())
        self.pos_before = [self.root_states[k::11, :2].clone() for k in range(10)]
        self.box_before, _, self.goal_before = tm.ten_ant_box_goals(self.root_states[10::11, :])
        self.progress_buf[env_ids] = 0
        self.reset_buf[env_ids] = 0

    # ten_ant.py:712-808
    def compute_observations(self):
        c = self.cfg
        for k in range(10):
            self.obs_k[k][:] = tm.ant_observations_38(
                self.root_states[k::11, :], self.targets, self.inv_start_rot, self.dof_pos[k], self.dof_vel[k],
                self.dof_limits_lower, self.dof_limits_upper, c["dofVelocityScale"],
                self.actions[:, 8 * k:8 * k + 8], self.heading_vec, self.up_vec)
        bp, bq, self.goals = tm.ten_ant_box_goals(self.root_states[10::11, :])
        self.box_pos[:] = bp
        self.box_quat[:] = bq
        self.obs_buf = torch.cat(tuple(self.obs_k) + (self.box_pos, self.box_quat, self.box_targets), dim=-1)

    # ten_ant.py:635-710
    def compute_reward(self):
        c = self.cfg
        self.rew_buf[:], self.reset_buf[:] = tm.ten_ant_reward(
            self.obs_k, self.reset_buf, self.progress_buf, self.actions, c["upWeight"], c["actionsCost"],
            c["energyCost"], c["jointsAtLimitCost"], c["terminationHeight"], c["deathCost"],
            c["episodeLength"], self.pos_before, self.goal_before, self.box_quat, 0.0, 1.0, 0.0, 0.0, 500.0,
            self.box_targets_k, 500.0, self.goals)

    def step(self, actions, frame_root, frame_dof, noise=None):
        """BaseTask.step: pre_physics_step (ten_ant.py:886-891) -> [simulate] -> post_physics_step
        (ten_ant.py:894-926); the frame becomes visible at the refresh inside compute_observations."""
        self.last = {}
        self.actions = actions.clone().to(self.device)
        self.last["forces"] = self.actions * self.joint_gears * self.cfg["powerScale"]
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.last["env_ids"] = env_ids.clone()
        if len(env_ids) > 0:
            self.reset_idx(env_ids, noise)
        self.root_states.copy_(frame_root)      # gym.refresh_actor_root_state_tensor
        self.dof_state.copy_(frame_dof)         # gym.refresh_dof_state_tensor
        self.compute_observations()
        self.compute_reward()
        self.pos_before = [self.obs_k[k][:, :2].clone() for k in range(10)]
        self.box_before = self.box_pos[:, :2].clone()
        self.goal_before = list(self.goals)
        return self.obs_buf, self.rew_buf, self.reset_buf


class OneAntOracle(_TaskOracleBase):
    """reference agents/tasks/one_ant.py (OneAnt)"""

    def __init__(self, num_envs, cfg=None, device="cpu", initial_root=None):
        cfg = dict(DEFAULT_ENV_CFG, **(cfg or {}))
        self.cfg = cfg
        self._alloc_base(num_envs, 60, device)
        N = num_envs
        self.root_states = (one_ant_initial_root(N, device) if initial_root is None else initial_root.clone())
        self.initial_root_states = self.root_states.clone()
        self.initial_root_states[:, 7:13] = 0
        self.dof_state = torch.zeros(N * 8, 2, device=device)
        self.dof_pos = self.dof_state.view(N, 8, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, 8, 2)[..., 1]
        self.vec_sensor_tensor = torch.zeros(N * 4, 6, device=device).view(N, 24)
        self.dof_limits_lower, self.dof_limits_upper = ant_dof_limits(device)
        zero = torch.tensor([0.0], device=device)
        init = torch.zeros_like(self.dof_pos)
        self.initial_dof_pos = torch.where(self.dof_limits_lower > zero, self.dof_limits_lower,
                                           torch.where(self.dof_limits_upper < zero, self.dof_limits_upper, init))
        self.joint_gears = torch.full((8,), ANT_GEAR, device=device)
        self.ant_pos = torch.zeros(N, 2, device=device)
        self.box_pos = torch.zeros(N, 2, device=device)
        self.box_quat = torch.zeros(N, 4, device=device)
        self.up_vec = torch.tensor([0.0, 0.0, 1.0], device=device).repeat(N, 1)
        self.heading_vec = torch.tensor([1.0, 0.0, 0.0], device=device).repeat(N, 1)
        self.inv_start_rot = itu.quat_conjugate(torch.tensor([0.0, 0.0, 0.0, 1.0], device=device)).repeat(N, 1)
        self.basis_vec0 = self.heading_vec.clone()
        self.basis_vec1 = self.up_vec.clone()
        self.targets = torch.zeros(N, 3, device=device)
        self.box_targets = torch.zeros(N, 2, device=device)
        self.potentials = torch.tensor([-4 / cfg["dt"]], dtype=torch.float32, device=device).repeat(N)  # one_ant.py:144
        self.prev_potentials = self.potentials.clone()
        self.pos_before = torch.zeros(2, device=device)
        self.box_before = torch.zeros(2, device=device)
        self.actions = torch.zeros(N, 8, device=device)
        ar = torch.arange(N, device=device, dtype=torch.long)
        self.ant_indices = 2 * ar
        self.box_indices = 2 * ar + 1

    # one_ant.py:363-391
    def reset_idx(self, env_ids, noise=None):
        ant_box_indices = torch.unique(torch.cat([self.ant_indices[env_ids],
                                                  self.box_indices[env_ids]]).to(torch.int32))
        positions, velocities = self._draw_noise(len(env_ids), noise)
        self.dof_pos[env_ids] = itu.tensor_clamp(self.initial_dof_pos[env_ids] + positions,
                                                 self.dof_limits_lower, self.dof_limits_upper)
        self.dof_vel[env_ids] = velocities
        ant_indices = self.ant_indices[env_ids].to(torch.int32)
        self.last.update(ant_box_indices=ant_box_indices, ant_indices=ant_indices,
                         noise_positions=positions.clone(), noise_velocities=velocities.clone(),
                         dof_pushed=self.dof_state.clone())
        self.pos_before = self.root_states[0::2, :2].clone()
        self.box_before = self.root_states[1::2, :2].clone()
        self.progress_buf[env_ids] = 0
        self.reset_buf[env_ids] = 0

    def step(self, actions, frame_root, frame_dof, frame_sensor, noise=None):
        """one_ant.py:396-415"""
        c = self.cfg
        self.last = {}
        self.actions = actions.clone().to(self.device)
        self.last["forces"] = self.actions * self.joint_gears * c["powerScale"]
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.last["env_ids"] = env_ids.clone()
        if len(env_ids) > 0:
            self.reset_idx(env_ids, noise)
        self.root_states.copy_(frame_root)
        self.dof_state.copy_(frame_dof)
        self.vec_sensor_tensor.copy_(frame_sensor.view(self.num_envs, 24))
        # one_ant.py:346-361
        (self.obs_buf[:], self.potentials[:], self.prev_potentials[:], self.up_vec[:], self.heading_vec[:],
         self.ant_pos[:]) = tm.one_ant_observations(
            self.root_states[0::2, :], self.root_states[1::2, :], self.targets, self.potentials,
            self.inv_start_rot, self.dof_pos, self.dof_vel, self.dof_limits_lower, self.dof_limits_upper,
            c["dofVelocityScale"], self.vec_sensor_tensor, self.actions, c["dt"], c["contactForceScale"],
            self.basis_vec0, self.basis_vec1, 2)
        self.box_pos[:] = self.root_states[1::2, :2]
        self.box_quat[:] = self.root_states[1::2, 3:7]
        # one_ant.py:314-344
        self.rew_buf[:], self.reset_buf[:] = tm.one_ant_reward(
            self.obs_buf, self.reset_buf, self.progress_buf, self.actions, c["upWeight"], c["actionsCost"],
            c["energyCost"], c["jointsAtLimitCost"], c["terminationHeight"], c["deathCost"], c["episodeLength"],
            self.pos_before, self.box_before, self.ant_pos, self.box_pos, self.box_quat, 0.0, 1.0, 0.0, 1.0,
            500.0, self.box_targets, 500.0)
        self.pos_before = self.ant_pos[:, :2].clone()
        self.box_before = self.box_pos[:, :2].clone()
        return self.obs_buf, self.rew_buf, self.reset_buf


class IngenuityOracle(_TaskOracleBase):
    """reference agents/tasks/multi_ingenuity.py (MultiIngenuity), flat (N,24) actions."""

    A = 4

    def __init__(self, num_envs, cfg=None, device="cpu", initial_root=None):
        cfg = dict(DEFAULT_ENV_CFG, **(cfg or {}))
        self.cfg = cfg
        self._alloc_base(num_envs, 52, device)
        N = num_envs
        self.root_states = (ingenuity_initial_root(N, device) if initial_root is None else initial_root.clone())
        self.initial_root_states = self.root_states.clone()
        self.dof_state = torch.zeros(N * 16, 2, device=device)
        v = self.dof_state.view(N, -1, 2)
        self.dof_vel = [v[:, 4 * h:4 * h + 4, 1] for h in range(4)]
        self.obs_h = [torch.zeros(N, 13, device=device) for _ in range(4)]
        self.goals = [torch.tensor(g, dtype=torch.float32, device=device).repeat(N, 1)
                      for g in ([4, 2, 1], [4, -2, 1], [4, 6, 1], [4, -6, 1])]      # multi_ingenuity.py:103-106
        self.thrusts = torch.zeros((N, 8, 3), dtype=torch.float32, device=device)
        self.forces = torch.zeros((N, 24, 3), dtype=torch.float32, device=device)
        ar = torch.arange(N, device=device, dtype=torch.long)
        self.actor_indices_h = [4 * ar + h for h in range(4)]

    # multi_ingenuity.py:231-266
    def reset_idx(self, env_ids):
        for h in range(4):
            self.dof_vel[h][:, 1] = -50
            self.dof_vel[h][:, 3] = 50
        self.thrusts[env_ids] = 0.0
        self.forces[env_ids] = 0.0
        actor_indices = torch.unique(torch.cat([idx[env_ids] for idx in self.actor_indices_h]).to(torch.int32))
        self.last.update(actor_indices=actor_indices, dof_pushed=self.dof_state.clone(),
                         forces_after_reset=self.forces.clone())
        self.reset_buf[env_ids] = 0
        self.progress_buf[env_ids] = 0

    def step(self, actions, frame_root):
        """multi_ingenuity.py:268-349.  The reference never refreshes the root tensor; the frame is
        made visible where the simulator would have written it (after pre_physics_step)."""
        self.last = {}
        tm.ingenuity_thrust_forces(actions.to(self.device), self.thrusts, self.forces, self.cfg["dt"])
        self.last["forces"] = self.forces.clone()
        self.root_states.copy_(frame_root)
        self.progress_buf += 1
        env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.last["env_ids"] = env_ids.clone()
        if len(env_ids) > 0:
            self.reset_idx(env_ids)
        for h in range(4):
            self.obs_h[h][:] = self.root_states[h::4, :]
        self.obs_buf = torch.cat(tuple(self.obs_h), dim=-1)
        self.rew_buf[:], self.reset_buf[:] = tm.ingenuity_reward(
            self.obs_h, self.goals, self.reset_buf, self.progress_buf, self.cfg["episodeLength"])
        return self.obs_buf, self.rew_buf, self.reset_buf


# ----------------------------------------------------------------------------------------------
# wrappers: agents/tasks/agent_base/vec_task.py:121-139 and multi_vec_task.py:89-175
# ----------------------------------------------------------------------------------------------


def vec_task_step(task_step, actions, clip_actions=1.0, clip_obs=5.0):
    """VecTaskPython.step: clamp actions, step, clamp obs.  ``task_step(actions)`` -> (obs, rew, reset)."""
    obs, rew, reset = task_step(torch.clamp(actions, -clip_actions, clip_actions))
    return torch.clamp(obs, -clip_obs, clip_obs), rew, reset


def multi_vec_task_step(task_step, actions_list, num_agents, own_width, clip_actions=1.0, clip_obs=7.0):
    """MultiVecTaskPython.step generalised over (A, own width): hstack, clamp, step, clamp obs, per agent
    cat(own slice, shared tail); state replicated per agent; reward (N,A,1); done (N,A)."""
    actions = torch.hstack(tuple(actions_list))
    obs, rew, reset = task_step(torch.clamp(actions, -clip_actions, clip_actions))
    obs_buf = torch.clamp(obs, -clip_obs, clip_obs)
    tail = obs_buf[:, num_agents * own_width:]
    ant_obs = [torch.cat([obs_buf[:, a * own_width:(a + 1) * own_width], tail], dim=1) for a in range(num_agents)]
    obs_all = torch.transpose(torch.stack(ant_obs), 1, 0)
    state_all = torch.transpose(torch.stack([obs_buf] * num_agents), 1, 0)
    reward_all = torch.transpose(torch.stack([rew.unsqueeze(-1)] * num_agents), 1, 0)
    done_all = torch.transpose(torch.stack([reset] * num_agents), 1, 0)
    return obs_all, state_all, reward_all, done_all

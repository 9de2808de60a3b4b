"""Torch restatement of the reference's per-step task tensor functions (TEST INFRASTRUCTURE).

Each function follows the cited reference function op for op (same torch ops in the same
order on the same strided views) so that fp32 rounding on a given device equals the
reference's.  Device-agnostic: run on CPU it is the pinned oracle (bit-checked against the
reference itself by tests/golden/make_golden.py); run on CUDA tensors it is the "torch eager on
GPU" second oracle tier of SURVEY.md section 4 used for the ill-conditioned rewards.

The only deliberate deviation: ``abs(ant_push - 1)`` on a bool tensor (reference
``agents/tasks/one_ant.py:505``, ``agents/tasks/ten_ant.py:1074...1164``) raises on current torch;
the intended 0/1 factor is computed through ``.long()`` (value-identical, SURVEY finding 5).
"""
from typing import List, Tuple

import torch
from torch import Tensor

from . import isaac_torch_utils as itu

# ----------------------------------------------------------------------------------------------
# agents/utils/torch_jit_utils.py
# ----------------------------------------------------------------------------------------------


def compute_heading_and_up(torso_rotation, inv_start_rot, to_target, vec0, vec1, up_idx: int):
    """reference agents/utils/torch_jit_utils.py:13-28"""
    num_envs = torso_rotation.shape[0]
    target_dirs = itu.normalize(to_target)
    torso_quat = itu.quat_mul(torso_rotation, inv_start_rot)
    up_vec = itu.get_basis_vector(torso_quat, vec1).view(num_envs, 3)
    heading_vec = itu.get_basis_vector(torso_quat, vec0).view(num_envs, 3)
    up_proj = up_vec[:, up_idx]
    heading_proj = torch.bmm(heading_vec.view(num_envs, 1, 3), target_dirs.view(num_envs, 3, 1)).view(num_envs)
    return torso_quat, up_proj, heading_proj, up_vec, heading_vec


def compute_rot(torso_quat, velocity, ang_velocity, targets, torso_positions):
    """reference agents/utils/torch_jit_utils.py:31-42 (note the (z, x) atan2 arguments)"""
    vel_loc = itu.quat_rotate_inverse(torso_quat, velocity)
    angvel_loc = itu.quat_rotate_inverse(torso_quat, ang_velocity)
    roll, pitch, yaw = itu.get_euler_xyz(torso_quat)
    walk_target_angle = torch.atan2(targets[:, 2] - torso_positions[:, 2],
                                    targets[:, 0] - torso_positions[:, 0])
    angle_to_target = walk_target_angle - yaw
    return vel_loc, angvel_loc, roll, pitch, yaw, angle_to_target


def quat_axis(q, axis: int = 0):
    """reference agents/utils/torch_jit_utils.py:45-50"""
    basis_vec = torch.zeros(q.shape[0], 3, device=q.device)
    basis_vec[:, axis] = 1
    return itu.quat_rotate(q, basis_vec)


# ----------------------------------------------------------------------------------------------
# shared by ten_ant.py / one_ant.py
# ----------------------------------------------------------------------------------------------


def compute_box_quat(box_quat):
    """reference ten_ant.py:951-962 / one_ant.py:429-440"""
    qw = box_quat[:, 3].clone()
    qx = box_quat[:, 0].clone()
    qy = box_quat[:, 1].clone()
    qz = box_quat[:, 2].clone()
    x = 2 * (qx * qy + qw * qz)
    y = 1 - 2 * (qx * qx + qz * qz)
    z = 2 * (qy * qz - qw * qx)
    return x, y, z


def compute_box_quat_dist(x_goal: float, y_goal: float, z_goal: float, x, y, z):
    """reference ten_ant.py:964-973 / one_ant.py:442-451.  The second divisor is a Python float
    (math.sqrt of the goal norm), so it reaches torch as ``Tensor / Scalar``."""
    x_1 = x * x_goal
    y_1 = y * y_goal
    z_1 = z * z_goal
    goal_norm = (x_goal ** 2 + y_goal ** 2 + z_goal ** 2) ** 0.5
    return (x_1 + y_1 + z_1) / (torch.sqrt(x ** 2 + y ** 2 + z ** 2)) / goal_norm


def l2_dist(a, b):
    """reference ten_ant.py:975-985 / one_ant.py:453-463"""
    c = a - b
    c1 = c[:, 0].clone()
    c2 = c[:, 1].clone()
    c = c1 ** 2 + c2 ** 2
    return torch.sqrt(c)


# ----------------------------------------------------------------------------------------------
# agents/tasks/ten_ant.py
# ----------------------------------------------------------------------------------------------

TEN_ANT_GOAL_OFFSETS = (1.5, 4.5, 7.5, 10.5, 13.5)


def compute_box_angle(box_quat):
    """reference ten_ant.py:935-947"""
    qw = box_quat[:, 3].clone()
    qz = box_quat[:, 2].clone()
    y = 2 * qw * qz
    x = 1 - 2 * qz * qz
    return torch.atan(y / x)


def ten_ant_box_goals(box_root) -> Tuple[Tensor, Tensor, List[Tensor]]:
    """reference ten_ant.py:1353-1393 (compute_box_pos + compute_other_goal): box_pos, box_quat and
    the ten goals ``box_pos +/- c*(sin a, -cos a)`` in ant order 1..10."""
    box_pos = box_root[:, :2]
    box_quat = box_root[:, 3:7]
    angle = compute_box_angle(box_quat)
    sin_value = torch.sin(angle).unsqueeze(1)
    cos_value = (-torch.cos(angle)).unsqueeze(1)
    goal_dist_0 = torch.cat((sin_value, cos_value), dim=-1)
    goals = []
    for c in TEN_ANT_GOAL_OFFSETS:
        goals.append(box_pos + c * goal_dist_0)
        goals.append(box_pos - c * goal_dist_0)
    return box_pos, box_quat, goals


def ant_observations_38(root_states, targets, inv_start_rot, dof_pos, dof_vel, dof_limits_lower,
                        dof_limits_upper, dof_vel_scale: float, actions, basis_vec0, basis_vec1):
    """reference ten_ant.py:1304-1350 (compute_ant_observations; the four prints are side effects)"""
    torso_position = root_states[:, 0:3]
    torso_rotation = root_states[:, 3:7]
    velocity = root_states[:, 7:10]
    ang_velocity = root_states[:, 10:13]
    to_target = targets - torso_position
    to_target[:, 2] = 0.0
    torso_quat, up_proj, heading_proj, up_vec, heading_vec = compute_heading_and_up(
        torso_rotation, inv_start_rot, to_target, basis_vec0, basis_vec1, 2)
    vel_loc, angvel_loc, roll, pitch, yaw, angle_to_target = compute_rot(
        torso_quat, velocity, ang_velocity, targets, torso_position)
    dof_pos_scaled = itu.unscale(dof_pos, dof_limits_lower, dof_limits_upper)
    return torch.cat((torso_position, vel_loc, angvel_loc,
                      yaw.unsqueeze(-1), roll.unsqueeze(-1), angle_to_target.unsqueeze(-1),
                      up_proj.unsqueeze(-1), heading_proj.unsqueeze(-1), dof_pos_scaled,
                      dof_vel * dof_vel_scale, actions), dim=-1)


def ten_ant_reward(obs_k: List[Tensor], reset_buf, progress_buf, actions, up_weight: float,
                   actions_cost_scale: float, energy_cost_scale: float, joints_at_limit_cost_scale: float,
                   termination_height: float, death_cost: float, max_episode_length: float,
                   pos_before: List[Tensor], goal_before: List[Tensor], box_quat,
                   x_goal: float, y_goal: float, z_goal: float, quat_reward_scale: float,
                   ant_dist_reward_scale: float, box_targets_k: List[Tensor],
                   goal_dist_reward_scale: float, goals: List[Tensor]):
    """reference ten_ant.py:988-1301 (compute_ant_reward).  ``heading_reward_k`` (lines 1184-1232) is
    computed by the reference but never added to the total, so it is omitted."""
    x, y, z = compute_box_quat(box_quat)
    quat_dist = compute_box_quat_dist(x_goal, y_goal, z_goal, x, y, z)
    quat_reward = quat_reward_scale * quat_dist

    ant_dist_reward = None
    goal_dist_reward = None
    goal_arrive_reward = None
    arrive = []
    for k in range(10):
        d_now = l2_dist(obs_k[k][:, :2], goals[k])
        ant_push = (d_now < 1.5)
        ant_push = abs(ant_push.long() - 1)
        ant_dist = l2_dist(pos_before[k], goal_before[k]) - l2_dist(obs_k[k][:, :2], goals[k])
        adr = ant_dist_reward_scale * ant_dist * ant_push
        goal_dist_before = l2_dist(box_targets_k[k], goal_before[k])
        goal_dist = l2_dist(box_targets_k[k], goals[k])
        goal_arrive = goal_dist < 0.5
        gdr = goal_dist_reward_scale * (goal_dist_before - goal_dist)
        gar = 2 * goal_arrive
        arrive.append(goal_arrive)
        ant_dist_reward = adr if ant_dist_reward is None else ant_dist_reward + adr
        goal_dist_reward = gdr if goal_dist_reward is None else goal_dist_reward + gdr
        goal_arrive_reward = gar if goal_arrive_reward is None else goal_arrive_reward + gar

    quat_arrive = quat_dist > 0.9
    success_reward = quat_arrive
    for k in range(10):
        success_reward = success_reward * arrive[k]
    success_reward = success_reward * 100

    up_sum = None
    for k in range(10):
        up_reward = torch.zeros_like(obs_k[k][:, 13])
        up_reward = torch.where(obs_k[k][:, 12] > 0.93, up_reward + up_weight, up_reward)
        up_sum = up_reward if up_sum is None else up_sum + up_reward
    up_reward = up_sum * 10

    actions_cost = torch.sum(actions ** 2, dim=-1)
    electricity_cost = None
    dof_at_limit_cost = None
    for k in range(10):
        ec = torch.sum(torch.abs(actions[:, 8 * k:8 * k + 8] * obs_k[k][:, 22:30]), dim=-1)
        dl = torch.sum(obs_k[k][:, 14:22] > 0.99, dim=-1)
        electricity_cost = ec if electricity_cost is None else electricity_cost + ec
        dof_at_limit_cost = dl if dof_at_limit_cost is None else dof_at_limit_cost + dl

    alive_reward = torch.ones_like(ant_dist_reward) * 5
    total_reward = alive_reward + up_reward + quat_reward + ant_dist_reward + goal_dist_reward + \
        goal_arrive_reward + success_reward - actions_cost_scale * actions_cost - \
        energy_cost_scale * electricity_cost - dof_at_limit_cost * joints_at_limit_cost_scale

    fallen = obs_k[0][:, 2] < termination_height
    for k in range(1, 10):
        fallen = fallen + (obs_k[k][:, 2] < termination_height)
    total_reward = torch.where(fallen, torch.ones_like(total_reward) * death_cost, total_reward)
    reset = torch.where(fallen, torch.ones_like(reset_buf), reset_buf)
    reset = torch.where(progress_buf >= max_episode_length - 1, torch.ones_like(reset_buf), reset)
    return total_reward, reset


# ----------------------------------------------------------------------------------------------
# agents/tasks/one_ant.py
# ----------------------------------------------------------------------------------------------


def one_ant_observations(root_states, root_states_box, targets, potentials, inv_start_rot, dof_pos,
                         dof_vel, dof_limits_lower, dof_limits_upper, dof_vel_scale: float,
                         sensor_force_torques, actions, dt: float, contact_force_scale: float,
                         basis_vec0, basis_vec1, up_axis_idx: int):
    """reference one_ant.py:563-618"""
    torso_position = root_states[:, 0:3]
    torso_rotation = root_states[:, 3:7]
    velocity = root_states[:, 7:10]
    ang_velocity = root_states[:, 10:13]
    ant_pos = root_states[:, 0:2]
    to_target = targets - torso_position
    to_target[:, 2] = 0.0
    box_torso_position = root_states_box[:, 0:3]
    to_target_box = targets - box_torso_position
    to_target_box[:, 2] = 0.0
    prev_potentials_new = potentials.clone()
    potentials = -torch.norm(to_target_box, p=2, dim=-1) / dt
    torso_quat, up_proj, heading_proj, up_vec, heading_vec = compute_heading_and_up(
        torso_rotation, inv_start_rot, to_target, basis_vec0, basis_vec1, 2)
    vel_loc, angvel_loc, roll, pitch, yaw, angle_to_target = compute_rot(
        torso_quat, velocity, ang_velocity, targets, torso_position)
    dof_pos_scaled = itu.unscale(dof_pos, dof_limits_lower, dof_limits_upper)
    obs = torch.cat((torso_position[:, up_axis_idx].view(-1, 1), vel_loc, angvel_loc,
                     yaw.unsqueeze(-1), roll.unsqueeze(-1), angle_to_target.unsqueeze(-1),
                     up_proj.unsqueeze(-1), heading_proj.unsqueeze(-1), dof_pos_scaled,
                     dof_vel * dof_vel_scale, sensor_force_torques.view(-1, 24) * contact_force_scale,
                     actions), dim=-1)
    return obs, potentials, prev_potentials_new, up_vec, heading_vec, ant_pos


def one_ant_reward(obs_buf, reset_buf, progress_buf, actions, up_weight: float,
                   actions_cost_scale: float, energy_cost_scale: float, joints_at_limit_cost_scale: float,
                   termination_height: float, death_cost: float, max_episode_length: float,
                   pos_before, box_before, ant_pos, box_pos, box_quat,
                   x_goal: float, y_goal: float, z_goal: float, quat_reward_scale: float,
                   ant_dist_reward_scale: float, box_targets, goal_dist_reward_scale: float):
    """reference one_ant.py:465-560 (heading_reward / progress_reward are computed there but unused)"""
    x, y, z = compute_box_quat(box_quat)
    quat_dist = compute_box_quat_dist(x_goal, y_goal, z_goal, x, y, z)
    quat_reward = quat_reward_scale * quat_dist
    ant_push = l2_dist(ant_pos, box_pos) < 1.5
    ant_push = abs(ant_push.long() - 1)
    ant_dist = l2_dist(pos_before, box_before) - l2_dist(ant_pos, box_pos)
    ant_dist_reward = ant_dist_reward_scale * ant_dist * ant_push
    goal_dist_before = l2_dist(box_targets, box_before)
    goal_dist = l2_dist(box_targets, box_pos)
    goal_arrive = goal_dist < 0.5
    goal_dist_reward = goal_dist_reward_scale * (goal_dist_before - goal_dist)
    goal_arrive_reward = 2 * goal_arrive
    quat_arrive = quat_dist > 0.9
    success_reward = quat_arrive * goal_arrive * 10
    up_reward = torch.zeros_like(obs_buf[:, 11])
    up_reward = torch.where(obs_buf[:, 10] > 0.93, up_reward + up_weight, up_reward)
    actions_cost = torch.sum(actions ** 2, dim=-1)
    electricity_cost = torch.sum(torch.abs(actions * obs_buf[:, 20:28]), dim=-1)
    dof_at_limit_cost = torch.sum(obs_buf[:, 12:20] > 0.99, dim=-1)
    alive_reward = torch.ones_like(up_reward) * 0.5
    total_reward = alive_reward + up_reward + quat_reward + ant_dist_reward + goal_dist_reward + \
        goal_arrive_reward + success_reward - actions_cost_scale * actions_cost - \
        energy_cost_scale * electricity_cost - dof_at_limit_cost * joints_at_limit_cost_scale
    total_reward = torch.where(obs_buf[:, 0] < termination_height, torch.ones_like(total_reward) * death_cost,
                               total_reward)
    reset = torch.where(obs_buf[:, 0] < termination_height, torch.ones_like(reset_buf), reset_buf)
    reset = torch.where(progress_buf >= max_episode_length - 1, torch.ones_like(reset_buf), reset)
    return total_reward, reset


# ----------------------------------------------------------------------------------------------
# agents/tasks/multi_ingenuity.py
# ----------------------------------------------------------------------------------------------


def ingenuity_reward(obs_h: List[Tensor], goals_h: List[Tensor], reset_buf, progress_buf,
                     max_episode_length: float):
    """reference multi_ingenuity.py:381-453 (compute_ingenuity_reward).  ``reset`` does NOT carry
    the old reset_buf (``die`` starts from zeros)."""
    target_dist = []
    pos_reward = None
    for h in range(4):
        root_positions = obs_h[h][:, :3]
        td = torch.sqrt(torch.square(goals_h[h] - root_positions).sum(-1))
        pr = 1.0 / (1.0 + td * td)
        target_dist.append(td)
        pos_reward = pr if pos_reward is None else pos_reward + pr
    up_reward = None
    for h in range(4):
        ups = quat_axis(obs_h[h][:, 3:7], 2)
        tiltage = torch.abs(1 - ups[..., 2])
        ur = 5.0 / (1.0 + tiltage * tiltage)
        up_reward = ur if up_reward is None else up_reward + ur
    spinnage_reward = None
    for h in range(4):
        spinnage = torch.abs(obs_h[h][:, 12])
        sr = 1.0 / (1.0 + spinnage * spinnage)
        spinnage_reward = sr if spinnage_reward is None else spinnage_reward + sr
    reward = pos_reward + pos_reward * (up_reward + spinnage_reward)
    ones = torch.ones_like(reset_buf)
    die = torch.zeros_like(reset_buf)
    die = torch.where((target_dist[0] > 8.0) | (target_dist[1] > 8.0) | (target_dist[2] > 8.0) |
                      (target_dist[3] > 8.0), ones, die)
    die = torch.where((obs_h[0][:, 2] < 0.5) | (obs_h[1][:, 2] < 0.5) | (obs_h[2][:, 2] < 0.5) |
                      (obs_h[3][:, 2] < 0.5), ones, die)
    reset = torch.where(progress_buf >= max_episode_length - 1, ones, die)
    return reward, reset


def ingenuity_thrust_forces(actions, thrusts, forces, dt: float, thrust_upper_limit: float = 2000.0,
                            thrust_lateral_component: float = 0.2):
    """reference multi_ingenuity.py:268-339 (pre_physics_step), in place on ``thrusts`` (N,8,3)
    and ``forces`` (N,24,3)."""
    thrust_action_speed_scale = 2000
    for h in range(4):
        b = 6 * h
        va = torch.clamp(actions[:, b + 2] * thrust_action_speed_scale, -thrust_upper_limit, thrust_upper_limit)
        vb = torch.clamp(actions[:, b + 5] * thrust_action_speed_scale, -thrust_upper_limit, thrust_upper_limit)
        la = torch.clamp(actions[:, b + 0:b + 2], -thrust_lateral_component, thrust_lateral_component)
        lb = torch.clamp(actions[:, b + 3:b + 5], -thrust_lateral_component, thrust_lateral_component)
        thrusts[:, 2 * h, 2] = dt * va
        thrusts[:, 2 * h, 0:2] = thrusts[:, 2 * h, 2, None] * la
        thrusts[:, 2 * h + 1, 2] = dt * vb
        thrusts[:, 2 * h + 1, 0:2] = thrusts[:, 2 * h + 1, 2, None] * lb
    for r, body in enumerate((1, 3, 7, 9, 13, 15, 19, 21)):
        forces[:, body] = thrusts[:, r]
    return forces

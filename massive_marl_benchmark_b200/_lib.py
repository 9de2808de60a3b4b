"""ctypes binding of libmmb_b200.so (include/mmb.h).  The one place Python touches the C ABI.

There is NO fallback: if the shared library is missing or a call fails, this raises.  The host
classes never route around the kernels (no CPU / eager path in the product).
"""
import ctypes as C
import math
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMB_LIB_PATH") or os.path.join(_HERE, "libmmb_b200.so")   # MMB_LIB_PATH: diagnostics builds (tools/probe)
ABI_VERSION = 4
MAX_GATHER_FIELDS = 16

FLAVOR_CUDA, FLAVOR_CPU = 0, 1
TASK_TEN_ANT, TASK_ONE_ANT, TASK_INGENUITY = 0, 1, 2

c_f, c_i32, c_i64, c_u64, c_vp, c_d = C.c_float, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p, C.c_double


class MmbError(RuntimeError):
    pass


class AntConsts(C.Structure):
    _fields_ = [(n, c_f) for n in (
        "dof_vel_scale", "contact_force_scale", "power_scale", "up_weight", "actions_cost", "energy_cost",
        "joints_at_limit_cost", "death_cost", "termination_height", "dt", "max_episode_length",
        "quat_reward_scale", "ant_dist_reward_scale", "goal_dist_reward_scale", "x_goal", "y_goal", "z_goal")] + [
        ("dof_lower", c_f * 8), ("dof_upper", c_f * 8), ("joint_gears", c_f * 8), ("inv_start_rot", c_f * 4),
        ("initial_dof_pos", c_f * 8)]


class TenAntParams(C.Structure):
    _fields_ = [
        ("num_envs", c_i32), ("num_frames", c_i32), ("flavor", c_i32), ("obs_layout", c_i32),
        ("root", c_vp), ("root_frame_stride", c_i64), ("dof", c_vp), ("dof_frame_stride", c_i64),
        ("actions", c_vp), ("actions_frame_stride", c_i64), ("clip_actions", c_f), ("clip_obs", c_f),
        ("pos_before", c_vp), ("goal_before", c_vp), ("box_before", c_vp), ("progress_buf", c_vp), ("reset_buf", c_vp),
        ("obs_raw", c_vp), ("obs_raw_frame_stride", c_i64), ("obs", c_vp), ("obs_frame_stride", c_i64),
        ("share_obs", c_vp), ("share_obs_frame_stride", c_i64), ("rewards", c_vp), ("rewards_frame_stride", c_i64),
        ("dones_i64", c_vp), ("dones_i64_frame_stride", c_i64), ("dones_u8", c_vp), ("dones_u8_frame_stride", c_i64),
        ("forces", c_vp), ("forces_frame_stride", c_i64), ("scratch", c_vp), ("overlap_prev", c_i32), ("scratch_per_set", c_i32),
        ("obs_agent_stride", c_i64), ("prev_root", c_vp), ("agent_actions", c_vp * 10),
        ("gae_values", c_vp), ("gae_values_frame_stride", c_i64), ("gae_last_values", c_vp),
        ("gae_returns", c_vp), ("gae_returns_frame_stride", c_i64),
        ("gae_advantages", c_vp), ("gae_advantages_frame_stride", c_i64),
        ("gae_stats", c_vp), ("gae_scratch", c_vp), ("gae_gamma", c_f), ("gae_lam", c_f),
        ("c", AntConsts)]


class OneAntParams(C.Structure):
    _fields_ = [
        ("num_envs", c_i32), ("num_frames", c_i32), ("flavor", c_i32), ("reserved0", c_i32),
        ("root", c_vp), ("root_frame_stride", c_i64), ("dof", c_vp), ("dof_frame_stride", c_i64),
        ("sensor", c_vp), ("sensor_frame_stride", c_i64), ("actions", c_vp), ("actions_frame_stride", c_i64),
        ("clip_actions", c_f), ("clip_obs", c_f),
        ("pos_before", c_vp), ("box_before", c_vp), ("potentials", c_vp), ("prev_potentials", c_vp),
        ("progress_buf", c_vp), ("reset_buf", c_vp),
        ("obs_raw", c_vp), ("obs_raw_frame_stride", c_i64), ("obs", c_vp), ("obs_frame_stride", c_i64),
        ("rewards", c_vp), ("rewards_frame_stride", c_i64), ("dones_i64", c_vp), ("dones_i64_frame_stride", c_i64),
        ("dones_u8", c_vp), ("dones_u8_frame_stride", c_i64), ("forces", c_vp), ("forces_frame_stride", c_i64),
        ("up_vec", c_vp), ("heading_vec", c_vp), ("ant_pos", c_vp), ("box_pos", c_vp), ("box_quat", c_vp),
        ("c", AntConsts)]


class IngenuityParams(C.Structure):
    _fields_ = [
        ("num_envs", c_i32), ("num_frames", c_i32), ("flavor", c_i32), ("obs_layout", c_i32),
        ("root", c_vp), ("root_frame_stride", c_i64), ("actions", c_vp), ("actions_frame_stride", c_i64),
        ("clip_actions", c_f), ("clip_obs", c_f), ("dt", c_f), ("max_episode_length", c_f),
        ("thrust_upper_limit", c_f), ("thrust_lateral_component", c_f), ("thrust_action_speed_scale", c_f),
        ("goals", (c_f * 3) * 4), ("progress_buf", c_vp), ("reset_buf", c_vp),
        ("obs_raw", c_vp), ("obs_raw_frame_stride", c_i64), ("obs", c_vp), ("obs_frame_stride", c_i64),
        ("rewards", c_vp), ("rewards_frame_stride", c_i64), ("dones_i64", c_vp), ("dones_i64_frame_stride", c_i64),
        ("dones_u8", c_vp), ("dones_u8_frame_stride", c_i64), ("forces", c_vp), ("forces_frame_stride", c_i64),
        ("forces_state", c_vp)]


class ResetParams(C.Structure):
    _fields_ = [
        ("task", c_i32), ("num_envs", c_i32), ("num_rows", c_i32), ("noise_mode", c_i32),
        ("flags_i64", c_vp), ("flags_i64_row_stride", c_i64), ("flags_u8", c_vp), ("flags_u8_row_stride", c_i64),
        ("env_ids", c_vp), ("env_ids_row_stride", c_i64), ("index_a", c_vp), ("index_a_row_stride", c_i64),
        ("index_b", c_vp), ("index_b_row_stride", c_i64), ("counts", c_vp),
        ("dof_state", c_vp), ("dof_state_row_stride", c_i64),
        ("noise_pos", c_vp), ("noise_vel", c_vp), ("noise_row_stride", c_i64),
        ("seed", c_u64), ("step", c_u64), ("forces_state", c_vp), ("scan_scratch", c_vp), ("step_counter", c_vp),
        ("c", AntConsts)]


MAX_COPY_SEGS = 12


class CopySeg(C.Structure):
    _fields_ = [("dst", c_vp), ("src", c_vp), ("n0", c_i32), ("n1", c_i32), ("n2", c_i32), ("_pad", c_i32),
                ("dst_s0", c_i64), ("dst_s1", c_i64), ("src_s0", c_i64), ("src_s1", c_i64)]


class CopyGroupParams(C.Structure):
    _fields_ = [("count", c_i32), ("_pad", c_i32), ("seg", CopySeg * MAX_COPY_SEGS)]


class RolloutAddParams(C.Structure):
    _fields_ = [("num_envs", c_i32), ("obs_dim", c_i32), ("states_dim", c_i32), ("act_dim", c_i32)] + [
        (n, c_vp) for n in ("observations", "states", "actions", "rewards", "dones", "values", "actions_log_prob",
                            "mu", "sigma", "dst_observations", "dst_states", "dst_actions", "dst_rewards",
                            "dst_dones", "dst_values", "dst_actions_log_prob", "dst_mu", "dst_sigma")] + [("values_stride", c_i64)]


ACT_COUNTER_WORDS = 2 + 64
MAX_RANKS = 16
MAX_GROUP = 16
STAT_SLOTS, STAT_SLOT_STRIDE = 32, 16
ADV_STATS_EXT_DOUBLES = 4 + STAT_SLOTS * STAT_SLOT_STRIDE
NORM_CLEAR, NORM_SLOTS = 1, 2


class Xchg(C.Structure):
    _fields_ = [("world", c_i32), ("rank", c_i32), ("slots", c_i32), ("timeout_ms", c_i32), ("state", c_vp),
                ("mailbox", c_vp * MAX_RANKS)]


class GaePpoParams(C.Structure):
    _fields_ = [("num_envs", c_i32), ("num_steps", c_i32), ("rewards", c_vp), ("values", c_vp), ("dones", c_vp),
                ("last_values", c_vp), ("gamma", c_d), ("lam", c_d), ("returns", c_vp), ("advantages", c_vp),
                ("stats", c_vp)]


class EpisodeParams(C.Structure):
    _fields_ = [("num_envs", c_i32), ("num_steps", c_i32), ("window", c_i32), ("_pad", c_i32),
                ("rewards", c_vp), ("rewards_row_stride", c_i64), ("dones_u8", c_vp), ("dones_u8_row_stride", c_i64),
                ("dones_i64", c_vp), ("dones_i64_row_stride", c_i64), ("cur_reward_sum", c_vp), ("cur_episode_length", c_vp),
                ("ep_reward", c_vp), ("ep_length", c_vp), ("reward_ring", c_vp), ("length_ring", c_vp), ("state", c_vp)]


class GaussianActParams(C.Structure):
    _fields_ = [("num_rows", c_i32), ("act_dim", c_i32), ("deterministic", c_i32), ("std_group_rows", c_i32),
                ("mean", c_vp), ("mean_stride", c_i64), ("std", c_vp), ("noise", c_vp), ("seed", c_u64), ("step", c_u64),
                ("actions", c_vp), ("logp_sum", c_vp), ("logp_per_dim", c_vp), ("sigma_src", c_vp), ("sigma_out", c_vp), ("step_counter", c_vp)]


class PpoLossParams(C.Structure):
    _fields_ = [("num_rows", c_i32), ("act_dim", c_i32), ("use_clipped_value_loss", c_i32), ("_pad", c_i32),
                ("mu", c_vp), ("mu_stride", c_i64), ("log_std", c_vp), ("actions", c_vp), ("old_logp", c_vp),
                ("advantages", c_vp), ("value", c_vp), ("target_values", c_vp), ("returns", c_vp), ("old_mu", c_vp),
                ("old_sigma", c_vp), ("clip_param", c_f), ("ratio_lo", c_f), ("ratio_hi", c_f), ("value_loss_coef", c_f),
                ("entropy_coef", c_f), ("k_log_2pi", c_f), ("logp", c_vp), ("grad_mu", c_vp), ("grad_value", c_vp), ("sums", c_vp),
                ("out", c_vp), ("ticket", c_vp)]


class MappoLossParams(C.Structure):
    _fields_ = [("num_rows", c_i32), ("act_dim", c_i32), ("use_huber_loss", c_i32), ("use_clipped_value_loss", c_i32),
                ("use_value_active_masks", c_i32), ("use_policy_active_masks", c_i32),
                ("mean", c_vp), ("mean_stride", c_i64), ("std", c_vp), ("actions", c_vp), ("old_logp", c_vp),
                ("adv_targ", c_vp), ("values", c_vp), ("value_preds", c_vp), ("returns", c_vp), ("active_masks", c_vp),
                ("mask_sum", c_vp), ("ret_mean", c_vp), ("ret_var", c_vp), ("ret_mean_orig", c_vp), ("ret_var_orig", c_vp),
                ("clip_param", c_f), ("ratio_lo", c_f), ("ratio_hi", c_f), ("huber_delta", c_f),
                ("imp_weights", c_vp), ("logp", c_vp), ("grad_mean", c_vp), ("grad_values", c_vp), ("sums", c_vp),
                ("out", c_vp), ("ticket", c_vp)]


class GaeMarlParams(C.Structure):
    _fields_ = [
        ("num_envs", c_i32), ("num_steps", c_i32), ("num_agents", c_i32),
        ("use_gae", c_i32), ("use_proper_time_limits", c_i32), ("use_denorm", c_i32), ("use_popart", c_i32),
        ("rewards", c_vp), ("rew_t", c_i64), ("rew_e", c_i64), ("rew_a", c_i64),
        ("value_preds", c_vp), ("val_t", c_i64), ("val_e", c_i64), ("val_a", c_i64),
        ("masks", c_vp), ("msk_t", c_i64), ("msk_e", c_i64), ("msk_a", c_i64),
        ("bad_masks", c_vp), ("bad_t", c_i64), ("bad_e", c_i64), ("bad_a", c_i64),
        ("next_value", c_vp), ("nv_e", c_i64), ("nv_a", c_i64),
        ("returns", c_vp), ("ret_t", c_i64), ("ret_e", c_i64), ("ret_a", c_i64),
        ("advantages", c_vp), ("adv_t", c_i64), ("adv_e", c_i64), ("adv_a", c_i64),
        ("denorm_mean", c_vp), ("denorm_var", c_vp), ("gamma", c_d), ("gae_lambda", c_d), ("stats", c_vp)]


class GatherParams(C.Structure):
    _fields_ = [
        ("num_fields", c_i32), ("index_mode", c_i32), ("total", c_i64), ("batch_start", c_i64), ("batch_size", c_i64),
        ("indices", c_vp), ("seed", c_u64), ("indices_out", c_vp),
        ("src", c_vp * MAX_GATHER_FIELDS), ("dst", c_vp * MAX_GATHER_FIELDS), ("row_bytes", c_i32 * MAX_GATHER_FIELDS),
        ("group", c_i32), ("_reserved", c_i32)]


ADAM_MAX_GROUPS = 64


class AdamParams(C.Structure):
    _fields_ = [("num_groups", c_i32), ("_pad", c_i32), ("total", c_i64), ("group_start", c_i64 * ADAM_MAX_GROUPS),
                ("params", c_vp), ("grads", c_vp), ("exp_avg", c_vp), ("exp_avg_sq", c_vp), ("sumsq", c_vp),
                ("step_size", c_f * ADAM_MAX_GROUPS), ("bc2_sqrt", c_f * ADAM_MAX_GROUPS), ("eps", c_f * ADAM_MAX_GROUPS),
                ("weight_decay", c_f * ADAM_MAX_GROUPS), ("max_grad_norm", c_f * ADAM_MAX_GROUPS),
                ("one_minus_beta1", c_f), ("beta2", c_f), ("one_minus_beta2", c_f), ("_pad2", c_f)]


class MlpLayerParams(C.Structure):
    _fields_ = [("M", c_i32), ("N", c_i32), ("K", c_i32), ("Mpad", c_i32), ("Kpad", c_i32), ("Npad", c_i32),
                ("n_tile", c_i32), ("epilogue", c_i32), ("x", c_vp), ("w", c_vp), ("bias", c_vp), ("ln_gamma", c_vp),
                ("ln_beta", c_vp), ("ln_eps", c_f), ("stages", c_i32), ("y", c_vp), ("y_stride", c_i64),
                ("overlap_prev", c_i32), ("operand_type", c_i32),
                ("ln_in_stats", c_vp), ("ln_in_parts", c_i32), ("ln_in_n", c_i32), ("ln_in_eps", c_f), ("ln_c", c_vp),
                ("ln_out_stats", c_vp)]


# name -> (restype, argtypes); every symbol include/mmb.h declares
SYMBOLS = {
    "mmb_abi_version": (c_i32, []),
    "mmb_strerror": (C.c_char_p, [c_i32]),
    "mmb_launch_count": (c_u64, []),
    "mmb_profile_enable": (c_i32, [c_i32]),
    "mmb_profile_collect": (c_i32, [c_i32, C.POINTER(c_d), C.POINTER(c_i64)]),
    "mmb_ten_ant_step": (c_i32, [C.POINTER(TenAntParams), c_vp]),
    "mmb_ten_ant_load_carry": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "mmb_one_ant_step": (c_i32, [C.POINTER(OneAntParams), c_vp]),
    "mmb_ingenuity_step": (c_i32, [C.POINTER(IngenuityParams), c_vp]),
    "mmb_reset_compact": (c_i32, [C.POINTER(ResetParams), c_vp]),
    "mmb_ten_ant_env_step": (c_i32, [C.POINTER(ResetParams), C.POINTER(TenAntParams), c_vp]),
    "mmb_rollout_add": (c_i32, [C.POINTER(RolloutAddParams), c_vp]),
    "mmb_copy_group": (c_i32, [C.POINTER(CopyGroupParams), c_vp]),
    "mmb_gae_ppo": (c_i32, [C.POINTER(GaePpoParams), c_vp]),
    "mmb_adv_normalize": (c_i32, [c_vp, c_i64, c_vp, c_f, c_i32, c_vp]),
    "mmb_adv_normalize_xchg": (c_i32, [c_vp, c_i64, c_vp, C.POINTER(Xchg), c_f, c_i32, c_vp]),
    "mmb_xchg_mailbox_bytes": (c_i64, [c_i32, c_i32]),
    "mmb_xchg_alloc": (c_i32, [c_i64, C.POINTER(c_vp), C.POINTER(C.c_uint8)]),
    "mmb_xchg_open": (c_i32, [C.POINTER(C.c_uint8), C.POINTER(c_vp)]),
    "mmb_xchg_close": (c_i32, [c_vp]),
    "mmb_xchg_free": (c_i32, [c_vp]),
    "mmb_rollout_statistics": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "mmb_gaussian_act": (c_i32, [C.POINTER(GaussianActParams), c_vp]),
    "mmb_ppo_loss": (c_i32, [C.POINTER(PpoLossParams), c_vp]),
    "mmb_mappo_loss": (c_i32, [C.POINTER(MappoLossParams), c_vp]),
    "mmb_episode_update": (c_i32, [C.POINTER(EpisodeParams), c_vp]),
    "mmb_gae_marl": (c_i32, [C.POINTER(GaeMarlParams), c_vp]),
    "mmb_marl_masks": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "mmb_shuffle_gather": (c_i32, [C.POINTER(GatherParams), c_vp]),
    "mmb_permutation": (c_i32, [c_i64, c_u64, c_i32, c_vp, c_vp]),
    "mmb_grad_sumsq_group": (c_i32, [C.POINTER(AdamParams), c_vp]),
    "mmb_adam_group": (c_i32, [C.POINTER(AdamParams), c_vp]),
    "mmb_mlp_layer": (c_i32, [C.POINTER(MlpLayerParams), c_vp]),
    "mmb_mlp_chain_trace": (c_i32, [c_vp, c_i32]),
    "mmb_mlp_chain": (c_i32, [C.POINTER(MlpLayerParams), c_i32, c_i32, C.POINTER(c_vp), c_vp]),
    "mmb_mlp_debug_status": (c_i32, [C.POINTER(C.c_uint32)]),
    "mmb_mlp_layer_group": (c_i32, [C.POINTER(MlpLayerParams), c_i32, c_vp]),
    "mmb_ln_cast_group": (c_i32, [C.POINTER(c_vp), c_i32, c_i32, c_i32, c_i32, c_i32, C.POINTER(c_vp), C.POINTER(c_vp), c_f, c_i32,
                                  C.POINTER(c_vp), c_vp]),
    "mmb_ln_cast": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_f, c_i32, c_vp, c_vp]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/ into libmmb_b200.so with nvcc for sm_100a (in-tree, travels with the snapshot)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise MmbError("building libmmb_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded shared library; raises MmbError when it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmbError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                           "`make -C massive_marl_benchmark_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.mmb_abi_version() != ABI_VERSION:
            raise MmbError("ABI mismatch: library %d, binding %d" % (l.mmb_abi_version(), ABI_VERSION))
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise MmbError("%s failed: %s (%d)" % (what, lib().mmb_strerror(rc).decode(), rc))


def launch_count():
    return int(lib().mmb_launch_count())


KERNEL_IDS = ("ten_ant", "ten_ant_chain", "ten_ant_carry", "one_ant", "one_ant_chain", "ingenuity", "ingenuity_chain",
              "reset", "rollout_add", "gae_ppo", "adv_norm", "stats", "gae_marl", "masks", "gather", "perm", "mlp_layer",
              "ln_cast", "adv_norm_xchg", "episode_scan", "episode_ring", "gauss_act", "ppo_loss", "mappo_loss", "adam_norm", "adam")


def profile_enable(on=True):
    lib().mmb_profile_enable(1 if on else 0)


def profile_collect():
    """{kernel name: (total_ms, launches)} since the previous collect (synchronises the recorded events)."""
    out = {}
    for i, name in enumerate(KERNEL_IDS):
        ms, n = c_d(0.0), c_i64(0)
        check(lib().mmb_profile_collect(i, C.byref(ms), C.byref(n)), "mmb_profile_collect")
        if n.value:
            out[name] = (ms.value, n.value)
    return out


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device.  Through torch's raw accessor where it exists: building
    a `torch.cuda.Stream` object per call (`current_stream()`) costs ~10-18 us, which is most of a per-step call's host time."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def default_ant_consts(env_cfg=None, quat_reward_scale=0.0, dt=0.0166):
    """AntConsts from the `env:` block of cfg/<Task>.yaml (defaults = cfg/TenAnt.yaml:6,39-52) plus the
    per-task literals of ten_ant.py:55-59,199-201 / one_ant.py:56-60."""
    e = dict(episodeLength=1000, powerScale=1.0, upWeight=0.1, actionsCost=0.005, energyCost=0.05,
             dofVelocityScale=0.2, contactForceScale=0.1, jointsAtLimitCost=0.1, deathCost=-2.0,
             terminationHeight=0.31)
    if env_cfg:
        e.update({k: env_cfg[k] for k in e if k in env_cfg})
    c = AntConsts()
    c.dof_vel_scale = e["dofVelocityScale"]
    c.contact_force_scale = e["contactForceScale"]
    c.power_scale = e["powerScale"]
    c.up_weight = e["upWeight"]
    c.actions_cost = e["actionsCost"]
    c.energy_cost = e["energyCost"]
    c.joints_at_limit_cost = e["jointsAtLimitCost"]
    c.death_cost = e["deathCost"]
    c.termination_height = e["terminationHeight"]
    c.dt = dt
    c.max_episode_length = e["episodeLength"]
    c.quat_reward_scale = quat_reward_scale
    c.ant_dist_reward_scale = 500.0
    c.goal_dist_reward_scale = 500.0
    c.x_goal, c.y_goal, c.z_goal = 0.0, 1.0, 0.0
    from .synthetic import ANT_DOF_RANGE_DEG
    for j, (lo, hi) in enumerate(ANT_DOF_RANGE_DEG):
        lo_r, hi_r = math.radians(lo), math.radians(hi)
        c.dof_lower[j], c.dof_upper[j] = lo_r, hi_r
        c.joint_gears[j] = 15.0
        import numpy as np
        lo32, hi32 = np.float32(lo_r), np.float32(hi_r)
        c.initial_dof_pos[j] = float(lo32) if lo32 > 0 else (float(hi32) if hi32 < 0 else 0.0)
    c.inv_start_rot[0], c.inv_start_rot[1], c.inv_start_rot[2], c.inv_start_rot[3] = -0.0, -0.0, -0.0, 1.0
    return c

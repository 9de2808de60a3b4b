"""CPU: the oracle (oracle/) reproduces the committed golden vectors, which were generated from the
reference's own classes (tests/golden/make_golden.py).  This is what pins the oracle on every machine,
including the GPU box where /root/reference does not exist."""
import numpy as np
import torch

from conftest import load_golden
from oracle import storage_oracle as so
from oracle.task_oracle import IngenuityOracle, OneAntOracle, TenAntOracle, multi_vec_task_step, vec_task_step


def test_ten_ant_oracle_reproduces_golden():
    g = load_golden("ten_ant_n37")
    F, N = g["rew"].shape
    orc = TenAntOracle(N)
    for t in range(F):
        a = g["actions"][t]
        o_all, s_all, r_all, d_all = multi_vec_task_step(
            lambda act: orc.step(act, g["root"][t], g["dof"][t], noise=(g["noise_pos"][t], g["noise_vel"][t])),
            [a[:, 8 * k:8 * k + 8] for k in range(10)], 10, 38)
        n = int(g["n_reset"][t])
        assert torch.equal(orc.obs_buf, g["obs"][t]) and torch.equal(orc.rew_buf, g["rew"][t])
        assert torch.equal(orc.reset_buf, g["reset"][t]) and torch.equal(orc.progress_buf, g["progress"][t])
        assert torch.equal(orc.last["forces"], g["forces"][t])
        assert torch.equal(orc.last["env_ids"], g["env_ids"][t][:n])
        if n:
            assert torch.equal(orc.last["ant_box_indices"], g["ant_box_indices"][t][:11 * n])
            assert torch.equal(orc.last["ant_indices"], g["ant_indices"][t][:10 * n])
            assert torch.equal(orc.last["dof_pushed"], g["dof_pushed"][t])
        assert torch.equal(o_all, g["obs_all"][t]) and torch.equal(r_all, g["reward_all"][t])
        assert torch.equal(d_all, g["done_all"][t])
        assert torch.equal(s_all[:, 0], torch.clamp(g["obs"][t], -7.0, 7.0))
        if t == 0:
            orc.progress_buf[:] = g["progress_after0"]
    # sanity of the fixture itself: it exercises resets from falls and from the episode limit
    assert int(g["n_reset"][0]) == N and (g["progress"] == 0).any() and (g["rew"] == -2.0).any()
    assert (g["progress"][:, 3] >= 997).any()


def test_one_ant_oracle_reproduces_golden():
    g = load_golden("one_ant_n64")
    F, N = g["rew"].shape
    orc = OneAntOracle(N)
    for t in range(F):
        obs_c, rew, done = vec_task_step(
            lambda act: orc.step(act, g["root"][t], g["dof"][t], g["sensor"][t], noise=(g["noise_pos"][t], g["noise_vel"][t])),
            g["actions"][t], 1.0, 5.0)
        for name, key in (("obs_buf", "obs"), ("rew_buf", "rew"), ("reset_buf", "reset"), ("progress_buf", "progress"),
                          ("potentials", "potentials"), ("prev_potentials", "prev_potentials"), ("up_vec", "up_vec"),
                          ("heading_vec", "heading_vec")):
            assert torch.equal(getattr(orc, name), g[key][t]), name
        assert torch.equal(obs_c, g["obs_clamped"][t]) and torch.equal(orc.last["forces"], g["forces"][t])
        if t == 0:
            orc.progress_buf[:] = g["progress_after0"]
    assert (g["progress"][:, 5] >= 998).any() and (g["reset"][:, 5] == 1).any()


def test_ingenuity_oracle_reproduces_golden():
    g = load_golden("ingenuity_n33")
    F, N = g["rew"].shape
    orc = IngenuityOracle(N)
    for t in range(F):
        orc.step(g["actions"][t], g["root"][t])
        assert torch.equal(orc.obs_buf, g["obs"][t]) and torch.equal(orc.rew_buf, g["rew"][t])
        assert torch.equal(orc.reset_buf, g["reset"][t]) and torch.equal(orc.progress_buf, g["progress"][t])
        assert torch.equal(orc.last["forces"], g["forces"][t])
        n = int(g["n_reset"][t])
        if n:
            assert torch.equal(orc.last["actor_indices"], g["actor_indices"][t][:4 * n])
        if t == 0:
            orc.progress_buf[:] = g["progress_after0"]
    assert (g["progress"][:, 2] >= 999).any()
    assert (g["reset"] == 1).any() and (g["reset"] == 0).any()


def test_storage_oracle_reproduces_golden():
    g = load_golden("storage_ppo")
    T, N = g["in_rew"].shape
    rewards, values = g["in_rew"].view(T, N, 1), g["in_val"]
    ret, adv = so.ppo_compute_returns(rewards, values, g["dones_u8"], g["last_values"], 0.96, 0.95)
    assert torch.equal(ret, g["returns"]) and torch.equal(adv, g["advantages"])
    ml, mr = so.ppo_get_statistics(g["dones_u8"], rewards)
    assert torch.equal(ml, g["mean_len"]) and torch.equal(mr, g["mean_rew"])
    assert [len(p) for p in so.ppo_minibatch_partition(N, T, 4)] == g["part4_sizes"].tolist()
    p3 = so.ppo_minibatch_partition(N, T, 3)
    assert [len(p) for p in p3] == g["part3_sizes"].tolist() and p3[-1][-1] == int(g["part3_last"])
    assert str(g["overflow_msg"]) == "Rollout buffer overflow"


def test_marl_oracle_reproduces_golden():
    g = load_golden("buffer_marl")
    T = g["in_rewards"].shape[0]
    vp = torch.cat([g["in_value_preds"], torch.zeros_like(g["in_value_preds"][:1])])
    ret, vps = so.marl_compute_returns(g["in_rewards"], vp, g["buf_masks"], torch.ones_like(g["buf_masks"]), g["next_value"],
                                       0.96, 0.95, denorm=(g["popart_mean"], g["popart_var"]))
    assert torch.equal(ret, g["returns"]) and torch.equal(vps, g["value_preds_after"])
    adv = so.marl_advantages(ret, vps, denorm=(g["popart_mean"], g["popart_var"]))
    assert torch.equal(adv, g["advantages"])
    for tag, denorm, ptl in (("plain", None, False), ("ptl", (g["popart_mean"], g["popart_var"]), True)):
        r2, _ = so.marl_compute_returns(g["in_rewards"], vp, g["buf_masks"], g["bad_masks_" + tag], g["next_value"], 0.96,
                                        0.95, denorm=denorm, use_proper_time_limits=ptl)
        assert torch.equal(r2, g["returns_" + tag]), tag
    m, a = so.runner_insert_masks(g["runner_dones"])
    assert torch.equal(m, g["runner_masks"]) and torch.equal(a, g["runner_active_masks"])
    # buffer slot semantics (separated_buffer.py:67-85): obs/masks at step+1, actions/values/rewards at step
    assert torch.equal(g["buf_share_obs"][1:], g["in_share_obs"]) and torch.equal(g["buf_masks"][1:], g["in_masks"])


def test_isaac_helpers_self_consistency():
    """isaacgym.torch_utils is unpinned (absent third-party module): property checks on the restatement."""
    from oracle import isaac_torch_utils as itu
    gen = torch.Generator().manual_seed(0)
    q = torch.randn(256, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
    v = torch.randn(256, 3, generator=gen)
    back = itu.quat_rotate_inverse(q, itu.quat_rotate(q, v))
    assert torch.allclose(back, v, atol=1e-5)
    assert torch.allclose(itu.quat_rotate(q, v).norm(dim=-1), v.norm(dim=-1), atol=1e-5)
    ident = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(256, 1)
    assert torch.allclose(itu.quat_mul(q, ident), q, atol=1e-6)
    assert torch.allclose(itu.quat_mul(q, itu.quat_conjugate(q)), ident, atol=1e-6)
    # euler of axis-aligned rotations
    for axis, idx in ((0, 0), (2, 2)):
        ang = torch.tensor([0.3, 1.0, 2.5])
        qq = torch.zeros(3, 4); qq[:, axis] = torch.sin(ang / 2); qq[:, 3] = torch.cos(ang / 2)
        roll, pitch, yaw = itu.get_euler_xyz(qq)
        got = roll if axis == 0 else yaw
        assert torch.allclose(got, ang, atol=1e-6)
    x = torch.tensor([[0.5236, 1.7453]]); lo = torch.tensor([0.5236]); hi = torch.tensor([1.7453])
    assert torch.allclose(itu.unscale(x, lo, hi), torch.tensor([[-1.0, 1.0]]), atol=1e-6)


def test_ppo_loss_oracle_reproduces_golden():
    """oracle/ppo_loss_oracle.py against outputs of the reference's own evaluate() + PPO.update loss lines
    (tests/golden/ppo_loss.npz).  Transcendentals may differ by an ulp between CPUs, hence a tight allclose."""
    from oracle.ppo_loss_oracle import ppo_loss_oracle
    g = load_golden("ppo_loss")
    for tag in ("a8", "a80", "a8u"):
        mb = {k[len(tag) + 5:]: v for k, v in g.items() if k.startswith(tag + "__in_")}
        cfg = {k[len(tag) + 6:]: float(v) for k, v in g.items() if k.startswith(tag + "__cfg_")}
        cfg["use_clipped_value_loss"] = bool(cfg["use_clipped_value_loss"])
        out = ppo_loss_oracle(**mb, **cfg)
        for k, v in out.items():
            ref = g["%s__out_%s" % (tag, k)]
            assert torch.allclose(v, ref, rtol=2e-5, atol=2e-6 * float(ref.abs().max()) + 1e-9), (tag, k)
        ratio = torch.exp(out["logp"] - mb["old_logp"].squeeze())
        assert ((ratio < 0.8).any() and (ratio > 1.2).any() and ((ratio > 0.8) & (ratio < 1.2)).any())   # all clip regimes


def _loss_case(g, tag, bool_keys):
    mb = {k[len(tag) + 5:]: v for k, v in g.items() if k.startswith(tag + "__in_")}
    cfg = {k[len(tag) + 6:]: float(v) for k, v in g.items() if k.startswith(tag + "__cfg_")}
    for k in bool_keys:
        cfg[k] = bool(cfg[k])
    want = {k[len(tag) + 6:]: v for k, v in g.items() if k.startswith(tag + "__out_")}
    return mb, cfg, want


MAPPO_BOOLS = ("use_huber_loss", "use_clipped_value_loss", "use_value_active_masks", "use_policy_active_masks")


def test_mappo_loss_oracle_reproduces_golden():
    """oracle/mappo_loss_oracle.py against outputs of the reference's own MAPPO.ppo_update (tests/golden/mappo_loss.npz)."""
    from oracle.mappo_loss_oracle import mappo_loss_oracle
    g = load_golden("mappo_loss")
    for tag in ("cfg", "mask", "mse"):
        mb, cfg, want = _loss_case(g, tag, MAPPO_BOOLS)
        out = mappo_loss_oracle(**mb, **cfg)
        for k, v in out.items():
            assert torch.allclose(v, want[k], rtol=2e-5, atol=2e-6 * float(want[k].abs().max()) + 1e-9), (tag, k)
        imp = out["imp_weights"]
        assert (imp < 1 - cfg["clip_param"]).any() and (imp > 1 + cfg["clip_param"]).any()
        if cfg["use_huber_loss"]:                                 # the fixture reaches all three huber branches
            ret_n = (mb["returns"] - mb["ret_mean_orig"]) / torch.sqrt(mb["ret_var_orig"])
            e = ret_n - mb["values"]
            assert not torch.equal(mb["ret_mean"], mb["ret_mean_orig"])      # the two PopArt calls saw different moments
            assert (e > cfg["huber_delta"]).any() and (e < -cfg["huber_delta"]).any() and (e.abs() <= cfg["huber_delta"]).any()

"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF (build container only).

Run:  python tests/golden/make_golden.py [--reference /root/reference]

The reference is Python, so it cannot travel to the GPU box; instead its own classes
(`TenAnt`, `OneAnt`, `MultiIngenuity`, `MultiVecTaskPython`, `RolloutStorage`,
`SeparatedReplayBuffer`, `PopArt`) are imported here under oracle/refshim, driven with seeded
synthetic Isaac-layout frames, and their inputs + outputs are committed as small .npz fixtures.
While generating, every output is also compared bit for bit with oracle/ (the CPU restatement):
that comparison is what pins the oracle.  The script fails loudly on any mismatch.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from oracle import storage_oracle as so  # noqa: E402
from oracle.task_oracle import IngenuityOracle, OneAntOracle, TenAntOracle, multi_vec_task_step  # noqa: E402
from massive_marl_benchmark_b200 import synthetic  # noqa: E402


def _eq(name, a, b):
    if not torch.equal(a, b):
        raise SystemExit("oracle != reference for %s" % name)


def _np(d):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def _pad_rows(t, n, width=None):
    """Variable-length per-step outputs are stored padded with -1 to n rows (+ a count)."""
    out = torch.full((n,) + tuple(t.shape[1:]), -1, dtype=t.dtype)
    out[:t.shape[0]] = t
    return out


def gen_ten_ant(N=37, F=8, seed=101):
    task, g = refshim.make_task("TenAnt", N, True)
    orc = TenAntOracle(N)
    fr = synthetic.ten_ant_frames(N, F, seed=seed, fall_prob=0.01)
    npos, nvel = synthetic.reset_noise(N, F, seed=seed + 1)
    # exercise the episode-length reset: after step 0 (which resets every env because reset_buf starts as ones)
    # a few envs are moved close to max_episode_length; tests inject the same `progress_after0`
    prog_bump = {3: 996, 17: 993}
    rec = {k: [] for k in ("obs", "rew", "reset", "progress", "forces", "n_reset", "env_ids", "ant_box_indices",
                           "ant_indices", "dof_pushed", "obs_all", "reward_all", "done_all")}
    from agents.tasks.agent_base.multi_vec_task import MultiVecTaskPython
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        env = MultiVecTaskPython(task, "cpu")
    for t in range(F):
        g.push_frame(fr["root"][t], fr["dof"][t])
        a = fr["actions"][t] * 1.2          # some actions outside +-1 so the wrapper clamp matters
        a_list = [a[:, 8 * k:8 * k + 8] for k in range(10)]
        n_res = int(task.reset_buf.sum())
        # the reference draws its reset noise from the global generator: inject the fixture's rows by
        # seeding a generator state that reproduces them is impossible, so patch torch_rand_float instead
        import agents.tasks.ten_ant as ta
        noise_iter = iter([npos[t, :n_res], nvel[t, :n_res]])
        ta.torch_rand_float = lambda lo, hi, shape, device, _it=noise_iter: next(_it).clone()
        g.log.clear()
        if t == 0:
            # first call through the reference's own reset() path semantics: reset() == step(zeros)
            obs_all, state_all, rew_all, done_all, _, _ = env.step(a_list)
        else:
            obs_all, state_all, rew_all, done_all, _, _ = env.step(a_list)
        o_all, s_all, r_all, d_all = multi_vec_task_step(
            lambda act: orc.step(act, fr["root"][t], fr["dof"][t], noise=(npos[t], nvel[t])), a_list, 10, 38)
        _eq("ten_ant obs t%d" % t, task.obs_buf, orc.obs_buf)
        _eq("ten_ant rew", task.rew_buf, orc.rew_buf)
        _eq("ten_ant reset", task.reset_buf, orc.reset_buf)
        _eq("ten_ant progress", task.progress_buf, orc.progress_buf)
        _eq("ten_ant obs_all", obs_all, o_all)
        _eq("ten_ant state_all", state_all, s_all)
        _eq("ten_ant reward_all", rew_all, r_all)
        _eq("ten_ant done_all", done_all, d_all)
        logs = {e[0]: e for e in g.log}
        _eq("ten_ant forces", logs["dof_forces"][1], orc.last["forces"])
        env_ids = orc.last["env_ids"]
        assert len(env_ids) == n_res
        if n_res:
            _eq("ten_ant ant_box_indices", logs["root_indexed"][1], orc.last["ant_box_indices"])
            _eq("ten_ant ant_indices", logs["dof_indexed"][1], orc.last["ant_indices"])
            _eq("ten_ant dof_pushed", logs["dof_indexed"][3], orc.last["dof_pushed"])
            abi, ai, pushed = logs["root_indexed"][1], logs["dof_indexed"][1], logs["dof_indexed"][3]
        else:
            abi = torch.zeros(0, dtype=torch.int32)
            ai = torch.zeros(0, dtype=torch.int32)
            pushed = task.dof_state.clone()
        rec["obs"].append(task.obs_buf.clone()); rec["rew"].append(task.rew_buf.clone())
        rec["reset"].append(task.reset_buf.clone()); rec["progress"].append(task.progress_buf.clone())
        rec["forces"].append(logs["dof_forces"][1]); rec["n_reset"].append(torch.tensor(n_res))
        rec["env_ids"].append(_pad_rows(env_ids, N)); rec["ant_box_indices"].append(_pad_rows(abi, 11 * N))
        rec["ant_indices"].append(_pad_rows(ai, 10 * N)); rec["dof_pushed"].append(pushed)
        rec["obs_all"].append(obs_all.clone())
        rec["reward_all"].append(rew_all.clone()); rec["done_all"].append(done_all.clone())
        if t == 0:
            for e, v in prog_bump.items():
                task.progress_buf[e] = v
                orc.progress_buf[e] = v
            prog_after0 = task.progress_buf.clone()
    out = {k: torch.stack(v) for k, v in rec.items()}
    out.update(root=fr["root"], dof=fr["dof"], actions=fr["actions"] * 1.2, noise_pos=npos, noise_vel=nvel,
               progress_after0=prog_after0, initial_root=orc.initial_root_states)
    return out


def gen_one_ant(N=64, F=8, seed=202):
    task, g = refshim.make_task("OneAnt", N, False)
    orc = OneAntOracle(N)
    fr = synthetic.one_ant_frames(N, F, seed=seed, fall_prob=0.03)
    npos, nvel = synthetic.reset_noise(N, F, seed=seed + 1)
    prog_bump = {5: 997}
    import agents.tasks.one_ant as oa
    rec = {k: [] for k in ("obs", "obs_clamped", "rew", "reset", "progress", "forces", "n_reset", "env_ids", "ant_box_indices",
                           "ant_indices", "dof_pushed", "potentials", "prev_potentials", "up_vec", "heading_vec")}
    from agents.tasks.agent_base.vec_task import VecTaskPython
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        env = VecTaskPython(task, "cpu")
    for t in range(F):
        g.push_frame(fr["root"][t], fr["dof"][t], fr["sensor"][t])
        a = fr["actions"][t] * 1.2
        n_res = int(task.reset_buf.sum())
        noise_iter = iter([npos[t, :n_res], nvel[t, :n_res]])
        oa.torch_rand_float = lambda lo, hi, shape, device, _it=noise_iter: next(_it).clone()
        g.log.clear()
        obs_c, rew, done, _ = env.step(a)
        orc.step(torch.clamp(a, -1.0, 1.0), fr["root"][t], fr["dof"][t], fr["sensor"][t], noise=(npos[t], nvel[t]))
        for nm in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "potentials", "prev_potentials", "up_vec", "heading_vec"):
            _eq("one_ant " + nm, getattr(task, nm), getattr(orc, nm))
        logs = {e[0]: e for e in g.log}
        _eq("one_ant forces", logs["dof_forces"][1], orc.last["forces"])
        if n_res:
            _eq("one_ant abi", logs["root_indexed"][1], orc.last["ant_box_indices"])
            _eq("one_ant ai", logs["dof_indexed"][1], orc.last["ant_indices"])
            _eq("one_ant pushed", logs["dof_indexed"][3], orc.last["dof_pushed"])
            abi, ai, pushed = logs["root_indexed"][1], logs["dof_indexed"][1], logs["dof_indexed"][3]
        else:
            abi = torch.zeros(0, dtype=torch.int32); ai = torch.zeros(0, dtype=torch.int32); pushed = task.dof_state.clone()
        rec["obs"].append(task.obs_buf.clone()); rec["obs_clamped"].append(obs_c.clone()); rec["rew"].append(task.rew_buf.clone())
        rec["reset"].append(task.reset_buf.clone()); rec["progress"].append(task.progress_buf.clone())
        rec["forces"].append(logs["dof_forces"][1]); rec["n_reset"].append(torch.tensor(n_res))
        rec["env_ids"].append(_pad_rows(orc.last["env_ids"], N)); rec["ant_box_indices"].append(_pad_rows(abi, 2 * N))
        rec["ant_indices"].append(_pad_rows(ai, N)); rec["dof_pushed"].append(pushed)
        rec["potentials"].append(task.potentials.clone()); rec["prev_potentials"].append(task.prev_potentials.clone())
        rec["up_vec"].append(task.up_vec.clone()); rec["heading_vec"].append(task.heading_vec.clone())
        if t == 0:
            for e, v in prog_bump.items():
                task.progress_buf[e] = v
                orc.progress_buf[e] = v
            prog_after0 = task.progress_buf.clone()
    out = {k: torch.stack(v) for k, v in rec.items()}
    out.update(root=fr["root"], dof=fr["dof"], sensor=fr["sensor"], actions=fr["actions"] * 1.2, noise_pos=npos,
               noise_vel=nvel, progress_after0=prog_after0, initial_root=orc.initial_root_states)
    return out


def gen_ingenuity(N=33, F=8, seed=303):
    task, g = refshim.make_task("MultiIngenuity", N, False)
    orc = IngenuityOracle(N)
    fr = synthetic.ingenuity_frames(N, F, seed=seed)
    prog_bump = {2: 998}
    rec = {k: [] for k in ("obs", "rew", "reset", "progress", "forces", "n_reset", "env_ids", "actor_indices", "dof_pushed")}
    for t in range(F):
        g.push_frame(fr["root"][t])
        a = fr["actions"][t]
        n_res = int(task.reset_buf.sum())
        g.log.clear()
        task.step(a)
        orc.step(a, fr["root"][t])
        for nm in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "forces"):
            _eq("ingenuity " + nm, getattr(task, nm), getattr(orc, nm))
        logs = {e[0]: e for e in g.log}
        _eq("ingenuity body forces", logs["body_forces"][1], orc.last["forces"])
        if n_res:
            _eq("ingenuity idx", logs["root_indexed"][1], orc.last["actor_indices"])
            _eq("ingenuity pushed", logs["dof_indexed"][3], orc.last["dof_pushed"])
            ai, pushed = logs["root_indexed"][1], logs["dof_indexed"][3]
        else:
            ai = torch.zeros(0, dtype=torch.int32); pushed = task.dof_state.clone()
        rec["obs"].append(task.obs_buf.clone()); rec["rew"].append(task.rew_buf.clone())
        rec["reset"].append(task.reset_buf.clone()); rec["progress"].append(task.progress_buf.clone())
        rec["forces"].append(logs["body_forces"][1]); rec["n_reset"].append(torch.tensor(n_res))
        rec["env_ids"].append(_pad_rows(orc.last["env_ids"], N)); rec["actor_indices"].append(_pad_rows(ai, 4 * N))
        rec["dof_pushed"].append(pushed)
        if t == 0:
            for e, v in prog_bump.items():
                task.progress_buf[e] = v
                orc.progress_buf[e] = v
            prog_after0 = task.progress_buf.clone()
    out = {k: torch.stack(v) for k, v in rec.items()}
    out.update(root=fr["root"], actions=fr["actions"], progress_after0=prog_after0, initial_root=orc.initial_root_states)
    return out


def gen_storage_ppo(T=8, N=50, seed=404):
    from agents.algorithms.rl.ppo.storage import RolloutStorage
    gen = torch.Generator().manual_seed(seed)
    st = RolloutStorage(N, T, (60,), (0,), (8,), "cpu", "random")
    ins = []
    for t in range(T):
        obs = torch.randn(N, 60, generator=gen); states = torch.zeros(N, 0)
        act = torch.randn(N, 8, generator=gen); rew = torch.randn(N, generator=gen)
        done = (torch.rand(N, generator=gen) < 0.1).long()
        val = torch.randn(N, 1, generator=gen); logp = torch.randn(N, generator=gen)
        mu = torch.randn(N, 8, generator=gen); sig = torch.randn(N, 8, generator=gen)
        st.add_transitions(obs, states, act, rew, done, val, logp, mu, sig)
        ins.append((obs, act, rew, done, val, logp, mu, sig))
    try:
        st.add_transitions(*([ins[0][0], torch.zeros(N, 0)] + list(ins[0][1:])))
        raise SystemExit("overflow not raised")
    except AssertionError as e:
        overflow_msg = str(e)
    last_values = torch.randn(N, 1, generator=gen)
    # Reference quirk: get_statistics does `done = self.dones.cpu(); done[-1] = 1` (storage.py:68-69).  On the
    # reference's production device (CUDA) .cpu() copies; on CPU it ALIASES, so the CPU run would overwrite the
    # last row of its own dones before compute_returns.  The fixture keeps the device semantics: dones restored.
    dones_before = st.dones.clone()
    mean_len, mean_rew = st.get_statistics()
    st.dones.copy_(dones_before)
    st.compute_returns(last_values, 0.96, 0.95)
    ret, adv = so.ppo_compute_returns(st.rewards, st.values, st.dones, last_values, 0.96, 0.95)
    _eq("ppo returns", st.returns, ret)
    _eq("ppo advantages", st.advantages, adv)
    ml, mr = so.ppo_get_statistics(st.dones, st.rewards)
    _eq("ppo stats", torch.stack([mean_len, mean_rew]), torch.stack([ml, mr]))
    # sequential partition
    st.sampler = "sequential"
    part = [list(b) for b in st.mini_batch_generator(4)]
    assert part == so.ppo_minibatch_partition(N, T, 4)
    part3 = [list(b) for b in st.mini_batch_generator(3)]     # 400 // 3 = 133, drop_last drops one index
    assert part3 == so.ppo_minibatch_partition(N, T, 3)
    names = ("obs", "act", "rew", "done", "val", "logp", "mu", "sig")
    out = {"in_" + n: torch.stack([x[i] for x in ins]) for i, n in enumerate(names)}
    out.update(last_values=last_values, returns=st.returns, advantages=st.advantages,
               dones_u8=st.dones, mean_len=mean_len, mean_rew=mean_rew,
               part4_sizes=torch.tensor([len(p) for p in part]), part3_sizes=torch.tensor([len(p) for p in part3]),
               part3_last=torch.tensor(part3[-1][-1]))
    out["overflow_msg"] = np.array(overflow_msg)
    return out


def gen_buffer_marl(T=8, N=24, seed=505):
    from agents.algorithms.marl.utils.separated_buffer import SeparatedReplayBuffer
    from agents.algorithms.marl.utils.popart import PopArt
    from gym import spaces
    gen = torch.Generator().manual_seed(seed)
    cfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=16, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
               use_gae=True, use_popart=True, use_valuenorm=False, use_proper_time_limits=False)
    obs_space = spaces.Box(low=-np.inf, high=np.inf, shape=(46,))
    sh_space = spaces.Box(low=-np.inf, high=np.inf, shape=(388,))
    act_space = spaces.Box(low=-np.ones(8), high=np.ones(8))
    buf = SeparatedReplayBuffer(cfg, obs_space, sh_space, act_space, "cpu")
    pa = PopArt(1)
    pa(torch.randn(256, 1, generator=gen) * 1.7 + 0.4)          # one running-moment update
    pa(torch.randn(256, 1, generator=gen) * 1.1 - 0.2)
    ins = []
    for t in range(T):
        sh = torch.randn(N, 388, generator=gen); ob = torch.randn(N, 46, generator=gen)
        rs = torch.zeros(N, 1, 16); act = torch.randn(N, 8, generator=gen); lp = torch.randn(N, 8, generator=gen)
        vp = torch.randn(N, 1, generator=gen); rw = torch.randn(N, 1, generator=gen)
        mk = (torch.rand(N, 1, generator=gen) > 0.1).float(); am = (torch.rand(N, 1, generator=gen) > 0.05).float()
        buf.insert(sh, ob, rs, rs, act, lp, vp, rw, mk, None, am, None)
        ins.append((sh, ob, act, lp, vp, rw, mk, am))
    next_value = torch.randn(N, 1, generator=gen)
    buf.compute_returns(next_value, pa)
    mean, var = so.popart_running_mean_var(pa.running_mean, pa.running_mean_sq, pa.debiasing_term)
    m2, v2 = pa.running_mean_var()
    _eq("popart mean", mean, m2); _eq("popart var", var, v2)
    ret, vps = so.marl_compute_returns(buf.rewards, buf.value_preds, buf.masks, buf.bad_masks, next_value,
                                       0.96, 0.95, denorm=(mean, var))
    _eq("marl returns", buf.returns, ret)
    adv_ref = buf.returns[:-1] - pa.denormalize(buf.value_preds[:-1])
    adv_ref = (adv_ref - torch.mean(adv_ref.clone())) / (torch.std(adv_ref.clone()) + 1e-5)
    adv = so.marl_advantages(buf.returns, buf.value_preds, denorm=(mean, var))
    _eq("marl advantages", adv_ref, adv)
    # no-normaliser branch and proper-time-limits branch
    outs = {}
    for tag, kw in (("plain", dict(use_popart=False, use_valuenorm=False)),
                    ("ptl", dict(use_proper_time_limits=True))):
        c2 = dict(cfg, **kw)
        b2 = SeparatedReplayBuffer(c2, obs_space, sh_space, act_space, "cpu")
        for x in ("rewards", "value_preds", "masks"):
            getattr(b2, x).copy_(getattr(buf, x))
        b2.bad_masks.copy_((torch.rand(T + 1, N, 1, generator=gen) > 0.1).float())
        vn = pa if (c2["use_popart"] or c2["use_valuenorm"]) else None
        b2.compute_returns(next_value, vn)
        r2, _ = so.marl_compute_returns(b2.rewards, buf.value_preds, b2.masks, b2.bad_masks, next_value, 0.96, 0.95,
                                        denorm=(mean, var) if vn is not None else None,
                                        use_proper_time_limits=c2["use_proper_time_limits"])
        _eq("marl returns " + tag, b2.returns, r2)
        outs["returns_" + tag] = b2.returns.clone()
        outs["bad_masks_" + tag] = b2.bad_masks.clone()
    # Runner.insert mask logic (runner.py:229-255) restated inline from the reference semantics
    dones = (torch.rand(N, 10, generator=gen) < 0.2).long()
    dones[3] = 1
    masks, active = so.runner_insert_masks(dones)
    names = ("share_obs", "obs", "actions", "logp", "value_preds", "rewards", "masks", "active_masks")
    out = {"in_" + n: torch.stack([x[i] for x in ins]) for i, n in enumerate(names)}
    out.update(next_value=next_value, popart_mean=mean, popart_var=var, returns=buf.returns, value_preds_after=buf.value_preds,
               advantages=adv_ref, buf_masks=buf.masks, buf_active_masks=buf.active_masks, buf_share_obs=buf.share_obs,
               buf_obs=buf.obs, runner_dones=dones, runner_masks=masks, runner_active_masks=active, **outs)
    return out


def marl_mlp_functional(sd, x, head):
    """fp32 restatement of MLPBase + head (agents/algorithms/utils/mlp.py:31-65, actor_critic.py:60-69,165-168)."""
    import torch.nn.functional as F
    h = F.layer_norm(x, (x.shape[1],), sd["base.feature_norm.weight"], sd["base.feature_norm.bias"])
    h = F.layer_norm(F.elu(F.linear(h, sd["base.mlp.fc1.0.weight"], sd["base.mlp.fc1.0.bias"])), (512,),
                     sd["base.mlp.fc1.2.weight"], sd["base.mlp.fc1.2.bias"])
    for i in range(2):
        h = F.layer_norm(F.elu(F.linear(h, sd["base.mlp.fc2.%d.0.weight" % i], sd["base.mlp.fc2.%d.0.bias" % i])), (512,),
                         sd["base.mlp.fc2.%d.2.weight" % i], sd["base.mlp.fc2.%d.2.bias" % i])
    return F.linear(h, sd[head + ".weight"], sd[head + ".bias"])


def gen_mlp_marl(reference_root, M=200, seed=606):
    """The shipped TenAnt MAPPO checkpoint of agent 0's ACTOR (logs/ten_ant/mappo/models_seed-1/actor_agent0.pt: real
    weights with real dynamic range) + inputs + the fp32 output of the reference's own MLPBase module."""
    from agents.algorithms.utils.mlp import MLPBase
    sd = torch.load(os.path.join(reference_root, "logs/ten_ant/mappo/models_seed-1/actor_agent0.pt"), map_location="cpu",
                    weights_only=False)
    cfg = dict(use_feature_normalization=True, use_orthogonal=True, use_ReLU=True, stacked_frames=1, layer_N=2, hidden_size=512)
    base = MLPBase(cfg, (46,))
    base.load_state_dict({k[len("base."):]: v for k, v in sd.items() if k.startswith("base.")})
    gen = torch.Generator().manual_seed(seed)
    x = torch.clamp(torch.randn(M, 46, generator=gen) * 2.0, -7, 7)
    with torch.no_grad():
        feat = base(x)
        mean = torch.nn.functional.linear(feat, sd["act.action_out.fc_mean.weight"], sd["act.action_out.fc_mean.bias"])
        mine = marl_mlp_functional(sd, x, "act.action_out.fc_mean")
    _eq("marl mlp functional vs MLPBase", mean, mine)
    out = {"w_" + k.replace(".", "__"): v for k, v in sd.items()}
    out.update(x=x, features=feat, mean=mean)
    return out


def gen_ppo_loss(reference_root):
    """Three minibatches through the reference's own `ActorCritic.evaluate` + the loss lines of `PPO.update`
    (oracle/ref_ppo_loss.py executes them): OneAnt width (8), TenAnt width (80), and the unclipped value loss."""
    from oracle.ppo_loss_oracle import ppo_loss_oracle, synthetic_minibatch
    from oracle.ref_ppo_loss import reference_ppo_loss
    out = {}
    cases = [("a8", 96, 8, 707, dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, use_clipped_value_loss=True), 0.15),
             ("a80", 70, 80, 808, dict(clip_param=0.2, value_loss_coef=2.0, entropy_coef=0.01, use_clipped_value_loss=True), 0.05),
             ("a8u", 64, 8, 909, dict(clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.0, use_clipped_value_loss=False), 0.15)]
    for tag, B, A, seed, cfg, spread in cases:
        mb = synthetic_minibatch(B, A, seed, ratio_spread=spread)
        ref = reference_ppo_loss(reference_root, mb, **cfg)
        mine = ppo_loss_oracle(**mb, **cfg)
        for k in ref:
            _eq("ppo loss %s %s" % (tag, k), ref[k], mine[k])
        out.update({"%s__in_%s" % (tag, k): v for k, v in mb.items()})
        out.update({"%s__out_%s" % (tag, k): v for k, v in ref.items()})
        out.update({"%s__cfg_%s" % (tag, k): np.float64(v) for k, v in cfg.items()})
    return out


MAPPO_LOSS_BASE = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.0, huber_delta=10.0, use_huber_loss=True,
                       use_clipped_value_loss=True, use_value_active_masks=False, use_policy_active_masks=False,
                       std_x_coef=1.0, std_y_coef=0.5)                       # cfg/mappo/config.yaml


def gen_mappo_loss():
    """Minibatches through the reference's own `MAPPO.ppo_update` (+ its ACTLayer / DiagGaussian / PopArt,
    oracle/ref_mappo_loss.py): the shipped configuration, the masked + small-delta variant, the mse / unclipped one."""
    from oracle.mappo_loss_oracle import mappo_loss_oracle, synthetic_minibatch
    from oracle.ref_mappo_loss import reference_mappo_loss
    out = {}
    cases = [("cfg", 96, 8, 1001, {}),
             ("mask", 80, 8, 1002, dict(use_value_active_masks=True, use_policy_active_masks=True, entropy_coef=0.01, huber_delta=1.0)),
             ("mse", 64, 6, 1003, dict(use_huber_loss=False, use_clipped_value_loss=False, value_loss_coef=0.5, clip_param=0.1))]
    for tag, B, A, seed, over in cases:
        cfg = dict(MAPPO_LOSS_BASE, **over)
        mb = synthetic_minibatch(B, A, seed, huber_delta=cfg["huber_delta"])
        ref, mb = reference_mappo_loss(mb, cfg)
        mine = mappo_loss_oracle(**mb, **cfg)
        for k in ref:
            _eq("mappo loss %s %s" % (tag, k), ref[k], mine[k])
        out.update({"%s__in_%s" % (tag, k): v for k, v in mb.items()})
        out.update({"%s__out_%s" % (tag, k): v for k, v in mine.items()})       # (= ref, plus the per-dimension log-probs)
        out.update({"%s__cfg_%s" % (tag, k): np.float64(v) for k, v in cfg.items()})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default=None, help="regenerate one fixture (by name)")
    args = ap.parse_args()
    refshim.install(args.reference)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    for name, fn in (("ten_ant_n37", gen_ten_ant), ("one_ant_n64", gen_one_ant), ("ingenuity_n33", gen_ingenuity),
                     ("storage_ppo", gen_storage_ppo), ("buffer_marl", gen_buffer_marl),
                     ("mlp_marl_actor0", lambda: gen_mlp_marl(args.reference)),
                     ("ppo_loss", lambda: gen_ppo_loss(args.reference)), ("mappo_loss", gen_mappo_loss)):
        if args.only and name != args.only:
            continue
        data = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **_np(data))
        print("%-16s %7.1f KB  keys=%d" % (name, os.path.getsize(path) / 1024, len(data)))


if __name__ == "__main__":
    main()

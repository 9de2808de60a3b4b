"""Drop-in proof: the reference's own training loops drive the replacement classes.

1. The exact loop of the reference's `PPO.run` (ppo.py:127-139: `current_obs = reset()`, `step`, `add_transitions(current_obs,
   ...)`, `current_obs.copy_(next_obs)`), restated inline, stores the observations the oracle predicts - a returned tensor is
   never rewritten by a later step (the reference returns a fresh `torch.clamp` result every step).
2. With the unmodified reference installed in baseline/_ref (baseline/make_ref.py; skipped when absent), the reference's
   UNMODIFIED `PPO.run` trains OneAnt N = 64 for 3 iterations (BASELINE configs[0]) on `tasks.OneAnt` + `VecTaskPython` +
   `storage.RolloutStorage` with only the storage import swapped; the first rollout it stored is
   then compared with the reference's own OneAnt / VecTaskPython / RolloutStorage fed the same frames and the recorded actions.
3. Likewise the reference's unmodified MARL `Runner.run` (MAPPO, TenAnt, 2 episodes) on `tasks.TenAnt` + `MultiVecTaskPython` +
   `separated_buffer.SeparatedReplayBuffer`.
The reference-driven runs execute in a subprocess: `oracle/refshim` must register its `gym` stand-in before the product
package is imported, so that the product's spaces are instances of the class the reference type-checks (ppo.py:34-39).
"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, assert_close_obs

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "agents")), reason="baseline/_ref not installed (python baseline/make_ref.py)")


def test_reference_ppo_loop_stores_the_right_observations(cuda_device):
    from oracle.task_oracle import TenAntOracle
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython
    dev = cuda_device
    N, T = 96, 6
    fr = synthetic.ten_ant_frames(N, T + 1, seed=17, fall_prob=0.02)
    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
    task = TenAnt(cfg, None, None, "cuda", 0, True, False,
                  provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev, loop=False), flavor=L.FLAVOR_CUDA)
    env = VecTaskPython(task, dev)
    st = RolloutStorage(N, T, (388,), (0,), (80,), dev)
    gen = torch.Generator().manual_seed(3)
    values = torch.randn(T, N, 1, generator=gen).to(dev)
    states = torch.zeros(N, 0, device=dev)
    torch.manual_seed(11)
    a_reset = 0.01 * (1 - 2 * torch.rand([N, 80], dtype=torch.float32, device=dev))   # what reset() will draw (vec_task.py:134)
    torch.manual_seed(11)
    # --- the reference's loop, verbatim in structure (ppo.py:128-139) ---
    current_obs = env.reset()
    for t in range(T):
        actions = fr["actions"][t + 1].to(dev)
        next_obs, rews, dones, infos = env.step(actions)
        st.add_transitions(current_obs, states, actions, rews, dones, values[t], rews, actions, actions)
        current_obs.copy_(next_obs)
    torch.cuda.synchronize()
    # --- oracle: obs stored in slot t is the observation BEFORE step t ---
    orc = TenAntOracle(N, device="cuda")
    obs, _, _ = orc.step(a_reset, fr["root"][0].to(dev), fr["dof"][0].to(dev))
    for t in range(T):
        assert_close_obs(st.observations[t], torch.clamp(obs, -5, 5), angle_cols=tuple(38 * k + c for k in range(10) for c in (9, 10, 11)),
                         what="stored obs slot %d" % t)
        a = torch.clamp(fr["actions"][t + 1].to(dev), -1, 1)
        obs, rew, done = orc.step(a, fr["root"][t + 1].to(dev), fr["dof"][t + 1].to(dev))
        assert torch.equal(st.dones[t, :, 0].long(), done)
        rel = (st.rewards[t, :, 0] - rew).abs() / rew.abs().clamp(min=1e-6)
        assert float(rel.max()) <= 1e-5
    assert_close_obs(current_obs, torch.clamp(obs, -5, 5), angle_cols=tuple(38 * k + c for k in range(10) for c in (9, 10, 11)), what="current_obs")


_PPO_SCRIPT = r'''
import contextlib, io, os, sys, torch, yaml
ROOT, REF, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
from oracle import refshim
refshim.install(REF)                                   # registers the gym stand-in BEFORE the product package is imported
from massive_marl_benchmark_b200 import _lib as L, synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.storage import RolloutStorage
from massive_marl_benchmark_b200.tasks import OneAnt
from massive_marl_benchmark_b200.vec_task import VecTaskPython
import importlib
ns = sys.modules["agents.algorithms.rl.ppo"]            # refshim registers the package as a bare namespace (its __init__ pulls PPO in)
ns.ActorCritic = importlib.import_module("agents.algorithms.rl.ppo.module").ActorCritic      # ppo/__init__.py:2, unchanged
ns.RolloutStorage = RolloutStorage                      # ppo/__init__.py:1: the ONE import swapped (INTEGRATION.md section 1)
import agents.algorithms.rl.ppo.ppo as refppo           # the reference's own, unmodified PPO
assert refppo.RolloutStorage is RolloutStorage

dev = torch.device("cuda", 0)
N, ITERS = 64, 3
cfg_train = yaml.safe_load(open(os.path.join(REF, "cfg", "ppo", "config.yaml")))
T = cfg_train["learn"]["nsteps"]
F = 1 + ITERS * T
fr = synthetic.one_ant_frames(N, F, seed=99, fall_prob=0.02)
cfg = {"env": {"numEnvs": N, "env_name": "one_ant"}, "sim": {"dt": 0.0166}, "seed": 5}
task = OneAnt(cfg, None, None, "cuda", 0, True, False, provider=ReplayProvider(fr, device=dev, loop=False), flavor=L.FLAVOR_CPU)
env = VecTaskPython(task, "cuda:0")
rec = {"actions": [], "first": None}
task_step = task.step
def recording_step(a):
    rec["actions"].append(a.detach().clone().cpu())
    return task_step(a)
task.step = recording_step
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ppo = refppo.PPO(vec_env=env, cfg_train=cfg_train, device="cuda:0", sampler="sequential", log_dir=os.path.join(OUT, "run"),
                     is_testing=False, print_log=True, apply_reset=False, asymmetric=False)
assert type(ppo.storage) is RolloutStorage
compute_returns = ppo.storage.compute_returns
def snapshot_returns(last_values, gamma, lam):
    compute_returns(last_values, gamma, lam)
    if rec["first"] is None:
        s = ppo.storage
        rec["first"] = {k: getattr(s, k).detach().clone().cpu() for k in
                        ("observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma", "returns", "advantages")}
        rec["first"]["last_values"] = last_values.detach().clone().cpu()
ppo.storage.compute_returns = snapshot_returns
with contextlib.redirect_stdout(io.StringIO()):
    ppo.run(num_learning_iterations=ITERS, log_interval=1)
torch.cuda.synchronize()
assert all(os.path.exists(os.path.join(OUT, "run", "model_%d.pt" % i)) for i in range(ITERS + 1))
assert len(rec["actions"]) == F and ppo.tot_timesteps == ITERS * T * N
for p in ppo.actor_critic.parameters():
    assert torch.isfinite(p).all()

# ---- the same first rollout through the reference's OWN classes (CPU, FakeGym frames), fed the recorded actions ----
ref_task, gym = refshim.make_task("OneAnt", N, False)
from agents.tasks.agent_base.vec_task import VecTaskPython as RefVecTask
from agents.algorithms.rl.ppo.storage import RolloutStorage as RefStorage
with contextlib.redirect_stdout(io.StringIO()):
    ref_env = RefVecTask(ref_task, "cpu")
for t in range(T + 1):
    gym.push_frame(fr["root"][t], fr["dof"][t], fr["sensor"][t])
ref_st = RefStorage(N, T, (60,), (0,), (8,), "cpu", "sequential")
first = rec["first"]
ref_task.step(rec["actions"][0])                                       # reset(): vec_task.py:133-139 with the recorded draw
cur = torch.clamp(ref_task.obs_buf, -5.0, 5.0).clone()
states = torch.zeros(N, 0)
for t in range(T):
    obs, rew, done, _ = ref_env.step(rec["actions"][1 + t])
    ref_st.add_transitions(cur, states, rec["actions"][1 + t], rew, done, first["values"][t], first["actions_log_prob"][t].view(-1),
                           first["mu"][t], first["sigma"][t])
    cur.copy_(obs)
ref_st.compute_returns(first["last_values"], cfg_train["learn"]["gamma"], cfg_train["learn"]["lam"])

def close(a, b, rtol, atol, what, angle_cols=()):
    import math
    d = (a.double() - b.double()).abs()
    for c in angle_cols:
        d[..., c] = torch.minimum(d[..., c], (d[..., c] - 2 * math.pi).abs())
    bad = d > rtol * b.double().abs() + atol
    assert not bad.any(), "%s: %d / %d out of tolerance (max %g)" % (what, int(bad.sum()), bad.numel(), float(d.max()))

assert torch.equal(first["dones"], ref_st.dones), "dones"
assert torch.equal(first["actions"], ref_st.actions)
close(first["observations"], ref_st.observations, 1e-5, 1e-6, "observations", angle_cols=(7, 8, 9))
close(first["rewards"], ref_st.rewards, 1e-5, 1e-5, "rewards")
close(first["returns"], ref_st.returns, 1e-5, 1e-4, "returns")
close(first["advantages"], ref_st.advantages, 1e-4, 1e-4, "advantages")
print("DROPIN_PPO_OK iterations=%d timesteps=%d" % (ITERS, ppo.tot_timesteps))
'''


@needs_ref
def test_reference_ppo_run_drives_the_dropins(cuda_device, tmp_path):
    res = subprocess.run([sys.executable, "-c", _PPO_SCRIPT, ROOT, REF, str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "DROPIN_PPO_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


_MARL_SCRIPT = r'''
import contextlib, io, os, sys, torch, yaml
ROOT, REF, OUT = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
from oracle import refshim
refshim.install(REF)
from massive_marl_benchmark_b200 import _lib as L, synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.separated_buffer import SeparatedReplayBuffer
from massive_marl_benchmark_b200.tasks import TenAnt
from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython
import agents.algorithms.marl.runner as refrunner       # the reference's own, unmodified Runner
refrunner.SeparatedReplayBuffer = SeparatedReplayBuffer  # INTEGRATION.md section 1: the one import swapped in runner.py:17

dev = torch.device("cuda", 0)
N, EPISODES = 32, 2
config = yaml.safe_load(open(os.path.join(REF, "cfg", "mappo", "config.yaml")))
T = config["episode_length"]
config.update(n_rollout_threads=N, n_eval_rollout_threads=N, num_env_steps=EPISODES * T * N, run_dir=os.path.join(OUT, "run"),
              experiment_name="dropin", save_interval=1, log_interval=1, use_eval=False, ppo_epoch=2)
F = 1 + EPISODES * T
fr = synthetic.ten_ant_frames(N, F, seed=123, fall_prob=0.01)
cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
task = TenAnt(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev, loop=False),
              flavor=L.FLAVOR_CPU)
env = MultiVecTaskPython(task, "cuda:0")
rec = {"actions": [], "first": None}
task_step = task.step
def recording_step(a):
    rec["actions"].append(a.detach().clone().cpu())
    return task_step(a)
task.step = recording_step
task_step_agents = task.step_agent_actions
def recording_step_agents(al):                          # the multi-agent wrapper hands the per-agent tensors over unstacked
    rec["actions"].append(torch.cat([a.detach() for a in al], dim=1).clone().cpu())
    return task_step_agents(al)
task.step_agent_actions = recording_step_agents
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    runner = refrunner.Runner(vec_env=env, config=config, model_dir="")
assert all(type(b) is SeparatedReplayBuffer for b in runner.buffer)
compute = runner.compute
def snapshot_compute():
    compute()
    if rec["first"] is None:
        rec["first"] = [{k: getattr(b, k).detach().clone().cpu() for k in ("share_obs", "obs", "rewards", "masks", "active_masks", "value_preds", "returns", "actions")}
                        for b in runner.buffer]
runner.compute = snapshot_compute
with contextlib.redirect_stdout(io.StringIO()):
    runner.run()
torch.cuda.synchronize()
assert len(rec["actions"]) == F, (len(rec["actions"]), F)
save_dir = os.path.join(OUT, "run", "ten_ant", "mappo", "models_seed1")
assert all(os.path.exists(os.path.join(save_dir, "actor_agent%d.pt" % a)) for a in range(10))
for tr in runner.trainer:
    for p in tr.policy.actor.parameters():
        assert torch.isfinite(p).all()

# ---- the same first episode through the reference's OWN TenAnt + MultiVecTaskPython + SeparatedReplayBuffer (CPU) ----
ref_task, gym = refshim.make_task("TenAnt", N, True)
from agents.tasks.agent_base.multi_vec_task import MultiVecTaskPython as RefMulti
from agents.algorithms.marl.utils.separated_buffer import SeparatedReplayBuffer as RefBuffer
with contextlib.redirect_stdout(io.StringIO()):
    ref_env = RefMulti(ref_task, "cpu")
for t in range(T + 1):
    gym.push_frame(fr["root"][t], fr["dof"][t])
obs, share, _ = ref_env.reset()                        # zero actions (multi_vec_task.py:146-175): same as the recorded ones
assert float(rec["actions"][0].abs().sum()) == 0.0
first = rec["first"]

def close(a, b, rtol, atol, what, angle_w=None):
    import math
    d = (a.double() - b.double()).abs()
    if angle_w:
        for k in range(a.shape[-1] // angle_w if angle_w == 38 else 1):
            for c in (9, 10, 11):
                col = k * angle_w + c
                d[..., col] = torch.minimum(d[..., col], (d[..., col] - 2 * math.pi).abs())
    bad = d > rtol * b.double().abs() + atol
    assert not bad.any(), "%s: %d / %d out of tolerance (max %g)" % (what, int(bad.sum()), bad.numel(), float(d.max()))

for a in range(10):
    close(first[a]["obs"][0], obs[:, a], 1e-5, 1e-6, "obs[0] agent %d" % a, angle_w=46)
    close(first[a]["share_obs"][0], share[:, a], 1e-5, 1e-6, "share_obs[0] agent %d" % a, angle_w=38)
for t in range(T):
    acts = rec["actions"][1 + t]
    obs, share, rew, done, info, _ = ref_env.step([acts[:, 8 * a:8 * a + 8] for a in range(10)])
    dones_env = torch.all(done, dim=1)
    for a in range(10):
        close(first[a]["obs"][t + 1], obs[:, a], 1e-5, 1e-6, "obs[%d] agent %d" % (t + 1, a), angle_w=46)
        close(first[a]["share_obs"][t + 1], share[:, a], 1e-5, 1e-6, "share_obs[%d] agent %d" % (t + 1, a), angle_w=38)
        close(first[a]["rewards"][t], rew[:, a], 1e-5, 1e-5, "rewards[%d] agent %d" % (t, a))
        want_mask = torch.ones(N, 1); want_mask[dones_env] = 0.0                 # runner.py:232-236
        assert torch.equal(first[a]["masks"][t + 1], want_mask), "masks[%d] agent %d" % (t + 1, a)
        assert torch.equal(first[a]["actions"][t], acts[:, 8 * a:8 * a + 8])
print("DROPIN_MARL_OK episodes=%d" % EPISODES)
'''


@needs_ref
def test_reference_marl_runner_drives_the_dropins(cuda_device, tmp_path):
    res = subprocess.run([sys.executable, "-c", _MARL_SCRIPT, ROOT, REF, str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "DROPIN_MARL_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


_IPPO_SCRIPT = r'''
import contextlib, io, os, sys, torch, yaml
ROOT, REF, OUT, ALGO, FUSED = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] in ("1", "2")
TEAM = sys.argv[5] == "2"          # grouped tensor-core forward for collect() + one shared buffer for the team
sys.path.insert(0, ROOT)
from oracle import refshim
refshim.install(REF)
from massive_marl_benchmark_b200 import _lib as L, synthetic
from massive_marl_benchmark_b200.providers import ReplayProvider
from massive_marl_benchmark_b200.runner import Runner, process_MultiAgentRL, resolve_algorithm
from massive_marl_benchmark_b200.tasks import TenAnt
from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython

dev = torch.device("cuda", 0)
N, EPISODES = 32, 2
config = yaml.safe_load(open(os.path.join(REF, "cfg", ALGO, "config.yaml")))
T = config["episode_length"]
config.update(n_rollout_threads=N, n_eval_rollout_threads=N, num_env_steps=EPISODES * T * N, run_dir=os.path.join(OUT, "run"),
              experiment_name="dropin", save_interval=1, log_interval=1, use_eval=False, ppo_epoch=2)
F = 1 + EPISODES * T
fr = synthetic.ten_ant_frames(N, F, seed=321, fall_prob=0.05)
cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}
task = TenAnt(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev, loop=False),
              flavor=L.FLAVOR_CPU)
env = MultiVecTaskPython(task, "cuda:0")
rec = {"actions": [], "first": None}
task_step = task.step
def recording_step(a):
    rec["actions"].append(a.detach().clone().cpu())
    return task_step(a)
task.step = recording_step
task_step_agents = task.step_agent_actions
def recording_step_agents(al):
    rec["actions"].append(torch.cat([a.detach() for a in al], dim=1).clone().cpu())
    return task_step_agents(al)
task.step_agent_actions = recording_step_agents
torch.manual_seed(0)
class Args: algo = ALGO
with contextlib.redirect_stdout(io.StringIO()):
    runner = Runner(vec_env=env, config=config, model_dir="", fused_update=FUSED, team_forward=TEAM,
                    shared_buffer=TEAM and config["use_centralized_V"])
TrainAlgo, Policy = resolve_algorithm(ALGO)
assert all(isinstance(t, TrainAlgo) for t in runner.trainer) and all(isinstance(p, Policy) for p in runner.policy)
compute = runner.compute
def snapshot_compute():
    compute()
    if rec["first"] is None:
        rec["first"] = [{k: getattr(b, k).detach().clone().cpu() for k in ("share_obs", "obs", "rewards", "masks", "active_masks", "actions", "returns", "value_preds")}
                        for b in runner.buffer]
runner.compute = snapshot_compute
before = [p.actor.base.mlp.fc1[0].weight.detach().clone() for p in runner.policy]
with contextlib.redirect_stdout(io.StringIO()) as log:
    runner.run()
torch.cuda.synchronize()
assert len(rec["actions"]) == F
save_dir = os.path.join(OUT, "run", "ten_ant", ALGO, "models_seed1")
assert all(os.path.exists(os.path.join(save_dir, "%s_agent%d.pt" % (k, a))) for a in range(10) for k in ("actor", "critic"))
for p, w0 in zip(runner.policy, before):
    w1 = p.actor.base.mlp.fc1[0].weight
    assert torch.isfinite(w1).all() and not torch.equal(w1, w0), "the update must have moved every agent's actor"

# env side of the first episode against the reference's OWN TenAnt + MultiVecTaskPython on the recorded actions
ref_task, gym = refshim.make_task("TenAnt", N, True)
from agents.tasks.agent_base.multi_vec_task import MultiVecTaskPython as RefMulti
with contextlib.redirect_stdout(io.StringIO()):
    ref_env = RefMulti(ref_task, "cpu")
for t in range(T + 1):
    gym.push_frame(fr["root"][t], fr["dof"][t])
obs, share, _ = ref_env.reset()
first = rec["first"]
centralized = config["use_centralized_V"]

def close(a, b, what, w):
    import math
    d = (a.double() - b.double()).abs()
    for k in range(a.shape[-1] // w if w == 38 else 1):
        for c in (9, 10, 11):
            d[..., k * w + c] = torch.minimum(d[..., k * w + c], (d[..., k * w + c] - 2 * math.pi).abs())
    bad = d > 1e-5 * b.double().abs() + 1e-6
    assert not bad.any(), "%s: %d / %d out of tolerance (max %g)" % (what, int(bad.sum()), bad.numel(), float(d.max()))

n_done_env = 0
for t in range(T + 1):
    if t > 0:
        acts = rec["actions"][t]
        obs, share, rew, done, info, _ = ref_env.step([acts[:, 8 * a:8 * a + 8] for a in range(10)])
        dones_env = torch.all(done, dim=1)
        n_done_env += int(dones_env.sum())
    for a in range(10):
        close(first[a]["obs"][t], obs[:, a], "obs[%d] agent %d" % (t, a), 46)
        if centralized:
            close(first[a]["share_obs"][t], share[:, a], "share_obs[%d] agent %d" % (t, a), 38)
        else:                                            # IPPO: the critic sees the agent's own observation (runner.py:190-191)
            assert torch.equal(first[a]["share_obs"][t], first[a]["obs"][t])
        if t > 0:
            want_mask = torch.ones(N, 1); want_mask[dones_env] = 0.0
            assert torch.equal(first[a]["masks"][t], want_mask)
            rel = (first[a]["rewards"][t - 1] - rew[:, a]).abs() / rew[:, a].abs().clamp(min=1e-6)
            assert float(rel.max()) <= 1e-5, "reward[%d] agent %d: %g" % (t - 1, a, float(rel.max()))
assert n_done_env > 0, "the frames were meant to end some episodes inside the first rollout"
assert "some episodes done" in log.getvalue()
print("RUNNER_OK algo=%s fused=%s episodes=%d" % (ALGO, FUSED, EPISODES))
'''


@needs_ref
@pytest.mark.parametrize("algo,fused", [("ippo", 0), ("ippo", 1), ("mappo", 1), ("happo", 1), ("mappo", 2), ("ippo", 2)])
def test_runner_mirror_with_ippo_dispatch(cuda_device, tmp_path, algo, fused):
    """`runner.Runner` (the reference's Runner interface + the `ippo` branch its dispatch lacks, device-side insert masks and
    episode bookkeeping) trains TenAnt for 2 episodes with the reference's own trainers / policies - optionally with this
    library's fused update bodies (1), and with the team forward + shared buffer for the rollout (2) - and the first episode
    it stored equals the reference's own env classes on the same frames and actions."""
    res = subprocess.run([sys.executable, "-c", _IPPO_SCRIPT, ROOT, REF, str(tmp_path), algo, str(int(fused))],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "RUNNER_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]

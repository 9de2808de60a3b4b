"""GPU parity of OneAnt (BASELINE config 0 shapes: N=64) and MultiIngenuity against the golden vectors made
from the reference, and against the oracle run as torch-eager-on-GPU."""
import pytest
import torch

from conftest import assert_close_obs, load_golden

pytestmark = pytest.mark.gpu


def _cfg(N, name):
    return {"env": {"numEnvs": N, "env_name": name}, "sim": {"dt": 0.0166}, "seed": 1}


def test_one_ant_matches_reference_golden(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import OneAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython
    g = load_golden("one_ant_n64")
    F, N = g["rew"].shape
    dev = cuda_device
    prov = ReplayProvider({"root": g["root"], "dof": g["dof"], "sensor": g["sensor"]}, device=dev, loop=False)
    task = OneAnt(_cfg(N, "one_ant"), None, None, "cuda", 0, True, False, provider=prov, flavor=L.FLAVOR_CPU)
    env = VecTaskPython(task, dev)
    assert env.observation_space.shape == (60,) and env.action_space.shape == (8,)
    for t in range(F):
        task.reset_noise = (g["noise_pos"][t].to(dev), g["noise_vel"][t].to(dev))
        obs, rew, done, info = env.step(g["actions"][t].to(dev))
        torch.cuda.synchronize()
        n_res = int(g["n_reset"][t])
        assert torch.equal(task.reset_buf.cpu(), g["reset"][t]) and torch.equal(done.cpu(), g["reset"][t])
        assert torch.equal(task.progress_buf.cpu(), g["progress"][t])
        assert int(task.reset_count.item()) == n_res
        assert torch.equal(task.env_ids[:n_res].cpu(), g["env_ids"][t][:n_res])
        assert torch.equal(task.ant_box_indices[:2 * n_res].cpu(), g["ant_box_indices"][t][:2 * n_res])
        assert torch.equal(task.ant_indices[:n_res].cpu(), g["ant_indices"][t][:n_res])
        ids = g["env_ids"][t][:n_res]
        if n_res:
            assert torch.equal(task.dof_reset_staging.view(N, 16).cpu()[ids], g["dof_pushed"][t].view(N, 16)[ids])
        assert_close_obs(task.obs_buf, g["obs"][t], angle_cols=(7, 8, 9), what="one_ant obs t=%d" % t)
        assert_close_obs(obs, g["obs_clamped"][t], angle_cols=(7, 8, 9), what="one_ant clamped obs")
        assert torch.equal(task.forces.cpu(), g["forces"][t])
        assert_close_obs(task.potentials, g["potentials"][t], what="potentials")
        assert_close_obs(task.prev_potentials, g["prev_potentials"][t], what="prev_potentials")
        assert_close_obs(task.up_vec, g["up_vec"][t], what="up_vec")
        assert_close_obs(task.heading_vec, g["heading_vec"][t], what="heading_vec")
        # OneAnt's distance terms use copied positions only: reward reproducible on any IEEE platform
        assert_close_obs(rew, g["rew"][t], rtol=1e-5, atol=1e-5, what="one_ant reward t=%d" % t)
        if t == 0:
            task.progress_buf.copy_(g["progress_after0"].to(dev))
    assert info == {}


def test_one_ant_vs_eager_gpu_and_replay(cuda_device):
    from oracle.task_oracle import OneAntOracle
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import OneAnt
    dev = cuda_device
    for N in (1, 64, 1000):
        T = 5
        fr = synthetic.one_ant_frames(N, T, seed=3 + N, fall_prob=0.02)
        npos, nvel = synthetic.reset_noise(N, T, seed=9)
        task = OneAnt(_cfg(N, "one_ant"), provider=ReplayProvider(fr, device=dev, loop=False))
        orc = OneAntOracle(N, device="cuda")
        outs = []
        for t in range(T):
            a = torch.clamp(fr["actions"][t].to(dev), -1, 1)
            task.reset_noise = (npos[t].to(dev), nvel[t].to(dev))
            task.step(a)
            orc.step(a, fr["root"][t].to(dev), fr["dof"][t].to(dev), fr["sensor"][t].to(dev), noise=(npos[t].to(dev), nvel[t].to(dev)))
            assert torch.equal(task.reset_buf, orc.reset_buf) and torch.equal(task.progress_buf, orc.progress_buf)
            assert_close_obs(task.obs_buf, orc.obs_buf, angle_cols=(7, 8, 9), what="one_ant eager N=%d" % N)
            rel = (task.rew_buf - orc.rew_buf).abs() / orc.rew_buf.abs().clamp(min=1e-6)
            assert float(rel.max()) <= 1e-5, float(rel.max())
            assert_close_obs(task.potentials, orc.potentials, what="potentials (CUDA flavour: x * (1/dt))")
            outs.append((task.obs_buf.clone(), task.rew_buf.clone(), task.reset_buf.clone()))
        rep = OneAnt(_cfg(N, "one_ant"), provider=ReplayProvider(fr, device=dev, loop=False))
        frd = {k: v.to(dev) for k, v in fr.items()}
        obs = torch.zeros(T, N, 60, device=dev); rew = torch.zeros(T, N, device=dev)
        d8 = torch.zeros(T, N, device=dev, dtype=torch.uint8)
        rep.replay(frd, torch.clamp(frd["actions"], -1, 1), obs, rew, d8)
        for t in range(T):
            assert torch.equal(obs[t], outs[t][0]) and torch.equal(rew[t], outs[t][1]) and torch.equal(d8[t].long(), outs[t][2])
        for nm in ("pos_before", "box_before", "potentials", "prev_potentials", "progress_buf", "reset_buf", "up_vec"):
            assert torch.equal(getattr(rep, nm), getattr(task, nm)), nm


def test_ingenuity_matches_reference_golden(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import MultiIngenuity
    g = load_golden("ingenuity_n33")
    F, N = g["rew"].shape
    dev = cuda_device
    task = MultiIngenuity(_cfg(N, "multi_ingenuity"), provider=ReplayProvider({"root": g["root"]}, device=dev, loop=False),
                          flavor=L.FLAVOR_CPU)
    for t in range(F):
        task.step(g["actions"][t].to(dev))
        torch.cuda.synchronize()
        n_res = int(g["n_reset"][t])
        # the reset bit depends on the computed float target_dist > 8: bit-exact with the CPU association
        assert torch.equal(task.reset_buf.cpu(), g["reset"][t]), "ingenuity reset t=%d" % t
        assert torch.equal(task.progress_buf.cpu(), g["progress"][t])
        assert int(task.reset_count.item()) == n_res
        assert torch.equal(task.env_ids[:n_res].cpu(), g["env_ids"][t][:n_res])
        assert torch.equal(task.actor_indices[:4 * n_res].cpu(), g["actor_indices"][t][:4 * n_res])
        assert torch.equal(task.obs_buf.cpu(), g["obs"][t])          # obs = raw root rows
        assert torch.equal(task.forces_applied.cpu(), g["forces"][t])  # thrust map: exact (mul/clamp only)
        if n_res:
            assert torch.equal(task.dof_state.cpu(), g["dof_pushed"][t])
        assert_close_obs(task.rew_buf, g["rew"][t], what="ingenuity reward t=%d" % t)
        if t == 0:
            task.progress_buf.copy_(g["progress_after0"].to(dev))


def test_ingenuity_vs_eager_gpu_replay_and_wrapper(cuda_device):
    from oracle.task_oracle import IngenuityOracle
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import MultiIngenuity
    from massive_marl_benchmark_b200.vec_task import MultiVecTaskPython
    dev = cuda_device
    for N in (1, 63, 2048):
        T = 6
        fr = synthetic.ingenuity_frames(N, T, seed=21 + N)
        task = MultiIngenuity(_cfg(N, "mi"), provider=ReplayProvider({"root": fr["root"]}, device=dev, loop=False))
        orc = IngenuityOracle(N, device="cuda")
        outs = []
        for t in range(T):
            a = fr["actions"][t].to(dev)
            task.step(a)
            orc.step(a, fr["root"][t].to(dev))
            # CUDA flavour reproduces torch-CUDA's (a0+a2)+a1 association: reset bits exact vs eager GPU
            assert torch.equal(task.reset_buf, orc.reset_buf), "N=%d t=%d" % (N, t)
            assert torch.equal(task.progress_buf, orc.progress_buf)
            assert torch.equal(task.forces, orc.forces) and torch.equal(task.forces_applied, orc.last["forces"])
            assert_close_obs(task.rew_buf, orc.rew_buf, what="ingenuity reward vs eager")
            outs.append((task.obs_buf.clone(), task.rew_buf.clone(), task.reset_buf.clone()))
        rep = MultiIngenuity(_cfg(N, "mi"), provider=ReplayProvider({"root": fr["root"]}, device=dev, loop=False))
        frd = {k: v.to(dev) for k, v in fr.items()}
        obs = torch.zeros(T, N, 52, device=dev); rew = torch.zeros(T, N, device=dev)
        d64 = torch.zeros(T, N, device=dev, dtype=torch.long); forces = torch.zeros(T, N, 24, 3, device=dev)
        rep.replay(frd, frd["actions"], obs, rew, None, d64, forces)
        for t in range(T):
            assert torch.equal(obs[t], outs[t][0]) and torch.equal(rew[t], outs[t][1]) and torch.equal(d64[t], outs[t][2])
        assert torch.equal(rep.forces, task.forces) and torch.equal(rep.progress_buf, task.progress_buf)
    # the parametric multi-agent wrapper (SURVEY finding 7): 4 agents x 13 obs, 6 actions each
    N = 50
    fr = synthetic.ingenuity_frames(N, 3, seed=2)
    task = MultiIngenuity(_cfg(N, "mi"), is_multi_agent=True, provider=ReplayProvider({"root": fr["root"]}, device=dev))
    env = MultiVecTaskPython(task, dev)
    assert env.num_agents == 4 and env.observation_space[0].shape == (13,) and env.share_observation_space[0].shape == (52,)
    obs, state, _ = env.reset()
    assert obs.shape == (N, 4, 13) and state.shape == (N, 4, 52)
    a = fr["actions"][1].to(dev)
    obs, state, rew, done, info, _ = env.step([a[:, 6 * h:6 * h + 6] for h in range(4)])
    want = torch.clamp(fr["root"][1].to(dev).view(N, 52), -7, 7)
    assert torch.equal(obs.reshape(N, 52), want) and torch.equal(state[:, 2], want)
    assert rew.shape == (N, 4, 1) and done.shape == (N, 4)

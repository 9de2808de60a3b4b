"""The CUDA-graphed per-step path (vec_task.GraphedVecTaskPython / GraphedMultiVecTaskPython) against the eager per-step path,
which is itself pinned to the reference's golden vectors (test_gpu_ten_ant.py, test_gpu_one_ant_ingenuity.py): every output and
every piece of task state bit-equal step by step over several passes of the frame ring, Philox reset noise included (device-
resident counter vs the host's), plus the reference's `current_obs.copy_(next_obs)` loop on the graphed wrapper.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ten_ant(N, F, dev, multi, seed=3):
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import TenAnt
    fr = synthetic.ten_ant_frames(N, F, seed=seed)
    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant", "episodeLength": 11}, "sim": {"dt": 0.0166}, "seed": 5}
    task = TenAnt(cfg, None, None, "cuda", 0, True, multi, provider=ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, device=dev))
    return task, fr["actions"].to(dev)


def _one_ant(N, F, dev, seed=3):
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import OneAnt
    fr = synthetic.one_ant_frames(N, F, seed=seed)
    cfg = {"env": {"numEnvs": N, "env_name": "one_ant", "episodeLength": 9}, "sim": {"dt": 0.0166}, "seed": 5}
    return OneAnt(cfg, None, None, "cuda", 0, True, False, provider=ReplayProvider(fr, device=dev)), fr["actions"].to(dev)


def _ingenuity(N, F, dev, seed=3):
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.tasks import MultiIngenuity
    fr = synthetic.ingenuity_frames(N, F, seed=seed)
    cfg = {"env": {"numEnvs": N, "env_name": "multi_ingenuity", "episodeLength": 9}, "sim": {"dt": 0.0166}, "seed": 5}
    task = MultiIngenuity(cfg, None, None, "cuda", 0, True, True, provider=ReplayProvider({"root": fr["root"]}, device=dev))
    return task, fr["actions"].to(dev)


_STATE = ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "reset_count", "env_ids", "ant_box_indices", "ant_indices",
          "dof_reset_staging", "forces", "pos_before", "goal_before", "box_before", "potentials", "prev_potentials", "up_vec",
          "heading_vec", "actor_indices", "forces_applied", "randomize_buf", "root_states", "dof_state")


def _same_state(a, b, what):
    for k in _STATE:
        if hasattr(a, k):
            x, y = getattr(a, k), getattr(b, k)
            assert torch.equal(x, y), "%s: %s differs" % (what, k)


@pytest.mark.parametrize("N", [257, 4096, 9001])      # 9001: the multi-CTA reset scan advances the counter
def test_graphed_ten_ant_step_equals_eager(cuda_device, N):
    from massive_marl_benchmark_b200.vec_task import GraphedVecTaskPython, VecTaskPython
    dev, F = cuda_device, 6
    te, acts = _ten_ant(N, F, dev, False)
    tg, _ = _ten_ant(N, F, dev, False)
    eager, graphed = VecTaskPython(te, dev), GraphedVecTaskPython(tg, dev)
    for i in range(4 * F + 1):                       # passes 2-4 replay the graphs captured in pass 1
        a = acts[i % F]
        oe, re_, de, _ = eager.step(a)
        og, rg, dg, _ = graphed.step(a)
        torch.cuda.synchronize()
        assert torch.equal(oe, og) and torch.equal(re_, rg) and torch.equal(de, dg), "step %d" % i
        _same_state(te, tg, "step %d" % i)
    assert graphed._g.captures == F                  # one graph per ring slot, everything after is replay
    assert int(te.reset_count.item()) >= 0 and int(tg._step_counter_dev.item()) == te._step_count == tg._step_count
    # the episode length of 11 made every env reset at least once: the Philox staging rows compared above were exercised
    assert int((tg.dof_reset_staging != 0).sum().item()) > 0
    # static-tensor contract: the observation tensor is the same object every step
    assert graphed.step(acts[0])[0] is og


def test_graphed_multi_agent_ten_ant_equals_eager(cuda_device):
    from massive_marl_benchmark_b200.vec_task import GraphedMultiVecTaskPython, MultiVecTaskPython
    dev, N, F = cuda_device, 333, 5
    te, acts = _ten_ant(N, F, dev, True)
    tg, _ = _ten_ant(N, F, dev, True)
    eager, graphed = MultiVecTaskPython(te, dev), GraphedMultiVecTaskPython(tg, dev)
    r0, r1 = eager.reset(), graphed.reset()
    assert torch.equal(r0[0], r1[0]) and torch.equal(r0[1], r1[1])
    for i in range(3 * F):
        al = [acts[i % F][:, 8 * k:8 * k + 8].contiguous() for k in range(10)]
        e, g = eager.step(al), graphed.step(al)
        torch.cuda.synchronize()
        for k in range(4):
            assert e[k].shape == g[k].shape and torch.equal(e[k], g[k]), "step %d output %d" % (i, k)
        _same_state(te, tg, "step %d" % i)


def test_graphed_one_ant_and_ingenuity_equal_eager(cuda_device):
    from massive_marl_benchmark_b200.vec_task import (GraphedMultiVecTaskPython, GraphedVecTaskPython, MultiVecTaskPython,
                                                      VecTaskPython)
    dev, F = cuda_device, 5
    te, acts = _one_ant(64, F, dev)
    tg, _ = _one_ant(64, F, dev)
    eager, graphed = VecTaskPython(te, dev), GraphedVecTaskPython(tg, dev)
    for i in range(3 * F + 2):
        e, g = eager.step(acts[i % F]), graphed.step(acts[i % F])
        torch.cuda.synchronize()
        assert all(torch.equal(e[k], g[k]) for k in range(3)), "one_ant step %d" % i
        _same_state(te, tg, "one_ant step %d" % i)
    te, acts = _ingenuity(200, F, dev)
    tg, _ = _ingenuity(200, F, dev)
    eager, graphed = MultiVecTaskPython(te, dev), GraphedMultiVecTaskPython(tg, dev)
    for i in range(3 * F + 2):
        e, g = eager.step(acts[i % F]), graphed.step(acts[i % F])
        torch.cuda.synchronize()
        assert all(torch.equal(e[k], g[k]) for k in range(4)), "ingenuity step %d" % i
        _same_state(te, tg, "ingenuity step %d" % i)


def test_reference_loop_on_the_graphed_wrapper_stores_the_right_observations(cuda_device):
    """ppo.py:127-139: `current_obs = reset()`, then per step act(current_obs) -> step -> add_transitions(current_obs, ...) ->
    current_obs.copy_(next_obs).  With the graphed wrapper `next_obs` is a static tensor; `reset()` hands out a clone, so the
    stored observation of step t must still be the one the policy acted on."""
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.vec_task import GraphedVecTaskPython, VecTaskPython
    dev, N, F, T = cuda_device, 128, 4, 10
    te, acts = _ten_ant(N, F, dev, False)
    tg, _ = _ten_ant(N, F, dev, False)
    eager, graphed = VecTaskPython(te, dev), GraphedVecTaskPython(tg, dev)
    torch.manual_seed(0)
    want0 = eager.reset()
    torch.manual_seed(0)
    current_obs = graphed.reset()
    assert torch.equal(want0, current_obs)
    st = RolloutStorage(N, T, (388,), (0,), (80,), dev, sampler="sequential")
    want = [want0.clone()]
    for t in range(T):
        a = acts[t % F]
        z = torch.zeros(N, 1, device=dev)
        next_obs, rews, dones, _ = graphed.step(a)
        st.add_transitions(current_obs, torch.zeros(N, 0, device=dev), a, rews, dones, z, rews, a, a)
        current_obs.copy_(next_obs)
        want.append(eager.step(a)[0].clone())
    torch.cuda.synchronize()
    for t in range(T):
        assert torch.equal(st.observations[t], want[t]), "stored observation of step %d" % t

"""CPU, build container only: the oracle against the REFERENCE ITSELF (skipped where /root/reference is absent,
e.g. on the GPU box).  Runs the reference's own TenAnt / OneAnt / MultiIngenuity classes under oracle/refshim on
fresh random frames (different seeds from the committed fixtures) and demands bit-equality."""
import os

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "agents")), reason="reference tree not present")


@pytest.fixture(scope="module")
def shim():
    from oracle import refshim
    refshim.install(REF)
    return refshim


def test_ten_ant_class_bit_equal(shim):
    from massive_marl_benchmark_b200 import synthetic
    from oracle.task_oracle import TenAntOracle
    N, F = 19, 5
    task, g = shim.make_task("TenAnt", N, True)
    orc = TenAntOracle(N)
    fr = synthetic.ten_ant_frames(N, F, seed=4242, fall_prob=0.02)
    for t in range(F):
        g.push_frame(fr["root"][t], fr["dof"][t])
        st = torch.get_rng_state()
        task.step(fr["actions"][t])
        torch.set_rng_state(st)
        orc.step(fr["actions"][t], fr["root"][t], fr["dof"][t])
        assert torch.equal(task.obs_buf, orc.obs_buf) and torch.equal(task.rew_buf, orc.rew_buf)
        assert torch.equal(task.reset_buf, orc.reset_buf) and torch.equal(task.progress_buf, orc.progress_buf)


def test_one_ant_and_ingenuity_class_bit_equal(shim):
    from massive_marl_benchmark_b200 import synthetic
    from oracle.task_oracle import IngenuityOracle, OneAntOracle
    N, F = 21, 4
    task, g = shim.make_task("OneAnt", N, False)
    orc = OneAntOracle(N)
    fr = synthetic.one_ant_frames(N, F, seed=777, fall_prob=0.05)
    for t in range(F):
        g.push_frame(fr["root"][t], fr["dof"][t], fr["sensor"][t])
        st = torch.get_rng_state(); task.step(fr["actions"][t]); torch.set_rng_state(st)
        orc.step(fr["actions"][t], fr["root"][t], fr["dof"][t], fr["sensor"][t])
        assert torch.equal(task.obs_buf, orc.obs_buf) and torch.equal(task.rew_buf, orc.rew_buf)
        assert torch.equal(task.reset_buf, orc.reset_buf) and torch.equal(task.potentials, orc.potentials)
    task, g = shim.make_task("MultiIngenuity", N, False)
    orc = IngenuityOracle(N)
    fr = synthetic.ingenuity_frames(N, F, seed=778)
    for t in range(F):
        g.push_frame(fr["root"][t])
        task.step(fr["actions"][t]); orc.step(fr["actions"][t], fr["root"][t])
        assert torch.equal(task.obs_buf, orc.obs_buf) and torch.equal(task.rew_buf, orc.rew_buf)
        assert torch.equal(task.reset_buf, orc.reset_buf) and torch.equal(task.forces, orc.forces)

"""GroupedAdam (clip + Adam for many networks in two launches) against `nn.utils.clip_grad_norm_` + `torch.optim.Adam`, and
TeamUpdate (the whole MAPPO / IPPO team's update, minibatch-major, one grouped optimiser step) against the reference's own
per-agent trainers on the same minibatch permutations.  Tolerances: parameters after several steps within 2e-6 relative of
torch's (the kernel fuses the multiply-adds where torch's elementwise kernels do; sums differ in association only)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "agents")), reason="baseline/_ref not installed (python baseline/make_ref.py)")


def _nets(gen, dev, n):
    out = []
    for i in range(n):
        torch.manual_seed(100 + i)
        m = torch.nn.Sequential(torch.nn.LayerNorm(13), torch.nn.Linear(13, 37), torch.nn.ELU(), torch.nn.Linear(37, 5 if i % 2 == 0 else 1))
        out.append(m.to(dev))
    return out


@pytest.mark.parametrize("weight_decay,max_norm", [(0.0, 10.0), (0.01, 0.05), (0.0, 0.0)])
def test_grouped_adam_matches_torch_adam_with_clipping(cuda_device, weight_decay, max_norm):
    import copy
    from massive_marl_benchmark_b200.grouped_adam import GroupedAdam
    dev = cuda_device
    gen = torch.Generator().manual_seed(0)
    nets = _nets(gen, dev, 6)
    refs = copy.deepcopy(nets)
    lrs = [5e-4 * (1 + i) for i in range(6)]
    opt = GroupedAdam([dict(params=list(m.parameters()), lr=lr, eps=1e-5, weight_decay=weight_decay, max_grad_norm=max_norm)
                       for m, lr in zip(nets, lrs)])
    ropts = [torch.optim.Adam(m.parameters(), lr=lr, eps=1e-5, weight_decay=weight_decay) for m, lr in zip(refs, lrs)]
    for m, r in zip(nets, refs):       # re-pointing at the flat buffer preserved the values
        for p, q in zip(m.parameters(), r.parameters()):
            assert torch.equal(p, q)
    for step in range(5):
        x = torch.randn(64, 13, generator=gen).to(dev) * (3.0 if step == 2 else 1.0)
        opt.zero_grad()
        for m in nets:
            (m(x).pow(2).mean() * (50.0 if step == 2 else 1.0)).backward()      # step 2: large gradients -> the clip engages
        opt.collect_grads()
        opt.step()
        norms = opt.grad_norms()
        for i, (r, ro) in enumerate(zip(refs, ropts)):
            ro.zero_grad()
            (r(x).pow(2).mean() * (50.0 if step == 2 else 1.0)).backward()
            if max_norm > 0:
                tn = torch.nn.utils.clip_grad_norm_(r.parameters(), max_norm)
            else:
                tn = torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in r.parameters()]))
            assert abs(float(norms[i]) - float(tn)) <= 1e-5 * float(tn) + 1e-9, (step, i)
            ro.step()
        for m, r in zip(nets, refs):
            for p, q in zip(m.parameters(), r.parameters()):
                assert torch.allclose(p, q, rtol=2e-6, atol=1e-7), (step, float((p - q).abs().max()))
    # the modules still work as modules: state_dict round trip, and version counters moved with every step
    sd = nets[0].state_dict()
    assert all(torch.allclose(v, refs[0].state_dict()[k], rtol=2e-6, atol=1e-7) for k, v in sd.items())
    assert next(nets[0].parameters())._version >= 5
    st = opt.state_dict()
    opt.load_state_dict(st)
    assert st["steps"] == [5] * 6


def test_grouped_adam_rejects_bad_layouts(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.grouped_adam import GroupedAdam
    with pytest.raises(L.MmbError):
        GroupedAdam([dict(params=[torch.nn.Parameter(torch.zeros(3))])])          # CPU parameter: no CPU path
    p = L.AdamParams()
    assert L.lib().mmb_adam_group(p, None) == -1 and L.lib().mmb_grad_sumsq_group(None, None) == -1


_TEAM_SCRIPT = r'''
import contextlib, copy, io, os, sys, torch, yaml
ROOT, REF, ALGO = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
from oracle import refshim
refshim.install(REF)
from massive_marl_benchmark_b200 import spaces
from massive_marl_benchmark_b200.runner import resolve_algorithm
from massive_marl_benchmark_b200.separated_buffer import SeparatedReplayBuffer
from massive_marl_benchmark_b200.team_update import TeamUpdate
import numpy as np

dev = torch.device("cuda", 0)
A, N = 4, 48
config = yaml.safe_load(open(os.path.join(REF, "cfg", ALGO, "config.yaml")))
T = config["episode_length"]
config.update(n_rollout_threads=N, ppo_epoch=3, num_mini_batch=2, hidden_size=64)
TrainAlgo, Policy = resolve_algorithm(ALGO)
ob = spaces.Box(low=-np.inf, high=np.inf, shape=(46,)); sh = spaces.Box(low=-np.inf, high=np.inf, shape=(388,))
ac = spaces.Box(low=-np.ones(8), high=np.ones(8))
cent = sh if config["use_centralized_V"] else ob

def team(seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        pols = [Policy(config, ob, cent, ac, device=dev) for _ in range(A)]
        trs = [TrainAlgo(config, p, device=dev) for p in pols]
    return pols, trs

gen = torch.Generator().manual_seed(7)
def fill(bufs):
    g = torch.Generator().manual_seed(11)
    for b in bufs:
        b.share_obs.copy_(torch.randn(b.share_obs.shape, generator=g)); b.obs.copy_(torch.randn(b.obs.shape, generator=g))
        b.actions.copy_(torch.randn(b.actions.shape, generator=g) * 0.3)
        b.action_log_probs.copy_(-torch.rand(b.action_log_probs.shape, generator=g))
        b.value_preds.copy_(torch.randn(b.value_preds.shape, generator=g) * 0.5)
        b.rewards.copy_(torch.randn(b.rewards.shape, generator=g))
        b.masks.copy_((torch.rand(b.masks.shape, generator=g) > 0.05).float())

ref_pols, ref_trs = team(3)
our_pols, our_trs = team(3)
for p, q in zip(ref_pols, our_pols):
    assert all(torch.equal(x, y) for x, y in zip(p.actor.parameters(), q.actor.parameters()))
ref_bufs = [SeparatedReplayBuffer(config, ob, cent, ac, dev) for _ in range(A)]
our_bufs = [SeparatedReplayBuffer(config, ob, cent, ac, dev) for _ in range(A)]
fill(ref_bufs); fill(our_bufs)
perms = [[torch.randperm(T * N, generator=gen) for _ in range(config["ppo_epoch"])] for _ in range(A)]

class PermFeed:       # the same minibatch permutations for both sides: epoch e of agent a draws perms[a][e]
    def __init__(self, buf, seq):
        self.buf, self.seq, self.i = buf, seq, 0
        self.orig = buf.feed_forward_generator
    def __call__(self, advantages, num_mini_batch=None, mini_batch_size=None):
        self.buf.permutation_override = self.seq[self.i]; self.i += 1
        return self.orig(advantages, num_mini_batch, mini_batch_size)

# ---- reference: the reference's own trainers, agent after agent (runner.py:257-317 without the unused factor) ----
for a in range(A):
    nv = torch.randn(N, 1, generator=torch.Generator().manual_seed(50 + a)).to(dev)
    ref_bufs[a].compute_returns(nv, ref_trs[a].value_normalizer)
    our_bufs[a].compute_returns(nv, our_trs[a].value_normalizer)
    ref_bufs[a].feed_forward_generator = PermFeed(ref_bufs[a], perms[a])
    our_bufs[a].feed_forward_generator = PermFeed(our_bufs[a], perms[a])
ref_infos = []
for a in range(A):
    ref_trs[a].prep_training()
    ref_bufs[a].update_factor(torch.ones(T, N, 1, device=dev))
    ref_infos.append(ref_trs[a].train(ref_bufs[a]))
    ref_bufs[a].after_update()

# ---- ours: the whole team, minibatch-major, one grouped optimiser step per minibatch ----
tu = TeamUpdate(our_trs, our_bufs, config, algorithm=ALGO)
our_infos = tu.train()
torch.cuda.synchronize()
worst = 0.0
for a in range(A):
    for net in ("actor", "critic"):
        for (k, x), (_, y) in zip(getattr(ref_pols[a], net).state_dict().items(), getattr(our_pols[a], net).state_dict().items()):
            err = float(((x - y).abs() / (x.abs() + 1e-3)).max())
            worst = max(worst, err)
            assert err <= 2e-4, (a, net, k, err)
    for key in ("value_loss", "policy_loss", "dist_entropy", "ratio"):
        r, o = float(ref_infos[a][key]), float(our_infos[a][key])
        assert abs(r - o) <= 2e-4 * abs(r) + 2e-5, (a, key, r, o)
    for key in ("actor_grad_norm", "critic_grad_norm"):
        r, o = float(ref_infos[a][key]), float(our_infos[a][key])
        assert abs(r - o) <= 2e-3 * abs(r) + 1e-5, (a, key, r, o)
    if our_trs[a].value_normalizer is not None:
        for x, y in zip(ref_trs[a].value_normalizer.running_mean_var(), our_trs[a].value_normalizer.running_mean_var()):
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-6)
assert tu.replica_checksum() == (0.0, 0.0)
print("TEAM_OK algo=%s worst_param_rel_err=%.3g" % (ALGO, worst))
'''


@needs_ref
@pytest.mark.parametrize("algo", ["mappo", "ippo"])
def test_team_update_equals_the_reference_trainers(cuda_device, algo):
    res = subprocess.run([sys.executable, "-c", _TEAM_SCRIPT, ROOT, REF, algo], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "TEAM_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    print(res.stdout.strip().splitlines()[-1])

"""CPU: the C-ABI library loads and exports every symbol include/mmb.h declares, the ctypes structs have the
C layout, host-side logic (providers, spaces, sharding, synthetic layouts) behaves, and the product fails loudly
without a GPU instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmb.h")


@pytest.fixture(scope="module")
def built_lib():
    from massive_marl_benchmark_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(HEADER).read()
    declared = set(re.findall(r"MMB_API\s+[\w\s\*]+?\b(mmb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(built_lib.SYMBOLS), declared ^ set(built_lib.SYMBOLS)
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mmb_abi_version.restype = ctypes.c_int32
    assert lib.mmb_abi_version() == built_lib.ABI_VERSION
    lib.mmb_strerror.restype = ctypes.c_char_p
    assert lib.mmb_strerror(0) == b"ok" and b"alignment" in lib.mmb_strerror(-2)
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == declared, exported ^ declared      # nothing but the ABI is exported


def test_ctypes_structs_match_c_layout(built_lib):
    """Compile a C program against include/mmb.h that prints sizeof/offsetof and compare with ctypes."""
    L = built_lib
    probes = [("mmb_ant_consts", L.AntConsts, "initial_dof_pos"), ("mmb_ten_ant_params", L.TenAntParams, "c"),
              ("mmb_one_ant_params", L.OneAntParams, "c"), ("mmb_ingenuity_params", L.IngenuityParams, "forces_state"),
              ("mmb_reset_params", L.ResetParams, "c"), ("mmb_rollout_add_params", L.RolloutAddParams, "values_stride"),
              ("mmb_gae_ppo_params", L.GaePpoParams, "stats"), ("mmb_gae_marl_params", L.GaeMarlParams, "stats"),
              ("mmb_xchg", L.Xchg, "mailbox"), ("mmb_episode_params", L.EpisodeParams, "state"), ("mmb_gaussian_act_params", L.GaussianActParams, "step_counter"),
              ("mmb_ppo_loss_params", L.PpoLossParams, "ticket"), ("mmb_mappo_loss_params", L.MappoLossParams, "sums"),
              ("mmb_mappo_loss_params", L.MappoLossParams, "ticket"), ("mmb_gaussian_act_params", L.GaussianActParams, "std_group_rows"),
              ("mmb_mlp_layer_params", L.MlpLayerParams, "overlap_prev"), ("mmb_mlp_layer_params", L.MlpLayerParams, "ln_in_eps"),
              ("mmb_mlp_layer_params", L.MlpLayerParams, "ln_out_stats"), ("mmb_reset_params", L.ResetParams, "step_counter"),
              ("mmb_copy_seg", L.CopySeg, "src_s1"), ("mmb_copy_group_params", L.CopyGroupParams, "seg"),
              ("mmb_gather_params", L.GatherParams, "group"), ("mmb_adam_params", L.AdamParams, "one_minus_beta1")]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "mmb.h"\nint main(void){\n'
    for cname, _, field in probes:
        src += 'printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (cname, cname, field)
    src += 'printf("%d %d\\n", (int)MMB_ACT_COUNTER_WORDS, (int)MMB_ABI_VERSION);\n'
    src += "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")], text=True).split("\n")
    for (cname, cls, field), line in zip(probes, out):
        size, off = map(int, line.split())
        assert ctypes.sizeof(cls) == size, cname
        assert getattr(cls, field).offset == off, (cname, field)
    words, abi = map(int, out[len(probes)].split())
    assert words == L.ACT_COUNTER_WORDS and abi == L.ABI_VERSION      # constants the binding restates


def test_no_cpu_fallback(built_lib):
    """The product refuses to run without CUDA rather than computing on the CPU."""
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt
    with pytest.raises(built_lib.MmbError):
        RolloutStorage(4, 2, (3,), (0,), (2,), "cpu")
    with pytest.raises(built_lib.MmbError):
        TenAnt({"env": {"numEnvs": 4}}, None, None, "cpu", 0, True)
    import massive_marl_benchmark_b200 as pkg
    for mod in ("tasks", "vec_task", "storage", "separated_buffer", "_lib", "providers", "dist", "synthetic"):
        text = open(os.path.join(os.path.dirname(pkg.__file__), mod + ".py")).read()
        assert "import oracle" not in text and "from oracle" not in text, mod


def test_missing_library_fails_loudly(built_lib, monkeypatch):
    monkeypatch.setattr(built_lib, "_lib", None)
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libmmb_b200.so")
    with pytest.raises(built_lib.MmbError, match="no CPU fallback"):
        built_lib.lib()


def test_synthetic_layouts_and_providers():
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import ReplayProvider
    fr = synthetic.ten_ant_frames(6, 3, seed=1)
    assert fr["root"].shape == (3, 66, 13) and fr["dof"].shape == (3, 480, 2) and fr["actions"].shape == (3, 6, 80)
    assert torch.allclose(fr["root"][:, :, 3:7].norm(dim=-1), torch.ones(3, 66), atol=1e-5)
    assert (fr["root"][:, 10::11, 2] == 1.0).all() and (fr["root"][:, 10::11, 3:5] == 0).all()
    lo, hi = synthetic.ant_dof_limits()
    pos = fr["dof"][..., 0].view(3, 60, 8)
    assert (pos >= lo - 0.03).all() and (pos <= hi + 0.03).all()
    assert torch.equal(synthetic.ant_initial_dof_pos(), torch.where(lo > 0, lo, torch.where(hi < 0, hi, torch.zeros(8))))
    assert synthetic.one_ant_frames(5, 2)["sensor"].shape == (2, 20, 6)
    assert synthetic.ingenuity_frames(5, 2)["root"].shape == (2, 20, 13)
    again = synthetic.ten_ant_frames(6, 3, seed=1)
    assert all(torch.equal(fr[k], again[k]) for k in fr)
    p = ReplayProvider({"root": fr["root"], "dof": fr["dof"]}, loop=True)
    with pytest.raises(RuntimeError):
        p.frame()
    for t in range(5):
        p.simulate()
        assert torch.equal(p.frame()["root"], fr["root"][t % 3])
    assert p.window(1, 2)["dof"].shape[0] == 2
    with pytest.raises(IndexError):
        p.window(2, 2)
    q = ReplayProvider({"root": fr["root"]}, loop=False)
    for _ in range(4):
        q.simulate()
    with pytest.raises(IndexError):
        q.frame()


def test_spaces_and_shard_ranges():
    from massive_marl_benchmark_b200 import dist as mdist
    from massive_marl_benchmark_b200 import spaces
    b = spaces.Box(low=-np.inf, high=np.inf, shape=(46,))
    assert b.shape == (46,) and isinstance(b, spaces.Space)
    assert spaces.Box(np.ones(8) * -1.0, np.ones(8)).shape == (8,)
    for n, w in ((4096, 8), (16384, 4), (10, 3), (7, 8)):
        spans = [mdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_default_consts_match_reference_yaml():
    from massive_marl_benchmark_b200 import _lib as L
    c = L.default_ant_consts()
    assert abs(c.up_weight - 0.1) < 1e-7 and abs(c.death_cost + 2.0) < 1e-7 and c.max_episode_length == 1000.0
    assert np.allclose(list(c.initial_dof_pos), [0, 0.5236, 0, -0.5236, 0, -0.5236, 0, 0.5236], atol=1e-4)
    assert list(c.joint_gears) == [15.0] * 8 and list(c.inv_start_rot)[3] == 1.0
    assert np.signbit(np.float32(list(c.inv_start_rot)[0]))        # quat_conjugate((0,0,0,1)) has negative zeros
    c2 = L.default_ant_consts({"upWeight": 0.25, "episodeLength": 500})
    assert abs(c2.up_weight - 0.25) < 1e-7 and c2.max_episode_length == 500.0


def test_ppo_checkpoint_format(tmp_path):
    """`model_{it}.pt` = the state dict of two Sequentials with Linear layers at the even indices + `log_std`
    (ppo.py:90-97, module.py:25-55): written, read back, sizes inferred from the tensors, malformed files refused."""
    import pytest
    import torch
    from massive_marl_benchmark_b200 import checkpoints as ck
    m = ck.PPOModules([60, 1024, 1024, 512, 8], [60, 1024, 1024, 512, 1], 8)
    assert sorted(m.state_dict()) == sorted(["log_std"] + ["%s.%d.%s" % (n, i, p) for n in ("actor", "critic")
                                                           for i in (0, 2, 4, 6) for p in ("weight", "bias")])
    with torch.no_grad():
        m.log_std.fill_(-0.3)
    path = ck.save_ppo(m, str(tmp_path), 42)
    m2, it = ck.load_ppo(path)
    assert it == 42 and not m2.asymmetric
    x = torch.randn(3, 60)
    assert torch.equal(m2.actor(x), m.actor(x)) and torch.equal(m2.critic(x), m.critic(x)) and float(m2.log_std[0]) == float(m.log_std[0])
    assert [tuple(l.weight.shape) for l in m2.actor if hasattr(l, "weight")] == [(1024, 60), (1024, 1024), (512, 1024), (8, 512)]
    sd = dict(m.state_dict())
    bad = dict(sd); bad.pop("actor.2.weight")
    with pytest.raises(ValueError):
        ck.ppo_modules_from_state_dict(bad)
    bad = dict(sd); bad["log_std"] = torch.zeros(3)
    with pytest.raises(ValueError):
        ck.ppo_modules_from_state_dict(bad)
    bad = dict(sd); bad["critic.6.weight"] = torch.zeros(2, 512); bad["critic.6.bias"] = torch.zeros(2)
    with pytest.raises(ValueError):
        ck.ppo_modules_from_state_dict(bad)


def test_deferred_layernorm_fold_identity_on_cpu(built_lib):
    """Host side of the deferred LayerNorm (include/mmb.h `ln_in_stats`): with W' = bf16(W * gamma), c = row sums of W' and
    bias' = b + W beta as `_Layer.refresh` builds them, rstd * (e . W'^T - mean * c) + bias' equals LayerNorm(e) . W^T + b up
    to the bf16 rounding of W' - and exactly when W' is taken unrounded.  Re-folds when a source parameter changes."""
    import torch
    from massive_marl_benchmark_b200 import mlp as mm
    gen = torch.Generator().manual_seed(0)
    sd = {"base.feature_norm.weight": torch.ones(46), "base.feature_norm.bias": torch.zeros(46)}
    for n, k in (("base.mlp.fc1", 46), ("base.mlp.fc2.0", 512), ("base.mlp.fc2.1", 512)):
        sd[n + ".0.weight"] = torch.randn(512, k, generator=gen) / k ** 0.5
        sd[n + ".0.bias"] = 0.1 * torch.randn(512, generator=gen)
        sd[n + ".2.weight"] = 1.0 + 0.2 * torch.randn(512, generator=gen)
        sd[n + ".2.bias"] = 0.1 * torch.randn(512, generator=gen)
    sd["act.action_out.fc_mean.weight"] = 0.05 * torch.randn(8, 512, generator=gen)
    sd["act.action_out.fc_mean.bias"] = 0.1 * torch.randn(8, generator=gen)
    f = mm.FusedMLP.from_marl_state_dict(sd, "act.action_out.fc_mean", device="cpu")
    assert [l.fold is not None for l in f.layers] == [False, True, True, True]
    assert [l.stats_out for l in f.layers] == [True, True, True, False] and all(l.epilogue == (1 if l.act else 0) for l in f.layers)
    e = torch.nn.functional.elu(torch.randn(64, 512, generator=gen))
    for l, ln, lin in ((f.layers[1], "base.mlp.fc1.2", "base.mlp.fc2.0.0"), (f.layers[3], "base.mlp.fc2.1.2", "act.action_out.fc_mean")):
        W, b, gamma, beta = sd[lin + ".weight"], sd[lin + ".bias"], sd[ln + ".weight"], sd[ln + ".bias"]
        want = torch.nn.functional.layer_norm(e, (512,), gamma, beta, 1e-5) @ W.T + b
        mean, rstd = e.mean(1, keepdim=True), torch.rsqrt(e.var(1, unbiased=False, keepdim=True) + 1e-5)
        wf = l.w[:l.n_src, :l.K].float()
        got = rstd * (e @ wf.T - mean * l.c[:l.n_src][None, :]) + l.bias[:l.n_src][None, :]
        assert torch.allclose(got, want, rtol=0, atol=2e-2 * float(want.abs().max()))           # bf16 weights
        exact = rstd * (e @ (W * gamma[None, :]).T - mean * (W * gamma[None, :]).sum(1)[None, :]) + (b + W @ beta)[None, :]
        assert torch.allclose(exact, want, rtol=1e-4, atol=1e-4)
        assert torch.allclose(l.c[:l.n_src], wf.sum(1)) and torch.allclose(l.bias[:l.n_src], b + W @ beta, atol=1e-6)
    # a LayerNorm parameter changed in place -> stale -> the NEXT layer is re-folded
    assert not f.stale()
    old_c = f.layers[1].c.clone()
    sd["base.mlp.fc1.2.weight"].mul_(2.0)
    assert f.stale()
    f.refresh()
    assert not f.stale() and torch.allclose(f.layers[1].c, 2.0 * old_c, rtol=2e-2, atol=1e-3)

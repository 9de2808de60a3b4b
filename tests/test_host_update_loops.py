"""CPU: the host logic of the two drop-in update loops (massive_marl_benchmark_b200.ppo_update / mappo_update) with the
fused loss launch replaced by the oracle's statements - everything around the kernel (gathers, network calls, PopArt call
protocol, adaptive-KL schedule, the backward / clip / step sequences, loss bookkeeping, autograd attachment of
precomputed gradients) must reproduce the reference-pinned update oracles bit for bit.  The kernels themselves are
covered by the -m gpu tests."""
import copy
import importlib
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_attach_routes_precomputed_gradients_and_allows_separate_backwards():
    from massive_marl_benchmark_b200.mappo_loss import _Attach
    m = torch.randn(5, 3, requires_grad=True)
    ls = torch.randn(3, requires_grad=True)
    v = torch.randn(5, 1, requires_grad=True)
    s = torch.sigmoid(ls) * 0.5
    gm, gs, gv = torch.randn(5, 3), torch.randn(3), torch.randn(5, 1)
    pl = _Attach.apply(torch.tensor(1.5), 2, m, s, gm, gs)
    vl = _Attach.apply(torch.tensor(2.5), 1, v, gv)
    assert float(pl) == 1.5 and float(vl) == 2.5
    (pl * 2.0).backward()                       # two independent backward passes, as the reference trainer makes them
    (vl * 0.5).backward()
    assert torch.equal(m.grad, 2.0 * gm) and torch.equal(v.grad, 0.5 * gv)
    want, = torch.autograd.grad((torch.sigmoid(ls) * 0.5 * (2.0 * gs)).sum(), ls)
    assert torch.allclose(ls.grad, want)


def test_ppo_update_host_loop_equals_the_pinned_oracle(monkeypatch):
    t = importlib.import_module("test_gpu_ppo_update")
    from massive_marl_benchmark_b200 import ppo_update as pu
    from massive_marl_benchmark_b200.ppo_loss import PpoLossOut
    from oracle.ppo_loss_oracle import ppo_loss_terms, ppo_update_oracle

    def stand_in(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma, **cfg):
        loss, s, v, kl, logp, ent = ppo_loss_terms(mu, log_std, value, actions, old_logp, adv, tv, ret, old_mu, old_sigma, **cfg)
        return PpoLossOut(loss, s.detach(), v.detach(), kl.detach(), logp.detach(), ent.detach()[0])

    monkeypatch.setattr(pu, "ppo_loss", stand_in)
    T, N, obs_dim, A = 6, 40, 12, 8
    torch.manual_seed(11)
    ac1 = t._ActorCritic(obs_dim, A, hidden=(16, 8))
    ac2 = copy.deepcopy(ac1)
    st = types.SimpleNamespace(num_envs=N, num_transitions_per_env=T, states=torch.zeros(T, N, 0))
    widths = {"observations": obs_dim, "actions": A, "mu": A, "sigma": A}
    for k in ("observations", "actions", "actions_log_prob", "values", "returns", "advantages", "mu", "sigma"):
        setattr(st, k, torch.zeros(T, N, widths.get(k, 1)))
    t._fill(st, ac1, T, N, obs_dim, A, seed=21)
    gen = torch.Generator().manual_seed(9)
    orders = [torch.randperm(T * N, generator=gen) for _ in range(2)]
    mb = (T * N) // 3

    class _Batches:                                      # iterated once per epoch, a new order each time (storage.py:75-87)
        def __init__(self):
            self.epoch = 0

        def __iter__(self):
            o = orders[self.epoch]
            self.epoch += 1
            return iter([o[i * mb:(i + 1) * mb] for i in range(3)])

    st.mini_batch_generator = lambda n: _Batches()
    a = t._ppo(st, ac1, torch.optim.Adam(ac1.parameters(), lr=2e-3))
    b = t._ppo(st, ac2, torch.optim.Adam(ac2.parameters(), lr=2e-3))
    assert pu.ppo_update(a) == ppo_update_oracle(b, orders)
    assert a.step_size == b.step_size != 2e-3
    assert all(torch.equal(x, y) for x, y in zip(ac1.state_dict().values(), ac2.state_dict().values()))


def test_mappo_update_host_loop_equals_the_pinned_oracle(monkeypatch):
    t = importlib.import_module("test_gpu_mappo_update")
    from massive_marl_benchmark_b200 import mappo_update as mu
    from massive_marl_benchmark_b200.mappo_loss import MappoLossOut
    from oracle.mappo_loss_oracle import mappo_loss_terms, mappo_update_oracle

    def stand_in(mean, std, values, actions, old_logp, adv, vp, ret, active, m1=None, v1=None, m2=None, v2=None, **cfg):
        ls = torch.log(std / 0.5 / (1 - std / 0.5))     # logit: sigmoid(ls) * 0.5 == std, the graph continues through std
        pl, ent, vl, imp, lp = mappo_loss_terms(mean, ls, values, actions, old_logp, adv, vp, ret, active, m1, v1, m2, v2, **cfg)
        return MappoLossOut(pl, vl, ent, imp.detach(), lp.detach())

    monkeypatch.setattr(mu, "mappo_loss", stand_in)
    for over in (dict(), dict(_use_value_active_masks=True, _use_policy_active_masks=True, huber_delta=0.5),
                 dict(nopop=True, _use_huber_loss=False, _use_clipped_value_loss=False)):
        over = dict(over)
        nopop = over.pop("nopop", False)
        torch.manual_seed(4)
        a1, c1 = t._Actor(12, 6, hidden=16), t._Critic(20, hidden=16)
        a2, c2 = copy.deepcopy(a1), copy.deepcopy(c1)

        def state():
            if nopop:
                return None
            deb = torch.tensor(1.0 - 0.99999 ** 300)
            return {"running_mean": torch.tensor([0.9]) * deb, "running_mean_sq": torch.tensor([4.5]) * deb,
                    "debiasing_term": deb.clone()}

        ora, drop = t._trainer(a1, c1, state(), **over), t._trainer(a2, c2, state(), **over)
        for it in range(3):
            s = t._sample(a1, c1, 64, 12, 20, 6, 200 + it)
            want, got = mappo_update_oracle(ora, s), mu.mappo_ppo_update(drop, s)
            for x, y in zip(want[:5], got[:5]):
                assert abs(float(x) - float(y)) <= 1e-6 * abs(float(x)), over       # (the logit round trip of the stand-in)
        for x, y in zip(list(a1.state_dict().values()) + list(c1.state_dict().values()),
                        list(a2.state_dict().values()) + list(c2.state_dict().values())):
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-7), over
        if not nopop:                                    # PopArt was called exactly as often on both sides
            assert all(torch.equal(ora.popart[k], drop.popart[k]) for k in ora.popart)


def test_ippo_update_host_loop_equals_the_pinned_oracle(monkeypatch):
    """IPPO's normaliser protocol (one `update`, both error terms on the same moments) through the drop-in."""
    t = importlib.import_module("test_gpu_mappo_update")
    from massive_marl_benchmark_b200 import mappo_update as mu
    from massive_marl_benchmark_b200.mappo_loss import MappoLossOut
    from oracle.mappo_loss_oracle import mappo_loss_terms, mappo_update_oracle, popart_update

    seen = []

    def stand_in(mean, std, values, actions, old_logp, adv, vp, ret, active, m1=None, v1=None, m2=None, v2=None, **cfg):
        seen.append((m1 is not None, m2 is not None))
        ls = torch.log(std / 0.5 / (1 - std / 0.5))
        pl, ent, vl, imp, lp = mappo_loss_terms(mean, ls, values, actions, old_logp, adv, vp, ret, active, m1, v1, m2, v2, **cfg)
        return MappoLossOut(pl, vl, ent, imp.detach(), lp.detach())

    class _ValueNorm(t._PopArt):                         # valuenorm.py:39-55: update() only, no normalising call
        def update(self, x):
            popart_update(self.state, x)

        def __call__(self, x):
            raise AssertionError("IPPO never calls the normaliser in training mode")

    monkeypatch.setattr(mu, "mappo_loss", stand_in)
    torch.manual_seed(6)
    a1, c1 = t._Actor(12, 6, hidden=16), t._Critic(12, hidden=16)
    a2, c2 = copy.deepcopy(a1), copy.deepcopy(c1)

    def state():
        return {"running_mean": torch.zeros(1), "running_mean_sq": torch.zeros(1), "debiasing_term": torch.tensor(0.0)}

    ora = t._trainer(a1, c1, state(), _use_popart=False, _use_valuenorm=True)
    drop = t._trainer(a2, c2, state(), _use_popart=False, _use_valuenorm=True)
    drop.value_normalizer = _ValueNorm(drop.popart)
    for it in range(3):
        s = t._sample(a1, c1, 64, 12, 12, 6, 400 + it)
        want, got = mappo_update_oracle(ora, s, ippo=True), mu.ippo_ppo_update(drop, s)
        for x, y in zip(want[:5], got[:5]):
            assert abs(float(x) - float(y)) <= 1e-6 * abs(float(x))
    assert seen == [(True, False)] * 3
    assert all(torch.equal(ora.popart[k], drop.popart[k]) for k in ora.popart) and float(ora.popart["debiasing_term"]) > 0
    for x, y in zip(list(a1.state_dict().values()) + list(c1.state_dict().values()),
                    list(a2.state_dict().values()) + list(c2.state_dict().values())):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-7)


def test_happo_update_host_loop_equals_the_pinned_oracle(monkeypatch):
    """HAPPO's factor folded into the advantage (positive factor) against the oracle that keeps it inside the surrogate
    as the reference does."""
    t = importlib.import_module("test_gpu_mappo_update")
    from massive_marl_benchmark_b200 import mappo_update as mu
    from massive_marl_benchmark_b200.mappo_loss import MappoLossOut
    from oracle.mappo_loss_oracle import mappo_loss_terms, mappo_update_oracle

    def stand_in(mean, std, values, actions, old_logp, adv, vp, ret, active, m1=None, v1=None, m2=None, v2=None, **cfg):
        ls = torch.log(std / 0.5 / (1 - std / 0.5))
        pl, ent, vl, imp, lp = mappo_loss_terms(mean, ls, values, actions, old_logp, adv, vp, ret, active, m1, v1, m2, v2, **cfg)
        return MappoLossOut(pl, vl, ent, imp.detach(), lp.detach())

    monkeypatch.setattr(mu, "mappo_loss", stand_in)
    for over in (dict(), dict(_use_policy_active_masks=True)):
        torch.manual_seed(7)
        a1, c1 = t._Actor(12, 6, hidden=16), t._Critic(20, hidden=16)
        a2, c2 = copy.deepcopy(a1), copy.deepcopy(c1)

        def state():
            deb = torch.tensor(1.0 - 0.99999 ** 300)
            return {"running_mean": torch.tensor([0.9]) * deb, "running_mean_sq": torch.tensor([4.5]) * deb,
                    "debiasing_term": deb.clone()}

        ora, drop = t._trainer(a1, c1, state(), **over), t._trainer(a2, c2, state(), **over)
        del drop._use_valuenorm                           # the HAPPO trainer has no such attribute (happo_trainer.py:30-42)
        for it in range(3):
            s = t._sample(a1, c1, 64, 12, 20, 6, 700 + it)
            g = torch.Generator().manual_seed(800 + it)
            factor = torch.exp(0.3 * torch.randn(64, 1, generator=g))
            s = s[:12] + (factor if it else factor.repeat(1, 6) / 6,)
            want, got = mappo_update_oracle(ora, s, happo=True), mu.happo_ppo_update(drop, s)
            for x, y in zip(want[:5], got[:5]):
                assert abs(float(x) - float(y)) <= 2e-6 * abs(float(x)) + 1e-8, over
        for x, y in zip(list(a1.state_dict().values()) + list(c1.state_dict().values()),
                        list(a2.state_dict().values()) + list(c2.state_dict().values())):
            assert torch.allclose(x, y, rtol=1e-5, atol=2e-7), over

"""GPU parity of the rollout-storage path: RolloutStorage (PPO), SeparatedReplayBuffer (MARL), masks,
shuffle + gather.  Oracles: golden vectors from the reference classes and oracle/storage_oracle.py."""
import numpy as np
import pytest
import torch

from conftest import assert_close_obs, load_golden

pytestmark = pytest.mark.gpu


def test_rollout_storage_matches_reference_golden(cuda_device):
    from massive_marl_benchmark_b200.storage import RolloutStorage
    g = load_golden("storage_ppo")
    dev = cuda_device
    T, N = g["in_rew"].shape
    st = RolloutStorage(N, T, (60,), (0,), (8,), dev, "sequential")
    for t in range(T):
        st.add_transitions(g["in_obs"][t].to(dev), torch.zeros(N, 0, device=dev), g["in_act"][t].to(dev),
                           g["in_rew"][t].to(dev), g["in_done"][t].to(dev), g["in_val"][t].to(dev),
                           g["in_logp"][t].to(dev), g["in_mu"][t].to(dev), g["in_sig"][t].to(dev))
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):
        st.add_transitions(g["in_obs"][0].to(dev), torch.zeros(N, 0, device=dev), g["in_act"][0].to(dev),
                           g["in_rew"][0].to(dev), g["in_done"][0].to(dev), g["in_val"][0].to(dev),
                           g["in_logp"][0].to(dev), g["in_mu"][0].to(dev), g["in_sig"][0].to(dev))
    assert torch.equal(st.observations.cpu(), g["in_obs"]) and torch.equal(st.actions.cpu(), g["in_act"])
    assert torch.equal(st.dones.cpu(), g["dones_u8"]) and st.dones.dtype == torch.uint8
    assert torch.equal(st.rewards.cpu()[..., 0], g["in_rew"]) and torch.equal(st.mu.cpu(), g["in_mu"])
    mean_len, mean_rew = st.get_statistics()
    assert abs(float(mean_len) - float(g["mean_len"])) <= 1e-5 * float(g["mean_len"])
    assert abs(float(mean_rew) - float(g["mean_rew"])) <= 1e-6 + 1e-5 * abs(float(g["mean_rew"]))
    st.compute_returns(g["last_values"].to(dev), 0.96, 0.95)
    # the scan is elementwise-sequential: returns are bit-identical to the reference
    assert torch.equal(st.returns.cpu(), g["returns"])
    assert_close_obs(st.advantages, g["advantages"], rtol=1e-5, atol=1e-6, what="normalised advantages")
    # minibatch partition: sequential sampler, drop_last
    parts = [b for b in st.mini_batch_generator(4)]
    assert [len(b) for b in parts] == g["part4_sizes"].tolist()
    assert torch.equal(torch.cat(parts).cpu(), torch.arange(T * N))
    parts3 = [b for b in st.mini_batch_generator(3)]
    assert [len(b) for b in parts3] == g["part3_sizes"].tolist() and int(parts3[-1][-1]) == int(g["part3_last"])
    # what the unmodified PPO.update does with an index item
    obs_batch = st.observations.view(-1, 60)[parts[1]]
    assert torch.equal(obs_batch.cpu(), g["in_obs"].view(-1, 60)[parts[1].cpu()])
    st.clear()
    assert st.step == 0


def test_gae_large_and_random_sampler(cuda_device):
    from oracle import storage_oracle as so
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    gen = torch.Generator().manual_seed(5)
    for T, N in ((16, 4096), (3, 1), (17, 1000), (1, 77)):
        st = RolloutStorage(N, T, (4,), (0,), (2,), dev, "random")
        r = torch.randn(T, N, 1, generator=gen); v = torch.randn(T, N, 1, generator=gen)
        d = (torch.rand(T, N, 1, generator=gen) < 0.05).to(torch.uint8); lv = torch.randn(N, 1, generator=gen)
        st.rewards.copy_(r); st.values.copy_(v); st.dones.copy_(d)
        st.compute_returns(lv.to(dev), 0.96, 0.95)
        ret, adv = so.ppo_compute_returns(r, v, d, lv, 0.96, 0.95)
        assert torch.equal(st.returns.cpu(), ret), (T, N)
        if T * N > 1:
            assert_close_obs(st.advantages, adv, rtol=1e-5, atol=1e-6, what="adv T=%d N=%d" % (T, N))
        # float64 recursion within 1e-5 (property test of SURVEY.md section 4)
        adv64, ret64 = torch.zeros(N, 1, dtype=torch.float64), torch.zeros(T, N, 1, dtype=torch.float64)
        for t in reversed(range(T)):
            nv = lv.double() if t == T - 1 else v[t + 1].double()
            m = 1.0 - d[t].double()
            adv64 = r[t].double() + m * 0.96 * nv - v[t].double() + m * 0.96 * 0.95 * adv64
            ret64[t] = adv64 + v[t].double()
        assert_close_obs(st.returns, ret64, rtol=1e-5, atol=1e-5, what="returns vs float64")
        if T * N < 4:
            continue
        # random sampler: every epoch a fresh permutation of [0, T*N) cut into equal minibatches
        it = st.mini_batch_generator(4)
        e1 = torch.cat([b for b in it]).cpu(); e2 = torch.cat([b for b in it]).cpu()
        mb = (T * N) // 4
        assert e1.numel() == 4 * mb and len(set(e1.tolist())) == e1.numel() and int(e1.max()) < T * N and int(e1.min()) >= 0
        if T * N > 64:
            assert not torch.equal(e1, e2)
        # host-supplied permutation (parity mode): identical minibatches to the oracle partition
        perm = torch.randperm(T * N, generator=gen)
        st.permutation_override = perm
        got = [b.cpu().tolist() for b in st.mini_batch_generator(4)]
        assert got == so.ppo_minibatch_partition(N, T, 4, perm=perm)
        st.permutation_override = None
        # fused gather == per-field indexing
        idx = torch.randperm(T * N, generator=gen)[: max(1, mb)].to(dev)
        out = st.gather_minibatch(idx)
        for f in ("observations", "actions", "values", "returns", "advantages", "mu", "sigma", "actions_log_prob"):
            src = getattr(st, f)
            assert torch.equal(out[f], src.view(-1, *src.shape[2:])[idx]), f


def test_permutation_is_bijection(cuda_device):
    from massive_marl_benchmark_b200 import _lib as L
    dev = cuda_device
    for n in (1, 2, 3, 17, 1000, 65536, 1_000_003):
        out = torch.empty(n, device=dev, dtype=torch.int64)
        L.check(L.lib().mmb_permutation(n, 12345 + n, 1, L.ptr(out), L.stream_ptr()), "perm")
        s = torch.sort(out).values
        assert torch.equal(s, torch.arange(n, device=dev)), n
    a = torch.empty(4096, device=dev, dtype=torch.int64); b = torch.empty_like(a)
    L.lib().mmb_permutation(4096, 1, 1, L.ptr(a), L.stream_ptr()); L.lib().mmb_permutation(4096, 2, 1, L.ptr(b), L.stream_ptr())
    assert not torch.equal(a, b)
    assert float((a == torch.arange(4096, device=dev)).float().mean()) < 0.01


@pytest.mark.parametrize("group", [1, 2, 4, 8, 16])
def test_grouped_shuffle_is_a_permutation_and_matches_the_fused_gather(cuda_device, group):
    """Grouped shuffle (mmb_gather_params.group): a bijection on [0, T*N) whose aligned groups of `group` positions hold one
    aligned source group, rotated; the index-free fused path (`gather_epoch_minibatch`: shuffle + gather in one launch) returns
    exactly the rows the materialised order selects, for every field width (1552 / 320 / 4-byte planes) and minibatch."""
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200.storage import RolloutStorage
    dev = cuda_device
    for n in (group * 1, group * 37, 65536, group * 62501):
        out = torch.empty(n, device=dev, dtype=torch.int64)
        L.check(L.lib().mmb_permutation(n, 999 + n, group, L.ptr(out), L.stream_ptr()), "perm")
        assert torch.equal(torch.sort(out).values, torch.arange(n, device=dev)), n
        grp = out.view(-1, group)
        assert torch.equal(grp // group, (grp[:, :1] // group).expand_as(grp)), "positions of a group read one source group"
        rot = (grp - grp[:, :1]) % group
        assert torch.equal(rot, torch.arange(group, device=dev).expand_as(rot)), "rows of a group keep their cyclic order"
    assert L.lib().mmb_permutation(group * 3 + 1, 1, group, L.ptr(out), L.stream_ptr()) == (0 if group == 1 else -1)
    T, N = 5, 96
    st = RolloutStorage(N, T, (388,), (0,), (80,), dev, "random")
    for f in st.FIELDS:
        t = getattr(st, f)
        if t.numel():
            t.copy_(torch.randn(t.shape, device=dev))
    st.shuffle_group, st.shuffle_seed = group, 7
    for n_mb in (1, 3, 5):
        st._epoch = 3
        order = st._epoch_order()                 # epoch 3's materialised order
        st._epoch = 3
        st.new_epoch()
        mb = T * N // n_mb
        for k in range(n_mb):
            got = st.gather_epoch_minibatch(k, n_mb, want_indices=True)
            idx = order[k * mb:(k + 1) * mb]
            assert torch.equal(got["indices"], idx)
            want = st.gather_minibatch(idx)
            for f, v in want.items():
                assert torch.equal(got[f], v), (f, n_mb, k)
                src = getattr(st, f)
                assert torch.equal(v, src.reshape(-1, *src.shape[2:])[idx]), f


def test_separated_buffer_matches_reference_golden(cuda_device):
    from massive_marl_benchmark_b200 import spaces
    from massive_marl_benchmark_b200.separated_buffer import SeparatedReplayBuffer, runner_insert_masks
    g = load_golden("buffer_marl")
    dev = cuda_device
    T, N = g["in_rewards"].shape[:2]
    cfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=16, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
               use_gae=True, use_popart=True, use_valuenorm=False, use_proper_time_limits=False)
    ob = spaces.Box(low=-np.inf, high=np.inf, shape=(46,)); sh = spaces.Box(low=-np.inf, high=np.inf, shape=(388,))
    ac = spaces.Box(low=-np.ones(8), high=np.ones(8))

    class Norm:  # PopArt stand-in exposing running_mean_var() (the update rule stays PyTorch, SURVEY #15)
        def running_mean_var(self):
            return g["popart_mean"].to(dev), g["popart_var"].to(dev)

    buf = SeparatedReplayBuffer(cfg, ob, sh, ac, dev)
    rs = torch.zeros(N, 1, 16, device=dev)
    for t in range(T):
        buf.insert(g["in_share_obs"][t].to(dev), g["in_obs"][t].to(dev), rs, rs, g["in_actions"][t].to(dev),
                   g["in_logp"][t].to(dev), g["in_value_preds"][t].to(dev), g["in_rewards"][t].to(dev),
                   g["in_masks"][t].to(dev), None, g["in_active_masks"][t].to(dev), None)
    assert buf.step == 0
    assert torch.equal(buf.share_obs.cpu(), g["buf_share_obs"]) and torch.equal(buf.obs.cpu(), g["buf_obs"])
    assert torch.equal(buf.masks.cpu(), g["buf_masks"]) and torch.equal(buf.active_masks.cpu(), g["buf_active_masks"])
    buf.compute_returns(g["next_value"].to(dev), Norm())
    assert torch.equal(buf.value_preds.cpu(), g["value_preds_after"])
    assert torch.equal(buf.returns.cpu(), g["returns"])       # sequential scan: bit-identical
    adv = buf.normalized_advantages(1e-5)
    assert_close_obs(adv, g["advantages"], rtol=1e-5, atol=1e-6, what="MARL advantages")
    # other branches of compute_returns
    for tag, kw, norm in (("plain", dict(use_popart=False, use_valuenorm=False), None),
                          ("ptl", dict(use_proper_time_limits=True), Norm())):
        b2 = SeparatedReplayBuffer(dict(cfg, **kw), ob, sh, ac, dev)
        b2.rewards.copy_(buf.rewards); b2.masks.copy_(buf.masks)
        b2.value_preds.copy_(g["value_preds_after"].to(dev)); b2.bad_masks.copy_(g["bad_masks_" + tag].to(dev))
        b2.value_preds[:-1].copy_(torch.stack([g["in_value_preds"][t] for t in range(T)]).to(dev))
        b2.compute_returns(g["next_value"].to(dev), norm)
        assert torch.equal(b2.returns.cpu(), g["returns_" + tag]), tag
    # feed_forward_generator: 13-tuple, host permutation -> identical rows to direct indexing
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(1))
    buf.permutation_override = perm
    batches = list(buf.feed_forward_generator(adv, num_mini_batch=2))
    assert len(batches) == 2 and len(batches[0]) == 13
    mb = T * N // 2
    for i, tup in enumerate(batches):
        idx = perm[i * mb:(i + 1) * mb].to(dev)
        assert torch.equal(tup[0], buf.share_obs[:-1].reshape(-1, 388)[idx])
        assert torch.equal(tup[1], buf.obs[:-1].reshape(-1, 46)[idx])
        assert torch.equal(tup[4], buf.actions.reshape(-1, 8)[idx])
        assert torch.equal(tup[6], buf.returns[:-1].reshape(-1, 1)[idx])
        assert torch.equal(tup[9], buf.action_log_probs.reshape(-1, 8)[idx])
        assert torch.equal(tup[10], adv.reshape(-1, 1)[idx]) and tup[11] is None
        assert torch.equal(tup[12], buf.factor.reshape(-1, 1)[idx])
    buf.after_update()
    assert torch.equal(buf.share_obs[0], buf.share_obs[-1]) and torch.equal(buf.masks[0], buf.masks[-1])
    # Runner.insert mask logic
    masks, active = runner_insert_masks(g["runner_dones"].to(dev))
    assert torch.equal(masks.cpu(), g["runner_masks"]) and torch.equal(active.cpu(), g["runner_active_masks"])


@pytest.mark.parametrize("N,T,p_done", [(257, 1, 0.3), (1000, 16, 0.02), (4096, 16, 0.2), (50, 7, 0.0)])
def test_episode_tracker_matches_runner_bookkeeping(cuda_device, N, T, p_done):
    """EpisodeTracker == the reference runner's per-step bookkeeping (ppo.py:143-157): running sums bit-exact, the
    deque contents (order included) bit-exact, logged means within fp32 rounding; several updates so that the rings
    wrap, with uint8 and int64 done planes."""
    from massive_marl_benchmark_b200.episodes import EpisodeTracker
    from oracle.episode_oracle import EpisodeOracle
    dev = cuda_device
    gen = torch.Generator().manual_seed(N + T)
    tr, orc = EpisodeTracker(N, dev), EpisodeOracle(N)
    for it in range(5):
        rew = torch.randn(T, N, generator=gen)
        done = (torch.rand(T, N, generator=gen) < p_done)
        if it == 2 and p_done > 0:
            done[T - 1] = True                      # every env finishes at once: more than `window` entries in one row
        d = done.to(torch.uint8) if it % 2 == 0 else done.to(torch.int64) * 3
        orc.update(rew, d)
        tr.update(rew.to(dev).view(T, N, 1), d.to(dev).view(T, N, 1))
        torch.cuda.synchronize()
        assert torch.equal(tr.cur_reward_sum.cpu(), orc.cur_reward_sum)
        assert torch.equal(tr.cur_episode_length.cpu(), orc.cur_episode_length)
        assert int(tr.finished) == orc.finished
        rr, ll = tr.deques()
        assert rr == [float(x) for x in orc.rewbuffer] and ll == [float(x) for x in orc.lenbuffer]
        if orc.finished:
            mr, ml = tr.means()
            er, el = orc.means()
            assert abs(float(mr) - er) <= 1e-5 * max(1.0, abs(er)) and abs(float(ml) - el) <= 1e-5 * max(1.0, abs(el))


@pytest.mark.parametrize("use_popart,ptl", [(True, False), (False, False), (True, True)])
def test_shared_buffer_equals_per_agent_buffers(cuda_device, use_popart, ptl):
    """SharedReplayBuffer (share_obs once, agent-major planes, one GAE launch for all agents) against A independent
    SeparatedReplayBuffers (the reference's arrangement, itself pinned to the golden vectors above) fed the same
    rollout: every tensor an agent reads, the returns, the advantages and the generator tuples are bit-identical."""
    from massive_marl_benchmark_b200 import spaces
    from massive_marl_benchmark_b200.separated_buffer import SeparatedReplayBuffer
    from massive_marl_benchmark_b200.shared_buffer import SharedReplayBuffer
    dev = cuda_device
    T, N, A, O, S, ACT = 6, 50, 4, 46, 388, 8
    cfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=16, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
               use_gae=True, use_popart=use_popart, use_valuenorm=False, use_proper_time_limits=ptl)
    ob = spaces.Box(low=-np.inf, high=np.inf, shape=(O,)); sh = spaces.Box(low=-np.inf, high=np.inf, shape=(S,))
    ac = spaces.Box(low=-np.ones(ACT), high=np.ones(ACT))
    gen = torch.Generator().manual_seed(11)

    class Norm:
        def __init__(self, m, v):
            self.m, self.v = torch.tensor([m], device=dev), torch.tensor([v], device=dev)

        def running_mean_var(self):
            return self.m, self.v

    norms = [Norm(0.1 * i, 1.5 + 0.2 * i) for i in range(A)] if use_popart else None
    shared = SharedReplayBuffer(cfg, A, ob, sh, ac, dev)
    per = [SeparatedReplayBuffer(cfg, ob, sh, ac, dev) for _ in range(A)]
    rs = torch.zeros(N, 1, 16, device=dev)
    for t in range(T):
        share = torch.randn(N, S, generator=gen).to(dev)
        obs = torch.randn(N, A, O, generator=gen).to(dev); act = torch.randn(N, A, ACT, generator=gen).to(dev)
        logp = torch.randn(N, A, ACT, generator=gen).to(dev); val = torch.randn(N, A, 1, generator=gen).to(dev)
        rew = torch.randn(N, A, 1, generator=gen).to(dev)
        masks = (torch.rand(N, 1, 1, generator=gen) > 0.1).float().expand(N, A, 1).contiguous().to(dev)
        bad = (torch.rand(N, A, 1, generator=gen) > 0.05).float().to(dev)
        active = (torch.rand(N, A, 1, generator=gen) > 0.1).float().to(dev)
        shared.insert(share, obs, act, logp, val, rew, masks, bad, active)
        for i in range(A):
            per[i].insert(share, obs[:, i], rs, rs, act[:, i], logp[:, i], val[:, i], rew[:, i], masks[:, i], bad[:, i], active[:, i])
    assert shared.step == 0
    nv = torch.randn(N, A, 1, generator=gen).to(dev)
    shared.compute_returns(nv, norms)
    adv_all = shared.normalized_advantages(1e-5)
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(2))
    for i in range(A):
        per[i].compute_returns(nv[:, i].contiguous(), norms[i] if norms else None)
        view = shared.agent(i)
        for name in ("share_obs", "obs", "value_preds", "returns", "masks", "bad_masks", "active_masks", "actions",
                     "action_log_probs", "rewards", "raw_advantages"):
            assert torch.equal(getattr(view, name), getattr(per[i], name)), (i, name)
        adv_i = per[i].normalized_advantages(1e-5)
        assert torch.equal(adv_all[i], adv_i)
        view.permutation_override = per[i].permutation_override = perm
        for ta, tb in zip(view.feed_forward_generator(adv_all[i], num_mini_batch=2), per[i].feed_forward_generator(adv_i, num_mini_batch=2)):
            assert len(ta) == len(tb) == 13
            for x, y in zip(ta, tb):
                assert (x is None and y is None) or torch.equal(x, y)
    # the view can also run the per-agent path by itself (what an unmodified trainer does) and stays consistent
    v0 = shared.agent(0)
    v0.compute_returns(nv[:, 0].contiguous(), norms[0] if norms else None)
    assert torch.equal(v0.returns, per[0].returns)
    shared.after_update()
    assert torch.equal(shared.share_obs[0], shared.share_obs[-1]) and torch.equal(shared.obs[:, 0], shared.obs[:, -1])
    assert shared.share_obs.data_ptr() == shared.agent(3).share_obs.data_ptr()       # stored once


def test_episode_tracker_marl_runner_variant(cuda_device):
    """update_marl == the bookkeeping lines of the MARL runner (runner.py:135-144) transcribed literally."""
    from massive_marl_benchmark_b200.episodes import EpisodeTracker
    dev = cuda_device
    N, A, T = 123, 4, 9
    gen = torch.Generator().manual_seed(2)
    tr = EpisodeTracker(N, dev)
    train_episode_rewards = torch.zeros(1, N, device=dev)          # the runner keeps it on the training device
    done_episodes_rewards = []
    for it in range(4):
        rewards = torch.randn(T, N, A, 1, generator=gen).to(dev)
        dones = (torch.rand(T, N, 1, generator=gen) < 0.1).expand(T, N, A).clone().to(dev)
        dones[:, ::7, 0] = False                                   # some agents not done -> env not done
        for step in range(T):
            dones_env = torch.all(dones[step], dim=1)                            # runner.py:135
            reward_env = torch.mean(rewards[step], dim=1).flatten()              # runner.py:137
            train_episode_rewards += reward_env                                  # runner.py:139
            for t in range(N):                                                   # runner.py:141-144
                if dones_env[t]:
                    done_episodes_rewards.append(train_episode_rewards[:, t].clone())
                    train_episode_rewards[:, t] = 0
        tr.update_marl(rewards, dones)
        torch.cuda.synchronize()
        assert torch.equal(tr.cur_reward_sum, train_episode_rewards[0])
        assert int(tr.finished) == len(done_episodes_rewards)
        rr, _ = tr.deques()
        assert rr == [float(x) for x in done_episodes_rewards[-100:]]


def test_shared_insert_with_strided_sources(cuda_device):
    """`SharedReplayBuffer.insert` through `mmb_copy_group` with the sources as the rollout actually hands them over: the
    env's reward / done planes as stride-0 expand views, the team forward's agent-major outputs as transposed views, a
    column slice as share_obs - against plain torch copies of the same tensors."""
    from massive_marl_benchmark_b200 import spaces
    from massive_marl_benchmark_b200.shared_buffer import SharedReplayBuffer
    dev = cuda_device
    T, N, A, O, S, ACT = 3, 70, 5, 46, 388, 8
    cfg = dict(episode_length=T, n_rollout_threads=N, hidden_size=16, recurrent_N=1, gamma=0.96, gae_lambda=0.95,
               use_gae=True, use_popart=False, use_valuenorm=False, use_proper_time_limits=False)
    ob = spaces.Box(low=-np.inf, high=np.inf, shape=(O,)); sh = spaces.Box(low=-np.inf, high=np.inf, shape=(S,))
    ac = spaces.Box(low=-np.ones(ACT), high=np.ones(ACT))
    buf = SharedReplayBuffer(cfg, A, ob, sh, ac, dev)
    gen = torch.Generator().manual_seed(5)
    for t in range(T):
        state_all = torch.randn(N, S, generator=gen).to(dev).unsqueeze(1).expand(N, A, S)        # MultiVecTaskPython's state_all
        obs = torch.randn(N, A, O, generator=gen).to(dev)
        act_am = torch.randn(A, N, ACT, generator=gen).to(dev); logp_am = torch.randn(A, N, ACT, generator=gen).to(dev)
        val_am = torch.randn(A, N, 1, generator=gen).to(dev)
        rew = torch.randn(N, generator=gen).to(dev).view(N, 1, 1).expand(N, A, 1)
        masks = (torch.rand(N, generator=gen) > 0.2).float().to(dev).view(N, 1, 1).expand(N, A, 1)
        active = torch.ones(N, A, 1, device=dev)
        s = buf.step
        buf.insert(state_all[:, 0], obs, act_am.transpose(0, 1), logp_am.transpose(0, 1), val_am.transpose(0, 1), rew, masks, None, active)
        assert torch.equal(buf.share_obs[s + 1], state_all[:, 0]) and torch.equal(buf.obs[:, s + 1], obs.transpose(0, 1))
        assert torch.equal(buf.actions[:, s], act_am) and torch.equal(buf.action_log_probs[:, s], logp_am)
        assert torch.equal(buf.value_preds[:, s], val_am) and torch.equal(buf.rewards[:, s], rew.transpose(0, 1))
        assert torch.equal(buf.masks[:, s + 1], masks.transpose(0, 1)) and torch.equal(buf.active_masks[:, s + 1], active.transpose(0, 1))
    assert buf.step == 0

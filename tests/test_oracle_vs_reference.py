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


def test_episode_bookkeeping_oracle_against_the_reference_lines():
    """The episode bookkeeping lives inline in PPO.run() (ppo.py:143-157) and cannot be called on its own, so the
    reference's own source lines are executed here (textually extracted, dedented) and the oracle must reproduce
    the deques and running sums exactly."""
    import statistics
    import textwrap
    from collections import deque
    from oracle.episode_oracle import EpisodeOracle
    src = open(os.path.join(REF, "agents/algorithms/rl/ppo/ppo.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.strip() == "cur_reward_sum[:] += rews")
    end = next(i for i in range(start, len(src)) if src[i].strip() == "cur_episode_length[new_ids] = 0")
    body = compile(textwrap.dedent("\n".join(src[start:end + 1])), "ppo.py:%d-%d" % (start + 1, end + 1), "exec")
    N, T = 37, 9
    gen = torch.Generator().manual_seed(3)
    ns = {"cur_reward_sum": torch.zeros(N, dtype=torch.float), "cur_episode_length": torch.zeros(N, dtype=torch.float)}
    rewbuffer, lenbuffer = deque(maxlen=100), deque(maxlen=100)
    orc = EpisodeOracle(N)
    for it in range(6):
        rew = torch.randn(T, N, generator=gen)
        done = (torch.rand(T, N, generator=gen) < 0.15).to(torch.int64)
        ns["reward_sum"], ns["episode_length"] = [], []
        for t in range(T):                                   # the rollout loop of ppo.py:127
            ns["rews"], ns["dones"] = rew[t], done[t]
            exec(body, {}, ns)
        rewbuffer.extend(ns["reward_sum"]); lenbuffer.extend(ns["episode_length"])     # ppo.py:156-157
        orc.update(rew, done)
        assert list(orc.rewbuffer) == list(rewbuffer) and list(orc.lenbuffer) == list(lenbuffer)
        assert torch.equal(orc.cur_reward_sum, ns["cur_reward_sum"]) and torch.equal(orc.cur_episode_length, ns["cur_episode_length"])
        assert orc.means() == (statistics.mean(rewbuffer), statistics.mean(lenbuffer))


def test_ppo_loss_oracle_against_the_reference_lines(shim):
    """The PPO minibatch loss: the reference's own `ActorCritic.evaluate` (imported) and the loss block of `PPO.update`
    (source lines executed, oracle/ref_ppo_loss.py) against oracle/ppo_loss_oracle.py on fresh seeds - losses, KL,
    log-probs and all three gradients identical."""
    from oracle.ppo_loss_oracle import ppo_loss_oracle, synthetic_minibatch
    from oracle.ref_ppo_loss import reference_ppo_loss
    for B, A, seed, clipped, ec, vc, clip in [(57, 8, 11, True, 0.0, 1.0, 0.2), (31, 80, 12, True, 0.01, 2.0, 0.2),
                                             (40, 24, 13, False, 0.0, 0.5, 0.1)]:
        mb = synthetic_minibatch(B, A, seed)
        ref = reference_ppo_loss(REF, mb, clip, vc, ec, clipped)
        mine = ppo_loss_oracle(**mb, clip_param=clip, value_loss_coef=vc, entropy_coef=ec, use_clipped_value_loss=clipped)
        for k in ref:
            assert torch.equal(ref[k], mine[k]), (A, k)


def test_mappo_loss_oracle_against_the_reference_trainer(shim):
    """The MAPPO minibatch losses: the reference's own `MAPPO.ppo_update` with its own ACTLayer / DiagGaussian / PopArt
    (oracle/ref_mappo_loss.py; only the MLP trunks are stubbed) against oracle/mappo_loss_oracle.py on fresh seeds -
    losses, importance weights and all three gradients identical, in every flag combination the trainer has."""
    import itertools
    from oracle.mappo_loss_oracle import mappo_loss_oracle, synthetic_minibatch, without_popart
    from oracle.ref_mappo_loss import reference_mappo_loss
    base = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, huber_delta=1.0, std_x_coef=1.0, std_y_coef=0.5)
    seed = 40
    for huber, clipped, vmask, pmask, popart in itertools.product((True, False), repeat=5):
        seed += 1
        cfg = dict(base, use_huber_loss=huber, use_clipped_value_loss=clipped, use_value_active_masks=vmask,
                   use_policy_active_masks=pmask)
        mb = synthetic_minibatch(48, 8 if seed % 2 else 6, seed, huber_delta=1.0)
        if not popart:
            mb = without_popart(mb)
        ref, mb = reference_mappo_loss(mb, cfg)
        mine = mappo_loss_oracle(**mb, **cfg)
        for k in ref:
            assert torch.equal(ref[k], mine[k]), (cfg, popart, k)


def _fill_rollout(st, ac, T, N, obs_dim, A, seed):
    """A storage as it looks after a rollout + compute_returns: old policy = the network's current one (plus a little
    noise so that ratios leave 1), advantages normalised."""
    g = torch.Generator().manual_seed(seed)
    st.observations.copy_(torch.randn(T, N, obs_dim, generator=g))
    with torch.no_grad():
        mu = ac.actor(st.observations.view(-1, obs_dim)).view(T, N, A) + 0.02 * torch.randn(T, N, A, generator=g)
        val = ac.critic(st.observations.view(-1, obs_dim)).view(T, N, 1)
        std = ac.log_std.exp() * ac.log_std.exp()
        act = mu + std * torch.randn(T, N, A, generator=g)
        logp = (-0.5 * (((act - mu) / std) ** 2).sum(-1) - std.log().sum() - 0.5 * A * 1.8378770664093453).view(T, N, 1)
    st.mu.copy_(mu); st.sigma.copy_(ac.log_std.detach().repeat(T, N, 1)); st.actions.copy_(act)
    st.actions_log_prob.copy_(logp); st.values.copy_(val + 0.1 * torch.randn(T, N, 1, generator=g))
    st.returns.copy_(val + 0.5 * torch.randn(T, N, 1, generator=g))
    adv = st.returns - st.values
    st.advantages.copy_((adv - adv.mean()) / (adv.std() + 1e-8))


def test_ppo_update_oracle_against_the_reference_update(shim):
    """The whole minibatch loop: the reference's own unmodified `PPO.update` on the reference's `ActorCritic` and
    `RolloutStorage` (sequential sampler) against oracle.ppo_loss_oracle.ppo_update_oracle from the same initial weights -
    identical parameters, optimiser step size and returned loss means after 2 epochs x 3 minibatches, with Adam and the
    adaptive-KL schedule as in cfg/ppo/config.yaml."""
    import contextlib
    import copy
    import io
    import types
    import agents.algorithms.rl.ppo as pkg                 # a bare namespace under the shim: give it what ppo.py imports from it
    from agents.algorithms.rl.ppo.module import ActorCritic
    from agents.algorithms.rl.ppo.storage import RolloutStorage
    pkg.RolloutStorage, pkg.ActorCritic = RolloutStorage, ActorCritic
    from agents.algorithms.rl.ppo.ppo import PPO
    from oracle.ppo_loss_oracle import ppo_update_oracle
    T, N, obs_dim, A = 6, 20, 12, 8
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ac_ref = ActorCritic((obs_dim,), (0,), (A,), 0.8, {"pi_hid_sizes": [32, 16], "vf_hid_sizes": [32, 16], "activation": "elu"})
    ac_mine = copy.deepcopy(ac_ref)
    initial = copy.deepcopy(ac_ref.state_dict())
    st = RolloutStorage(N, T, (obs_dim,), (0,), (A,), "cpu", "sequential")
    _fill_rollout(st, ac_ref, T, N, obs_dim, A, seed=70)

    def make(ac):
        return types.SimpleNamespace(storage=st, actor_critic=ac, optimizer=torch.optim.Adam(ac.parameters(), lr=3e-4),
                                     num_mini_batches=3, num_learning_epochs=2, clip_param=0.2, value_loss_coef=2.0,
                                     entropy_coef=0.001, use_clipped_value_loss=True, desired_kl=0.016, schedule="adaptive",
                                     step_size=3e-4, max_grad_norm=1.0, asymmetric=False)

    ref, mine = make(ac_ref), make(ac_mine)
    out_ref = PPO.update(ref)
    out_mine = ppo_update_oracle(mine, [torch.arange(T * N)] * 2)
    assert out_ref == out_mine and ref.step_size == mine.step_size
    for (k, a), b in zip(ac_ref.state_dict().items(), ac_mine.state_dict().values()):
        assert torch.equal(a, b), k
    assert not torch.equal(ac_ref.state_dict()["actor.0.weight"], initial["actor.0.weight"])      # the update did move them


def _marl_config(**over):
    cfg = dict(clip_param=0.2, ppo_epoch=1, num_mini_batch=1, data_chunk_length=None, value_loss_coef=1.0, entropy_coef=0.01,
               max_grad_norm=10.0, huber_delta=10.0, use_valuenorm=False, use_recurrent_policy=False,
               use_naive_recurrent_policy=False, use_max_grad_norm=True, use_clipped_value_loss=True, use_huber_loss=True,
               use_popart=True, use_value_active_masks=False, use_policy_active_masks=False, std_x_coef=1.0, std_y_coef=0.5,
               actor_gain=0.01, gain=0.01, hidden_size=32, use_orthogonal=True, use_feature_normalization=True,
               use_ReLU=True, stacked_frames=1, layer_N=2, recurrent_N=1, lr=5e-4, critic_lr=5e-4, opti_eps=1e-5,
               weight_decay=0.0, algorithm_name="mappo")
    cfg.update(over)
    return cfg


def _marl_sample(policy, B, obs_dim, share_dim, A, seed):
    g = torch.Generator().manual_seed(seed)
    obs, share = torch.randn(B, obs_dim, generator=g), torch.randn(B, share_dim, generator=g)
    with torch.no_grad():
        head = policy.actor.act.action_out
        mean = head.fc_mean(policy.actor.base(obs)) + 0.05 * torch.randn(B, A, generator=g)
        std = torch.sigmoid(head.log_std / head.std_x_coef) * head.std_y_coef
        actions = mean + std * torch.randn(B, A, generator=g)
        old_logp = torch.distributions.Normal(mean, std).log_prob(actions)
        vals = policy.critic.v_out(policy.critic.base(share))
    value_preds = vals + 0.1 * torch.randn(B, 1, generator=g)
    returns = 1.0 + 2.0 * torch.randn(B, 1, generator=g)
    active = (torch.rand(B, 1, generator=g) > 0.2).float()
    adv = torch.randn(B, 1, generator=g)
    return (share, obs, torch.zeros(B, 1, 32), torch.zeros(B, 1, 32), actions, value_preds, returns, torch.ones(B, 1), active,
            old_logp, adv, None, None)


def test_mappo_update_oracle_against_the_reference_trainer(shim):
    """The whole `MAPPO.ppo_update`: the reference's own trainer, `MAPPO_Policy` (Actor / Critic with MLPBase, ACTLayer,
    DiagGaussian), PopArt and Adam optimisers against oracle.mappo_loss_oracle.mappo_update_oracle from the same initial
    state - identical parameters of both networks, PopArt statistics and returned values after three updates."""
    import contextlib
    import copy
    import io
    import types
    from agents.algorithms.marl.mappo_policy import MAPPO_Policy
    from agents.algorithms.marl.mappo_trainer import MAPPO
    from gym import spaces
    from oracle.mappo_loss_oracle import mappo_update_oracle
    obs_dim, share_dim, A, B = 10, 18, 6, 64
    for over in (dict(), dict(use_value_active_masks=True, use_policy_active_masks=True, huber_delta=0.5),
                 dict(use_popart=False, use_huber_loss=False, use_clipped_value_loss=False)):
        cfg = _marl_config(**over)
        torch.manual_seed(3)
        box = lambda n: spaces.Box(low=-1.0, high=1.0, shape=(n,))           # noqa: E731
        with contextlib.redirect_stdout(io.StringIO()):
            policy = MAPPO_Policy(cfg, box(obs_dim), box(share_dim), box(A))
        trainer = MAPPO(cfg, policy)
        mine_policy = types.SimpleNamespace(actor=copy.deepcopy(policy.actor), critic=copy.deepcopy(policy.critic))
        mine_policy.actor_optimizer = torch.optim.Adam(mine_policy.actor.parameters(), lr=cfg["lr"], eps=cfg["opti_eps"])
        mine_policy.critic_optimizer = torch.optim.Adam(mine_policy.critic.parameters(), lr=cfg["critic_lr"], eps=cfg["opti_eps"])
        mine = types.SimpleNamespace(policy=mine_policy, clip_param=cfg["clip_param"], value_loss_coef=cfg["value_loss_coef"],
                                     entropy_coef=cfg["entropy_coef"], max_grad_norm=cfg["max_grad_norm"],
                                     huber_delta=cfg["huber_delta"], _use_popart=cfg["use_popart"],
                                     _use_huber_loss=cfg["use_huber_loss"], _use_clipped_value_loss=cfg["use_clipped_value_loss"],
                                     _use_value_active_masks=cfg["use_value_active_masks"],
                                     _use_policy_active_masks=cfg["use_policy_active_masks"], popart=None)
        if cfg["use_popart"]:
            pa = trainer.value_normalizer
            mine.popart = dict(running_mean=pa.running_mean.clone(), running_mean_sq=pa.running_mean_sq.clone(),
                               debiasing_term=pa.debiasing_term.clone())
        for it in range(3):
            sample = _marl_sample(policy, B, obs_dim, share_dim, A, seed=100 + it)
            ref = trainer.ppo_update(sample)
            out = mappo_update_oracle(mine, sample)
            for a, b in zip(ref, out):
                assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), (over, it)
        for net_ref, net_mine in ((policy.actor, mine_policy.actor), (policy.critic, mine_policy.critic)):
            for (k, a), b in zip(net_ref.state_dict().items(), net_mine.state_dict().values()):
                assert torch.equal(a, b), (over, k)
        if cfg["use_popart"]:
            pa = trainer.value_normalizer
            assert torch.equal(pa.running_mean, mine.popart["running_mean"]) and torch.equal(pa.debiasing_term, mine.popart["debiasing_term"])
            assert float(pa.debiasing_term) > 0


def test_ippo_update_oracle_against_the_reference_trainer(shim):
    """`IPPO.ppo_update` (decentralised critic, ValueNorm as in cfg/ippo/config.yaml): the reference's own IPPO trainer +
    IPPO_Policy + ValueNorm + Adam against mappo_update_oracle(ippo=True) - identical parameters, normaliser statistics and
    returned values after three updates."""
    import contextlib
    import copy
    import io
    import types
    from agents.algorithms.marl.ippo_policy import IPPO_Policy
    from agents.algorithms.marl.ippo_trainer import IPPO
    from gym import spaces
    from oracle.mappo_loss_oracle import mappo_update_oracle
    obs_dim, A, B = 10, 6, 64
    for over in (dict(use_popart=False, use_valuenorm=True), dict(use_popart=False, use_valuenorm=False, use_huber_loss=False)):
        cfg = _marl_config(algorithm_name="ippo", **over)
        torch.manual_seed(5)
        box = lambda n: spaces.Box(low=-1.0, high=1.0, shape=(n,))           # noqa: E731
        with contextlib.redirect_stdout(io.StringIO()):
            policy = IPPO_Policy(cfg, box(obs_dim), box(obs_dim), box(A))     # use_centralized_V False: the critic sees obs
        trainer = IPPO(cfg, policy)
        mine_policy = types.SimpleNamespace(actor=copy.deepcopy(policy.actor), critic=copy.deepcopy(policy.critic))
        mine_policy.actor_optimizer = torch.optim.Adam(mine_policy.actor.parameters(), lr=cfg["lr"], eps=cfg["opti_eps"])
        mine_policy.critic_optimizer = torch.optim.Adam(mine_policy.critic.parameters(), lr=cfg["critic_lr"], eps=cfg["opti_eps"])
        mine = types.SimpleNamespace(policy=mine_policy, clip_param=cfg["clip_param"], value_loss_coef=cfg["value_loss_coef"],
                                     entropy_coef=cfg["entropy_coef"], max_grad_norm=cfg["max_grad_norm"],
                                     huber_delta=cfg["huber_delta"], _use_popart=cfg["use_popart"],
                                     _use_valuenorm=cfg["use_valuenorm"], _use_huber_loss=cfg["use_huber_loss"],
                                     _use_clipped_value_loss=cfg["use_clipped_value_loss"],
                                     _use_value_active_masks=cfg["use_value_active_masks"],
                                     _use_policy_active_masks=cfg["use_policy_active_masks"], popart=None)
        if cfg["use_valuenorm"]:
            vn = trainer.value_normalizer
            mine.popart = dict(running_mean=vn.running_mean.clone(), running_mean_sq=vn.running_mean_sq.clone(),
                               debiasing_term=vn.debiasing_term.clone())
        for it in range(3):
            sample = _marl_sample(policy, B, obs_dim, obs_dim, A, seed=300 + it)
            sample = (sample[1],) + sample[1:]                               # share_obs = obs
            ref = trainer.ppo_update(sample)
            out = mappo_update_oracle(mine, sample, ippo=True)
            for a, b in zip(ref, out):
                assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), (over, it)
        for net_ref, net_mine in ((policy.actor, mine_policy.actor), (policy.critic, mine_policy.critic)):
            for (k, a), b in zip(net_ref.state_dict().items(), net_mine.state_dict().values()):
                assert torch.equal(a, b), (over, k)
        if cfg["use_valuenorm"]:
            vn = trainer.value_normalizer
            assert torch.equal(vn.running_mean, mine.popart["running_mean"]) and float(vn.debiasing_term) > 0


def test_happo_update_oracle_against_the_reference_trainer(shim):
    """`HAPPO.ppo_update` (the factor of the previously updated agents inside the surrogate, PopArt): the reference's own
    HAPPO trainer + HAPPO_Policy against mappo_update_oracle(happo=True)."""
    import contextlib
    import copy
    import io
    import types
    from agents.algorithms.marl.happo_policy import HAPPO_Policy
    from agents.algorithms.marl.happo_trainer import HAPPO
    from gym import spaces
    from oracle.mappo_loss_oracle import mappo_update_oracle
    obs_dim, share_dim, A, B = 10, 18, 6, 64
    for over in (dict(), dict(use_policy_active_masks=True, use_value_active_masks=True, use_popart=False)):
        cfg = _marl_config(algorithm_name="happo", **over)
        torch.manual_seed(8)
        box = lambda n: spaces.Box(low=-1.0, high=1.0, shape=(n,))           # noqa: E731
        with contextlib.redirect_stdout(io.StringIO()):
            policy = HAPPO_Policy(cfg, box(obs_dim), box(share_dim), box(A))
        trainer = HAPPO(cfg, policy)
        mine_policy = types.SimpleNamespace(actor=copy.deepcopy(policy.actor), critic=copy.deepcopy(policy.critic))
        mine_policy.actor_optimizer = torch.optim.Adam(mine_policy.actor.parameters(), lr=cfg["lr"], eps=cfg["opti_eps"])
        mine_policy.critic_optimizer = torch.optim.Adam(mine_policy.critic.parameters(), lr=cfg["critic_lr"], eps=cfg["opti_eps"])
        mine = types.SimpleNamespace(policy=mine_policy, clip_param=cfg["clip_param"], value_loss_coef=cfg["value_loss_coef"],
                                     entropy_coef=cfg["entropy_coef"], max_grad_norm=cfg["max_grad_norm"],
                                     huber_delta=cfg["huber_delta"], _use_popart=cfg["use_popart"], _use_valuenorm=False,
                                     _use_huber_loss=cfg["use_huber_loss"], _use_clipped_value_loss=cfg["use_clipped_value_loss"],
                                     _use_value_active_masks=cfg["use_value_active_masks"],
                                     _use_policy_active_masks=cfg["use_policy_active_masks"], popart=None)
        if cfg["use_popart"]:
            pa = trainer.value_normalizer
            mine.popart = dict(running_mean=pa.running_mean.clone(), running_mean_sq=pa.running_mean_sq.clone(),
                               debiasing_term=pa.debiasing_term.clone())
        for it in range(3):
            sample = _marl_sample(policy, B, obs_dim, share_dim, A, seed=500 + it)
            g = torch.Generator().manual_seed(600 + it)
            factor = torch.exp(0.2 * torch.randn(B, 1, generator=g))                   # runner.py:271,312-313: [T, N, 1], exp(...)
            if it == 2:
                factor = factor.repeat(1, A) / A                                        # (the trainer also takes one column per dim)
            sample = sample[:12] + (factor,)
            ref = trainer.ppo_update(sample)
            out = mappo_update_oracle(mine, sample, happo=True)
            for a, b in zip(ref, out):
                assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), (over, it)
        for net_ref, net_mine in ((policy.actor, mine_policy.actor), (policy.critic, mine_policy.critic)):
            for (k, a), b in zip(net_ref.state_dict().items(), net_mine.state_dict().values()):
                assert torch.equal(a, b), (over, k)


def test_checkpoint_formats_round_trip_with_the_reference_classes(shim, tmp_path):
    """PPO: a `model_{it}.pt` written by the reference's `PPO.save` loads into checkpoints.load_ppo with identical
    forward outputs, and a file written by checkpoints.save_ppo loads into the reference's ActorCritic (strict keys).
    MARL: the shipped TenAnt MAPPO checkpoints (logs/ten_ant/mappo/models_seed-1) load into the reference's Actor / Critic
    and through checkpoints.load_marl_agents with the same tensors."""
    import contextlib
    import io
    import types
    import agents.algorithms.rl.ppo as pkg
    from agents.algorithms.rl.ppo.module import ActorCritic
    from agents.algorithms.rl.ppo.storage import RolloutStorage
    pkg.RolloutStorage, pkg.ActorCritic = RolloutStorage, ActorCritic
    from agents.algorithms.rl.ppo.ppo import PPO
    from massive_marl_benchmark_b200 import checkpoints as ck
    cfg = {"pi_hid_sizes": [48, 24, 12], "vf_hid_sizes": [40, 20], "activation": "elu"}
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ActorCritic((17,), (23,), (5,), 0.7, cfg, asymmetric=True)
    path = str(tmp_path / "model_1200.pt")
    PPO.save(types.SimpleNamespace(actor_critic=ref), path)                       # ppo.py:96-97
    mine, it = ck.load_ppo(path)
    assert it == 1200 and mine.asymmetric
    x, s = torch.randn(9, 17), torch.randn(9, 23)
    assert torch.equal(mine.actor(x), ref.actor(x)) and torch.equal(mine.critic(s), ref.critic(s))
    assert torch.equal(mine.log_std, ref.log_std)
    back = ck.save_ppo(mine, str(tmp_path), 7)
    assert back.endswith("model_7.pt")
    with contextlib.redirect_stdout(io.StringIO()):
        ref2 = ActorCritic((17,), (23,), (5,), 1.0, cfg, asymmetric=True)
    loader = types.SimpleNamespace(actor_critic=ref2, current_learning_iteration=0)
    PPO.load(loader, back)                                                        # ppo.py:90-94
    assert loader.current_learning_iteration == 7 and torch.equal(ref2.actor(x), ref.actor(x))

    model_dir = os.path.join(REF, "logs/ten_ant/mappo/models_seed-1")
    actors, critics = ck.load_marl_agents(model_dir, 10)
    assert len(actors) == len(critics) == 10
    assert actors[3]["act.action_out.fc_mean.weight"].shape == (8, 512) and critics[3]["v_out.weight"].shape == (1, 512)
    assert actors[0]["base.mlp.fc1.0.weight"].shape == (512, 46) and critics[0]["base.mlp.fc1.0.weight"].shape == (512, 388)

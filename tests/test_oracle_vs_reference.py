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
    from oracle.mappo_loss_oracle import mappo_loss_oracle, synthetic_minibatch
    from oracle.ref_mappo_loss import reference_mappo_loss
    base = dict(clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.01, huber_delta=1.0, std_x_coef=1.0, std_y_coef=0.5)
    seed = 40
    for huber, clipped, vmask, pmask, popart in itertools.product((True, False), repeat=5):
        seed += 1
        cfg = dict(base, use_huber_loss=huber, use_clipped_value_loss=clipped, use_value_active_masks=vmask,
                   use_policy_active_masks=pmask)
        mb = synthetic_minibatch(48, 8 if seed % 2 else 6, seed, huber_delta=1.0)
        if not popart:
            mb["ret_mean"] = mb["ret_var"] = None
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

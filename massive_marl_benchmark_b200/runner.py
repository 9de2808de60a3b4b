"""`Runner` with the reference's interface (agents/algorithms/marl/runner.py:24-112: `Runner(vec_env, config, model_dir)`,
`run / warmup / collect / insert / compute / train / save / restore / log_train`) whose per-step path sits on this
library, plus the algorithm dispatch the reference leaves open:

* IPPO.  `agents/utils/process_marl.py:23` only routes mappo / happo / hatrpo to the on-policy runner and
  `runner.py:74-82` has no branch for `ippo`, although the trainer, the policy and `cfg/ippo/config.yaml` ship
  (`ippo_trainer.py`, `ippo_policy.py`).  `ALGORITHMS` below registers all four and `process_MultiAgentRL` is
  process_marl.py:17-38 with the missing branch.
* per env step: the mask logic of `Runner.insert` (runner.py:229-255, ~12 boolean-mask index_put kernels) is one launch
  (`mmb_marl_masks`) writing straight into the buffers' planes; the reward / done bookkeeping (runner.py:135-144: a Python
  loop over all envs with one host sync per env per step) is done once per episode on the device (`EpisodeTracker`);
* buffers are `separated_buffer.SeparatedReplayBuffer` (one-launch insert / compute_returns / gathers);
* updates: the reference's own trainers (their `train()` loop over epochs and minibatches is kept), optionally with this
  library's fused update body (`fused_update=True`: `mappo_update.{mappo,ippo,happo}_ppo_update`);
* the probability-ratio factor of runner.py:266-317 (two extra full-buffer actor evaluations per agent) is only computed for
  the algorithms that use it (happo, hatrpo); mappo / ippo ignore it.

The trainers / policies are the reference's (imported by name from `agents.algorithms.marl`, i.e. this class lives next to
a checkout of the reference, exactly like the other drop-ins).  There is no CPU path.
"""
import importlib
import os
import time

import torch

from . import _lib as L
from .episodes import EpisodeTracker
from .separated_buffer import SeparatedReplayBuffer, runner_insert_masks

# algorithm_name -> (trainer module, trainer class, policy module, policy class) under agents.algorithms.marl
ALGORITHMS = {
    "mappo": ("mappo_trainer", "MAPPO", "mappo_policy", "MAPPO_Policy"),          # runner.py:80-82
    "happo": ("happo_trainer", "HAPPO", "happo_policy", "HAPPO_Policy"),          # runner.py:74-76
    "hatrpo": ("hatrpo_trainer", "HATRPO", "hatrpo_policy", "HATRPO_Policy"),     # runner.py:77-79
    "ippo": ("ippo_trainer", "IPPO", "ippo_policy", "IPPO_Policy"),               # missing in the reference's dispatch
}
USES_FACTOR = ("happo", "hatrpo")


def resolve_algorithm(name, package="agents.algorithms.marl"):
    """(TrainAlgo, Policy) classes of the reference for `algorithm_name`; raises like config.py:26-28 for unknown names."""
    if name not in ALGORITHMS:
        raise Exception("Unrecognized algorithm %r (on-policy MARL runner: %s)" % (name, ", ".join(sorted(ALGORITHMS))))
    tmod, tcls, pmod, pcls = ALGORITHMS[name]
    return (getattr(importlib.import_module(package + "." + tmod), tcls),
            getattr(importlib.import_module(package + "." + pmod), pcls))


def process_MultiAgentRL(args, env, config, model_dir=""):
    """agents/utils/process_marl.py:17-38 with the `ippo` branch the reference lacks."""
    config["n_rollout_threads"] = env.num_envs
    config["n_eval_rollout_threads"] = env.num_envs
    if args.algo in ALGORITHMS:
        return Runner(vec_env=env, config=config, model_dir=model_dir)
    raise Exception("Unrecognized algorithm %r for the on-policy MARL runner" % (args.algo,))


class Runner:
    def __init__(self, vec_env, config, model_dir="", fused_update=False, writer=True, team_forward=False, shared_buffer=False):
        """`team_forward=True`: `collect` evaluates all agents' actors and critics as two grouped tensor-core forwards
        (`mlp.MarlTeamForward`, bf16 operands; kept in step with the policies' live parameters) instead of 2 x num_agents
        torch module calls of ~10 kernels each (runner.py:205-217) - at N = 4096 the per-step cost of the reference's loop is
        those ~300 launches, not the environment.  `shared_buffer=True` (centralised critic only): one `SharedReplayBuffer`
        for the team - `share_obs` stored once per env, one insert for all agents - whose per-agent views the trainers use."""
        L.lib()
        self.envs = vec_env
        self.eval_envs = vec_env
        self.env_name = vec_env.task.cfg["env"]["env_name"]
        self.algorithm_name = config["algorithm_name"]
        self.experiment_name = config["experiment_name"]
        self.use_centralized_V = config["use_centralized_V"]
        self.use_obs_instead_of_state = config["use_obs_instead_of_state"]
        self.num_env_steps = config["num_env_steps"]
        self.episode_length = config["episode_length"]
        self.n_rollout_threads = config["n_rollout_threads"]
        self.n_eval_rollout_threads = config["n_eval_rollout_threads"]
        self.use_linear_lr_decay = config["use_linear_lr_decay"]
        self.hidden_size = config["hidden_size"]
        self.use_render = config["use_render"]
        self.recurrent_N = config["recurrent_N"]
        self.use_single_network = config["use_single_network"]
        self.save_interval = config["save_interval"]
        self.use_eval = config["use_eval"]
        self.eval_interval = config["eval_interval"]
        self.eval_episodes = config["eval_episodes"]
        self.log_interval = config["log_interval"]
        self.seed = self.envs.task.cfg["seed"]
        self.model_dir = model_dir
        self.num_agents = self.envs.num_agents
        self.device = self.envs.rl_device
        if config.get("use_recurrent_policy") or config.get("use_naive_recurrent_policy"):
            raise NotImplementedError("recurrent policies are outside the benchmark's configurations")

        self.run_dir = config["run_dir"]
        base = str(self.run_dir) + "/" + self.env_name + "/" + self.algorithm_name
        self.log_dir = base + "/logs_seed{}".format(self.seed)
        self.save_dir = base + "/models_seed{}".format(self.seed)
        os.makedirs(self.log_dir, exist_ok=True)
        os.makedirs(self.save_dir, exist_ok=True)
        self.writter = None
        if writer:
            from torch.utils.tensorboard import SummaryWriter
            self.writter = SummaryWriter(self.log_dir)

        TrainAlgo, Policy = resolve_algorithm(self.algorithm_name)
        if fused_update:
            from . import mappo_update as mu
            body = {"mappo": mu.mappo_ppo_update, "ippo": mu.ippo_ppo_update, "happo": mu.happo_ppo_update}.get(self.algorithm_name)
            if body is None:
                raise NotImplementedError("no fused update body for %r" % (self.algorithm_name,))
            TrainAlgo = type(TrainAlgo.__name__ + "Fused", (TrainAlgo,), {"ppo_update": body})

        def cent_space(a):
            return self.envs.share_observation_space[a] if self.use_centralized_V else self.envs.observation_space[a]

        self.policy = [Policy(config, self.envs.observation_space[a], cent_space(a), self.envs.action_space[a], device=self.device)
                       for a in range(self.num_agents)]
        if self.model_dir != "":
            self.restore()
        self.trainer, self.buffer, self.shared = [], [], None
        if shared_buffer:
            if not self.use_centralized_V:
                raise ValueError("shared_buffer needs a centralised critic (one share_obs row per env for all agents)")
            from .shared_buffer import SharedReplayBuffer
            self.shared = SharedReplayBuffer(config, self.num_agents, self.envs.observation_space[0], cent_space(0),
                                             self.envs.action_space[0], self.device)
        for a in range(self.num_agents):
            self.trainer.append(TrainAlgo(config, self.policy[a], device=self.device))
            self.buffer.append(self.shared.agent(a) if self.shared is not None else
                               SeparatedReplayBuffer(config, self.envs.observation_space[a], cent_space(a), self.envs.action_space[a],
                                                     self.device))
        self.team = None
        if team_forward:
            from .mlp import MarlTeamForward
            head = self.policy[0].actor.act.action_out
            self.team = MarlTeamForward([p.actor.state_dict() for p in self.policy], [p.critic.state_dict() for p in self.policy],
                                        std_x_coef=float(head.std_x_coef), std_y_coef=float(head.std_y_coef), device=self.device)
        self.episodes = EpisodeTracker(self.n_rollout_threads, self.device)
        self._finished_before = 0
        self._masks = torch.ones(self.n_rollout_threads, self.num_agents, 1, device=self.device)
        self._active = torch.ones_like(self._masks)

    # ------------------------------------------------------------------------------------------------------------
    def run(self):
        self.warmup()
        start = time.time()
        episodes = int(self.num_env_steps) // self.episode_length // self.n_rollout_threads
        for episode in range(episodes):
            if self.use_linear_lr_decay:
                for tr in self.trainer:                                  # (runner.py:125 calls it on the list: a latent bug)
                    tr.policy.lr_decay(episode, episodes)
            for step in range(self.episode_length):
                values, actions, action_log_probs, rnn_states, rnn_states_critic = self.collect(step)
                obs, share_obs, rewards, dones, infos, _ = self.envs.step(actions)
                self.insert((obs, share_obs, rewards, dones, infos, values, actions, action_log_probs, rnn_states, rnn_states_critic))
            aver = self._episode_bookkeeping()
            self.compute()
            train_infos = self.train()
            total_num_steps = (episode + 1) * self.episode_length * self.n_rollout_threads
            if episode % self.save_interval == 0 or episode == episodes - 1:
                self.save()
            if episode % self.log_interval == 0:
                end = time.time()
                print("\nAlgo {} Exp {} updates {}/{} episodes, total num timesteps {}/{}, FPS {}.\n".format(
                    self.algorithm_name, self.experiment_name, episode, episodes, total_num_steps, self.num_env_steps,
                    int(total_num_steps / (end - start))))
                self.log_train(train_infos, total_num_steps)
            if aver is not None:
                print("some episodes done, average rewards: ", aver)
                if self.writter is not None:
                    self.writter.add_scalars("train_episode_rewards", {"aver_rewards": aver}, total_num_steps)
            if episode % self.eval_interval == 0 and self.use_eval:
                raise NotImplementedError("evaluation rollouts (runner.py:351-409) are outside the hot path; use the reference's eval")

    def _episode_bookkeeping(self):
        """runner.py:135-144 for the T steps just collected, on the device: reward_env = mean over agents, an episode ends
        where all agents are done (= the buffers' masks), finished episodes' reward sums are averaged.  One host sync per
        rollout (the reference: one per env per step)."""
        T = self.episode_length
        reward_env = torch.stack([b.rewards[:, :, 0] for b in self.buffer], 0).mean(0)          # [T, N]
        dones_env = (self.buffer[0].masks[1:T + 1, :, 0] == 0).to(torch.uint8)                   # masks = 1 - dones_env
        self.episodes.update(reward_env, dones_env)
        finished = int(self.episodes.finished.item())
        new, self._finished_before = finished - self._finished_before, finished
        if new == 0:
            return None
        ep = self.episodes._scratch[0]                                          # finished-episode sums, valid where done
        d = dones_env.bool()
        return torch.where(d, ep, torch.zeros_like(ep)).sum() / d.sum()

    def warmup(self):
        obs, share_obs, _ = self.envs.reset()
        if not self.use_centralized_V:
            share_obs = obs
        if self.shared is not None:
            self.shared.share_obs[0].copy_(share_obs[:, 0])
            self.shared.obs[:, 0].copy_(obs.transpose(0, 1))
            return
        for a in range(self.num_agents):
            self.buffer[a].share_obs[0].copy_(share_obs[:, a])
            self.buffer[a].obs[0].copy_(obs[:, a])

    @torch.no_grad()
    def collect(self, step):
        if self.team is not None:
            # two grouped forwards for the whole team; the rnn states of a feed-forward policy are the zeros it was given
            A = self.num_agents
            obs = [self.buffer[a].obs[step] for a in range(A)]
            share = self.buffer[0].share_obs[step] if (self.use_centralized_V and self.shared is not None) else \
                [self.buffer[a].share_obs[step] for a in range(A)]
            if isinstance(share, list):
                share = torch.stack(share)
            values, actions, logps = self.team.get_actions(share, obs)           # [A, N, 1], [A, N, act], [A, N, act]
            self._team_out = (actions, logps)                                    # agent-major: what the shared insert wants
            b0 = self.buffer[0]
            rnn = b0.rnn_states[step].unsqueeze(1).expand(-1, A, -1, -1)
            return values.transpose(0, 1), list(actions.unbind(0)), list(logps.unbind(0)), rnn, rnn
        values, actions, logps, rnn, rnn_c = [], [], [], [], []
        for a in range(self.num_agents):
            self.trainer[a].prep_rollout()
            b = self.buffer[a]
            value, action, logp, rs, rsc = self.trainer[a].policy.get_actions(b.share_obs[step], b.obs[step], b.rnn_states[step],
                                                                               b.rnn_states_critic[step], b.masks[step])
            values.append(value.detach()); actions.append(action.detach()); logps.append(logp.detach())
            rnn.append(rs.detach()); rnn_c.append(rsc.detach())
        return (torch.transpose(torch.stack(values), 1, 0), actions, logps, torch.transpose(torch.stack(rnn), 1, 0),
                torch.transpose(torch.stack(rnn_c), 1, 0))

    def insert(self, data):
        obs, share_obs, rewards, dones, infos, values, actions, action_log_probs, rnn_states, rnn_states_critic = data
        runner_insert_masks(dones, self._masks, self._active)             # runner.py:232-241 in one launch
        if not self.use_centralized_V:
            share_obs = obs
        if self.shared is not None:                                       # the whole team in one insert; share_obs once per env
            team_out = self.__dict__.pop("_team_out", None)
            if team_out is not None:                                      # (N, A, .) views of the agent-major forward outputs
                acts, logps = team_out[0].transpose(0, 1), team_out[1].transpose(0, 1)
            else:
                stack = lambda xs: xs if torch.is_tensor(xs) else torch.stack(tuple(xs), dim=1)  # noqa: E731  per-agent list -> (N, A, .)
                acts, logps = stack(actions), stack(action_log_probs)
            self.shared.insert(share_obs[:, 0], obs, acts, logps, values, rewards, self._masks, None, self._active)
            for a in range(self.num_agents):
                self.buffer[a].step = self.shared.step
            return
        for a in range(self.num_agents):
            # feed-forward policies return the rnn states they were given (zeros): the planes stay as allocated
            self.buffer[a].insert(share_obs[:, a], obs[:, a], None, None, actions[a], action_log_probs[a], values[:, a],
                                  rewards[:, a], self._masks[:, a], None, self._active[:, a], None)

    @torch.no_grad()
    def compute(self):
        if self.team is not None and self.shared is not None:             # bootstrap values of all agents: one grouped forward,
            nv = self.team.get_values(self.shared.share_obs[-1])          # returns of all agents: one launch
            self.shared.compute_returns(nv.transpose(0, 1), [t.value_normalizer for t in self.trainer]
                                        if self.trainer[0].value_normalizer is not None else None)
            return
        for a in range(self.num_agents):
            self.trainer[a].prep_rollout()
            b = self.buffer[a]
            next_value = self.trainer[a].policy.get_values(b.share_obs[-1], b.rnn_states_critic[-1], b.masks[-1]).detach()
            b.compute_returns(next_value, self.trainer[a].value_normalizer)

    def train(self):
        train_infos = []
        T, N = self.episode_length, self.n_rollout_threads
        factor = torch.ones(T, N, 1, device=self.device)
        uses_factor = self.algorithm_name in USES_FACTOR
        for a in torch.randperm(self.num_agents).tolist():                # runner.py:266: random update order
            b, tr = self.buffer[a], self.trainer[a]
            tr.prep_training()
            b.update_factor(factor)
            if uses_factor:
                old_logp = self._evaluate(a)
            train_infos.append(tr.train(b))
            if uses_factor:                                               # runner.py:312-313
                new_logp = self._evaluate(a)
                act_dim = b.actions.shape[-1]
                factor = factor * torch.exp((new_logp - old_logp).reshape(T, N, act_dim).sum(dim=-1, keepdim=True)).detach()
            if self.shared is None:
                b.after_update()
        if self.shared is not None:
            self.shared.after_update()
        return train_infos

    @torch.no_grad()
    def _evaluate(self, a):
        b, actor = self.buffer[a], self.trainer[a].policy.actor
        flat = lambda t: t.reshape(-1, *t.shape[2:])                      # noqa: E731
        out = actor.evaluate_actions(flat(b.obs[:-1]), flat(b.rnn_states[0:1]), flat(b.actions), flat(b.masks[:-1]), None,
                                     flat(b.active_masks[:-1]))
        return out[0].detach()

    # ------------------------------------------------------------------------------------------------------------
    def save(self):                                                       # runner.py:319-328 (formats kept: checkpoints.py)
        for a in range(self.num_agents):
            if self.use_single_network:
                torch.save(self.trainer[a].policy.model.state_dict(), str(self.save_dir) + "/model_agent" + str(a) + ".pt")
            else:
                torch.save(self.trainer[a].policy.actor.state_dict(), str(self.save_dir) + "/actor_agent" + str(a) + ".pt")
                torch.save(self.trainer[a].policy.critic.state_dict(), str(self.save_dir) + "/critic_agent" + str(a) + ".pt")

    def restore(self):                                                    # runner.py:330-339
        for a in range(self.num_agents):
            if self.use_single_network:
                self.policy[a].model.load_state_dict(torch.load(str(self.model_dir) + "/model_agent" + str(a) + ".pt"))
            else:
                self.policy[a].actor.load_state_dict(torch.load(str(self.model_dir) + "/actor_agent" + str(a) + ".pt"))
                self.policy[a].critic.load_state_dict(torch.load(str(self.model_dir) + "/critic_agent" + str(a) + ".pt"))

    def log_train(self, train_infos, total_num_steps):                    # runner.py:341-345
        if self.writter is None:
            return
        for a in range(self.num_agents):
            for k, v in train_infos[a].items():
                agent_k = "agent%i/" % a + k
                self.writter.add_scalars(agent_k, {agent_k: v}, total_num_steps)

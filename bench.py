#!/usr/bin/env python
"""bench.py - env-steps/sec of the rollout hot path (obs + reward + reset + GAE) at TenAnt 4096 x 10.

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU torch pipeline (baseline/_ref)
  N > 1: torchrun launches one rank per GPU (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).

Workload (BASELINE.json configs[1]): TenAnt, 4096 envs x 10 agents per GPU, horizon T = 16.  One "step" =
one pass of the hot path over one batch = one rollout of T frames x N envs:
    mmb_ten_ant_step   T frames in one launch: action scaling, observations, reward, done written straight into the
                       rollout storage slots; progress / reset chain and carry in the unit that closes an env
  + mmb_reset_compact  over the T flag rows (reset-index lists + DOF re-randomisation)             [side stream]
  + mmb_gae_ppo        reverse-time scan of RolloutStorage.compute_returns + fp64 statistics        [side stream 2]
  + mmb_adv_normalize  (N > 1: mmb_adv_normalize_xchg, statistics exchanged over NVLink in-kernel)  [side stream 2]
(--fused-gae folds the scan into the step kernel's chain executor instead: one launch fewer, measured slower.)
`value` = env-steps/s with the state frames already resident in HBM; `e2e` = the same metric through the
reference-facing API (VecTaskPython.step / RolloutStorage) with the frames and actions in pinned HOST memory
(H2D every env-step, D2H of reward/done every env-step and of the advantages every rollout).
Synthetic Isaac-layout frames (PhysX is out of scope); 4 rotating frame/storage sets (> L2) so HBM is measured.
The timed region is a CUDA graph of min(K, 32) rollouts (replayed K / 32 times + a remainder graph): the result does
not depend on K.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1 before torch is imported
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_k, None)

import argparse
import json
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (obs+reward+reset+GAE) at TenAnt 4096x10"
UNIT = "env-steps/s"
WORKLOAD = "TenAnt PPO 4096 envs x 10 agents, obs/reward/reset + horizon-16 GAE"
N_ENVS, HORIZON = 4096, 16
GAMMA, LAM = 0.96, 0.95
# algorithmic bytes of one TenAnt env-step in the horizon-batched kernel:
#   reads  root 572 + dof 640 + actions 320 + value 4                          = 1536
#   writes obs 1552 + forces 320 + reward 4 + done 1 + return 4 + advantage 4   = 1885
BYTES_PER_ENV_STEP_KERNEL = 1536 + 1885
BYTES_PER_ENV_STEP_KERNEL_UNFUSED = 1532 + 1877   # --no-fused-gae: without the value read and the two GAE planes
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self._stop_evt, self.ok = [], threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, t0, t1):
        names = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
                 0x100: "display_clock_setting"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in names.items() if bits & b], "samples": len(inside)}


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed summary of
    the `ncu --set full` capture of this same command (profiles/); the newest round's summary wins; None if missing."""
    import csv
    for name in ("r02_ten_ant_ncu_summary.csv", "r01_ten_ant_v7_ncu_summary.csv"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            vals = {r[0]: (float(r[1]), r[2]) for r in csv.reader(open(path)) if len(r) == 3 and r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum")}
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            if len(vals) == 2:
                return int(sum(v * scale[u] for v, u in vals.values())), os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own torch pipeline on the host cores
# ------------------------------------------------------------------------------------------------------
def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_rollouts_reference(n_rollouts, warmup, n_envs, horizon, threads):
    """The reference's OWN code from baseline/_ref (installed unmodified by baseline/make_ref.py): its `TenAnt` task
    class stepped through its `VecTaskPython.step` (agents/tasks/agent_base/vec_task.py:126-131), its `RolloutStorage.
    add_transitions / compute_returns` (agents/algorithms/rl/ppo/storage.py:32-65) and the obs copy of `PPO.run`
    (ppo.py:138), on the same synthetic Isaac-layout frames the GPU arm consumes, served by the FakeGym frame provider of
    oracle/refshim in place of PhysX (out of scope).  TenAnt only constructs in its 10-agent mode (ten_ant.py:91 views
    the sensor tensor as [N * num_agents, 24]); its step takes the flat (N, 80) actions either way."""
    import contextlib
    import io
    from oracle import refshim
    from massive_marl_benchmark_b200 import synthetic
    torch.set_num_threads(threads)
    refshim.install(REF_DIR)
    task, gym = refshim.make_task("TenAnt", n_envs, True)
    from agents.algorithms.rl.ppo.storage import RolloutStorage
    from agents.tasks.agent_base.vec_task import VecTaskPython
    with contextlib.redirect_stdout(io.StringIO()):
        env = VecTaskPython(task, "cpu")
    T, N = horizon, n_envs
    st = RolloutStorage(N, T, (388,), (0,), (80,), "cpu", "sequential")
    fr = synthetic.ten_ant_frames(N, T, seed=1234)
    gym.frames = [dict(root=fr["root"][t], dof=fr["dof"][t], sensor=None) for t in range(T)]
    values = torch.randn(N, 1); logp = torch.randn(N); mu = torch.randn(N, 80); sigma = torch.randn(N, 80)
    states = torch.zeros(N, 0); last_values = torch.randn(N, 1)
    gym.cursor = -1
    cur = env.step(fr["actions"][0])[0].clone()

    def rollout():
        gym.cursor = -1                     # FakeGym serves frame t at step t
        gym.log.clear()
        for t in range(T):
            obs, rew, done, _ = env.step(fr["actions"][t])
            st.add_transitions(cur, states, fr["actions"][t], rew, done, values, logp, mu, sigma)
            cur.copy_(obs)
        st.compute_returns(last_values, GAMMA, LAM)
        st.clear()

    for _ in range(warmup):
        rollout()
    t0 = time.perf_counter()
    for _ in range(n_rollouts):
        rollout()
    dt = time.perf_counter() - t0
    return n_rollouts * T * N / dt, dt / n_rollouts, torch.get_num_threads()


def cpu_rollouts_port(n_rollouts, warmup, n_envs, horizon, threads):
    """Fallback when baseline/_ref is absent: the oracle port of the same pipeline."""
    from oracle import storage_oracle as so
    from oracle.task_oracle import TenAntOracle, vec_task_step
    from massive_marl_benchmark_b200 import synthetic
    torch.set_num_threads(threads)
    fr = synthetic.ten_ant_frames(n_envs, horizon, seed=1234)
    orc = TenAntOracle(n_envs)
    T, N = horizon, n_envs
    obs_st = torch.zeros(T, N, 388); act_st = torch.zeros(T, N, 80); rew_st = torch.zeros(T, N, 1)
    done_st = torch.zeros(T, N, 1, dtype=torch.uint8)
    val_st = torch.randn(T, N, 1); last_values = torch.randn(N, 1)
    cur_obs = torch.zeros(N, 388)

    def rollout():
        for t in range(T):
            obs, rew, done = vec_task_step(lambda a: orc.step(a, fr["root"][t], fr["dof"][t]), fr["actions"][t], 1.0, 5.0)
            obs_st[t].copy_(cur_obs); act_st[t].copy_(fr["actions"][t]); rew_st[t].copy_(rew.view(-1, 1))
            done_st[t].copy_(done.view(-1, 1))
            cur_obs.copy_(obs)
        return so.ppo_compute_returns(rew_st, val_st, done_st, last_values, GAMMA, LAM)

    for _ in range(warmup):
        rollout()
    t0 = time.perf_counter()
    for _ in range(n_rollouts):
        rollout()
    dt = time.perf_counter() - t0
    return n_rollouts * T * N / dt, dt / n_rollouts, torch.get_num_threads()


def cpu_rollouts(n_rollouts, warmup, n_envs=N_ENVS, horizon=HORIZON, threads=None):
    """(env-steps/s, seconds per rollout, threads used, kind)"""
    threads = threads or _host_threads()
    if os.path.isdir(os.path.join(REF_DIR, "agents")):
        return cpu_rollouts_reference(n_rollouts, warmup, n_envs, horizon, threads) + ("reference",)
    return cpu_rollouts_port(n_rollouts, warmup, n_envs, horizon, threads) + ("port",)


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone runs; the others exit at once (no process group is ever created)."""
    if rank != 0:
        return
    # the same config as our arm: 4096 envs per GPU of the job, i.e. --gpus x 4096 envs in the one CPU pipeline.
    # A bounded sample of it: at most ~40 / gpus timed rollouts whatever --steps is (a rollout of 4096 envs is ~0.3-0.5 s
    # on 8-32 host cores), same warm-up rule as our arm (>= 3).
    n_envs = N_ENVS * max(1, args.gpus)
    steps = max(1, min(args.steps, max(3, 40 // max(1, args.gpus))))
    warm = min(max(args.warmup, 3), 10)
    value, sec, threads, kind = cpu_rollouts(steps, warm, n_envs=n_envs)
    note = ("the reference's own TenAnt.step + VecTaskPython.step + RolloutStorage from baseline/_ref on the host CPU (FakeGym frame "
            "provider in place of PhysX)") if kind == "reference" else \
           "oracle port of the reference's torch pipeline on the host CPU (baseline/_ref absent)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_timed": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "num_envs_per_gpu": N_ENVS, "num_agents": 10, "horizon": HORIZON,
                       "env_steps_per_step": HORIZON * n_envs, "note": note},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "%d rollouts of %d steps x %d envs + GAE (after %d warm-up rollouts)" % (steps, HORIZON, n_envs, warm)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def mlp_forward_bench(dev, M, iters=200):
    """PPO `ActorCritic.act` forward (actor + critic as one grouped launch per layer, sampling + log-prob kernel) at batch M:
    ms per call (CUDA events around `iters` back-to-back calls), TFLOP/s of the two MLPs and the fraction of the measured
    sustained bf16 peak.  Secondary figure: the headline metric does not include the policy forward."""
    from massive_marl_benchmark_b200.mlp import PPOActorCriticForward

    def net(out_dim, gain):
        dims, mods = [388, 1024, 1024, 512, out_dim], []
        for i in range(4):
            lin = torch.nn.Linear(dims[i], dims[i + 1])
            torch.nn.init.orthogonal_(lin.weight, gain=(gain if i == 3 else 2 ** 0.5))
            mods.append(lin)
            if i < 3:
                mods.append(torch.nn.ELU())
        return torch.nn.Sequential(*mods)

    class AC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor, self.critic = net(80, 0.01), net(1, 1.0)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(80))

    torch.manual_seed(0)
    fwd = PPOActorCriticForward(AC().to(dev), dev)
    obs = torch.clamp(torch.randn(M, 388, device=dev) * 2.0, -5, 5)
    for _ in range(5):
        fwd.act(obs, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fwd.act(obs, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # the same call replayed from a CUDA graph (what a graphed rollout step pays: no per-launch host cost)
    ms_graph = None
    try:
        fwd.seed = 1
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fwd.act(obs, None)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / iters
    except Exception:  # pragma: no cover
        pass
    flops = 2.0 * M * ((388 * 1024 + 1024 * 1024 + 1024 * 512 + 512 * 80) + (388 * 1024 + 1024 * 1024 + 1024 * 512 + 512 * 1))
    peak = None
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peak = float(json.load(open(pth)).get("bf16_tflops_sustained", 0.0)) or None
    best = ms_graph if ms_graph else ms
    tf = flops / (best * 1e-3) / 1e12
    return {"op": "PPO ActorCritic.act forward (actor + critic, bf16 operands / fp32 accumulate, tcgen05)", "batch": M,
            "ms_eager": ms, "ms_graph_replay": ms_graph, "tflops": tf, "peak_tflops": peak, "frac": (tf / peak) if peak else None,
            "note": "ONE clustered launch for both networks (mmb_mlp_chain, input cast folded into layer 0) + the sampling / log-prob kernel; "
                    "TFLOP/s from the graph-replayed time; not part of the headline metric"}


def ppo_rollout_bench(dev, frames, N, T, rollouts=10):
    """frames: list of frame sets (dicts of [T, rows, cols] device tensors), concatenated into one ring.  The rollout phase of the reference's PPO loop WITH the policy in it (ppo.py:127-157: act -> step -> add_transitions per
    env step, compute_returns per rollout) as one CUDA-graph replay per rollout (`ppo_rollout.GraphedPPORollout`), frames
    resident in HBM: env-steps/s over `rollouts` replays, CUDA events.  Secondary figure (the headline metric has no policy)."""
    from massive_marl_benchmark_b200.mlp import PPOActorCriticForward
    from massive_marl_benchmark_b200.ppo_rollout import GraphedPPORollout
    from massive_marl_benchmark_b200.providers import ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt
    from massive_marl_benchmark_b200.vec_task import VecTaskPython

    def net(out_dim):
        dims, mods = [388, 1024, 1024, 512, out_dim], []
        for i in range(4):
            mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
            if i < 3:
                mods.append(torch.nn.ELU())
        return torch.nn.Sequential(*mods)

    class AC(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.asymmetric = False
            self.actor, self.critic = net(80), net(1)
            self.log_std = torch.nn.Parameter(torch.log(torch.tensor(0.8)) * torch.ones(80))

    torch.manual_seed(0)
    prov = ReplayProvider({"root": torch.cat([f["root"] for f in frames]), "dof": torch.cat([f["dof"] for f in frames])}, device=dev)
    n_frames = prov.num_frames
    task = TenAnt({"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1}, provider=prov)
    st = RolloutStorage(N, T, (388,), (0,), (80,), dev)
    ro = GraphedPPORollout(VecTaskPython(task, dev), PPOActorCriticForward(AC().to(dev), dev), st, GAMMA, LAM)
    phases = max(1, n_frames // T) + 2
    for _ in range(phases + 1):          # one eager rollout, then every phase of the frame ring captured
        ro.run()
        st.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rollouts):
        ro.run()
        st.clear()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / rollouts
    return {"op": "PPO rollout phase with the policy in the loop (act -> step -> add_transitions x %d, compute_returns), one CUDA-graph "
                  "replay per rollout" % T, "num_envs": N, "horizon": T, "us_per_env_step": ms * 1e3 / T,
            "env_steps_per_s": N * T / (ms * 1e-3), "graphs_captured": ro.captures,
            "note": "frames resident in HBM; actor + critic 388-1024-1024-512-80/1 on tcgen05 (bf16 operands); not the headline metric"}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200 import dist as mdist
    from massive_marl_benchmark_b200 import synthetic
    from massive_marl_benchmark_b200.providers import HostReplayProvider, ReplayProvider
    from massive_marl_benchmark_b200.storage import RolloutStorage
    from massive_marl_benchmark_b200.tasks import TenAnt, reset_replay
    from massive_marl_benchmark_b200.vec_task import VecTaskPython

    numa_bound = mdist.bind_to_gpu_numa(local_rank) if world > 1 else False
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    N, T, K, W = N_ENVS, HORIZON, args.steps, max(args.warmup, 3)
    SETS = 4  # rotating frame/storage sets: 4 x ~225 MB of traffic per step >> 126 MB L2
    GROUP = int(os.environ.get("MMB_BENCH_GROUP", 8 * SETS))  # rollouts captured per CUDA graph (side-stream tails joined once per graph)
    fused = args.fused_gae

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 1 + rank}
    frames = [synthetic.ten_ant_frames(N, T, seed=1234 + 1000 * rank + s) for s in range(SETS)]
    dev_frames = [{k: v.to(dev) for k, v in f.items()} for f in frames]
    task = TenAnt(cfg, None, None, "cuda", local_rank, True, False,
                  provider=ReplayProvider({"root": frames[0]["root"], "dof": frames[0]["dof"]}, device=dev))
    task.clip_actions, task.clip_obs = 1.0, 5.0
    storages = [RolloutStorage(N, T, (388,), (0,), (80,), dev, "sequential") for _ in range(SETS)]
    forces = [torch.zeros(T, N, 80, device=dev) for _ in range(SETS)]
    dof_push = torch.zeros(T, 80 * N, 2, device=dev)
    # cross-shard advantage statistics: "p2p" = the normalise kernel publishes into every rank's mailbox over NVLink and
    # waits on its own (dist.StatsExchange, no collective launch), "nccl" = all-reduce between the two launches (baseline)
    mode = args.stats_exchange if world > 1 else "none"   # "none" at N>1 is a diagnostic (shard-local moments), not the product
    xchg = mdist.StatsExchange() if mode == "p2p" else None
    for st in storages:
        st.values.normal_()
        st.process_group = True if mode == "nccl" else None
        st.stats_exchange = xchg
    last_values = torch.randn(N, 1, device=dev)
    reset_out = [None]
    launches_per_rollout = [0]
    # high priority: the short tail kernels must get SM slots as soon as their rollout's step kernel has finished, even
    # though the next step kernel (overlap_prev) is already filling the GPU
    side = torch.cuda.Stream(priority=-1)       # reset-index lists of the T steps
    side2 = torch.cuda.Stream(priority=-1)      # [GAE scan when not fused,] statistics exchange wait + normalisation

    side_done = [None] * SETS   # event: the side-stream tail of the last rollout that used set s has finished

    def rollout(i, join=True):
        """One rollout on frame/storage set i % SETS.  Main stream: the step kernel (the task-state chain orders
        consecutive rollouts; with the fused GAE it also leaves returns / raw advantages / statistics).  Side streams:
        the reset lists of the T steps; [statistics exchange] -> normalisation.  With join=False the side work is left
        running so that the NEXT rollouts' step kernels (other storage sets) overlap it; a set is reused only after its
        previous tails have finished (side_done)."""
        s = i % SETS
        st, fr = storages[s], dev_frames[s]
        main = torch.cuda.current_stream()
        if side_done[s] is not None:
            main.wait_event(side_done[s])
        # observation after step t lands in obs slot t+1; reward/done of step t in slot t (no add_transitions pass)
        # consecutive rollouts use different frame / storage sets, so the step kernel may overlap the previous one's tail
        task.replay(fr, fr["actions"], st.obs_slots[1:], st.rewards.view(T, N), st.dones.view(T, N), None, forces[s],
                    overlap_prev=not args.no_overlap, gae=st.fused_gae(last_values, GAMMA, LAM) if fused else None,
                    chain_scratch=st.chain_scratch())
        side.wait_stream(main)
        with torch.cuda.stream(side):
            reset_out[0] = reset_replay(task, st.dones.view(T, N), dof_out=dof_push, out=reset_out[0])
        side2.wait_stream(main)
        with torch.cuda.stream(side2):
            if not fused:
                st.compute_returns_scan(last_values, GAMMA, LAM)
            st.normalize_advantages()
            side2.wait_stream(side)
            if not join:
                ev = torch.cuda.Event()
                ev.record(side2)
                side_done[s] = ev
        if join:
            main.wait_stream(side2)
            side_done[s] = None

    _l0 = L.launch_count()
    rollout(0)
    launches_per_rollout[0] = L.launch_count() - _l0
    for i in range(max(W, SETS)):
        rollout(i)
    barrier()

    # The timed region replays CUDA graphs of `n` consecutive rollouts whose side-stream tails are joined ONCE, at the end
    # of the graph: n = min(K, GROUP) for the bulk and K mod n for the remainder, so the measurement is the same for any K.
    graph_cache = {}

    def group_graph(n):
        if n not in graph_cache:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for r_ in range(n):
                    rollout(r_, join=(r_ == n - 1))
                for s_ in range(SETS):
                    side_done[s_] = None
            graph_cache[n] = g
        return graph_cache[n]

    use_graph = not args.no_graph
    if use_graph:
        try:
            n_main = max(1, min(K, GROUP))
            for n_ in {n_main, K % n_main, max(1, min(W, GROUP)), W % max(1, min(W, GROUP))} - {0}:
                group_graph(n_).replay()
            torch.cuda.synchronize()
        except Exception as ex:  # pragma: no cover
            print("cuda graph capture failed, running eagerly: %r" % (ex,), file=sys.stderr)
            use_graph = False

    def run_steps(count):
        """`count` consecutive rollouts."""
        if not use_graph:
            for i in range(count):
                rollout(i, join=(i == count - 1))
            for s_ in range(SETS):
                side_done[s_] = None
            return
        n = max(1, min(count, GROUP))
        for _ in range(count // n):
            group_graph(n).replay()
        if count % n:
            group_graph(count % n).replay()

    run_steps(W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    # The K steps are timed ON THE DEVICE.  A ~0.2 ms delay kernel in front of the first event keeps the GPU busy while the
    # host enqueues the (graph) launches, so the interval between the two events is the device executing K steps back to back
    # - not the host's launch latency of the first graph (tens of microseconds, which at the driver's --steps 20 = 0.8 ms of
    # device work would otherwise be billed as ~1.5 us per step, and more with 8 processes sharing the host).
    if not args.no_prime:
        torch.cuda._sleep(int(2.0e5 * 1.9))
    ev0.record()
    run_steps(K)
    ev1.record()
    barrier()
    t_host1 = time.perf_counter()
    sampler.stop()
    launches = (L.launch_count() - launches0) if not use_graph else K * launches_per_rollout[0]
    my_ms = ev0.elapsed_time(ev1)
    ms = mdist.max_over_ranks(my_ms, dev)
    rank_ms = [my_ms]
    if world > 1:
        tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(tl, torch.tensor([my_ms], dtype=torch.float64, device=dev))
        rank_ms = [float(t.item()) for t in tl]
    value = K * T * N * world / (ms * 1e-3)
    chain_errors = task.chain_errors()

    # ---- per-kernel durations: the same rollouts launched eagerly, every kernel bracketed by CUDA events on its
    # launch stream (library-side, mmb_profile_*); inputs rotate exactly as above ---------------------------------
    KP = max(3, min(K, 200))
    L.profile_enable(True)
    L.profile_collect()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    pe0.record()
    for i in range(KP):
        rollout(i)
    pe1.record()
    barrier()
    L.profile_enable(False)
    prof = L.profile_collect()
    eager_ms_per_step = pe0.elapsed_time(pe1) / KP

    # ---- the dominant kernel alone, launched back to back over the same rotating inputs (a CUDA graph of GROUP
    # launches, replayed): its sustained duration without the event-pair and launch gaps of the eager pass -------------
    def step_only(i):
        s_ = i % SETS
        st_ = storages[s_]
        task.replay(dev_frames[s_], dev_frames[s_]["actions"], st_.obs_slots[1:], st_.rewards.view(T, N),
                    st_.dones.view(T, N), None, forces[s_], overlap_prev=not args.no_overlap,
                    gae=st_.fused_gae(last_values, GAMMA, LAM) if fused else None, chain_scratch=st_.chain_scratch())
    sustained_ms, RK = None, 0
    if not args.no_graph:
        try:
            for i in range(SETS):
                step_only(i)
            torch.cuda.synchronize()
            gk = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gk):
                for i in range(GROUP):
                    step_only(i)
            gk.replay()
            RK = max(2, min(K // GROUP, 25))
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            k0.record()
            for _ in range(RK):
                gk.replay()
            k1.record()
            torch.cuda.synchronize()
            sustained_ms = k0.elapsed_time(k1) / (RK * GROUP)
            for st_ in storages:      # the statistics this pass accumulated were never consumed: start clean
                st_._adv_stats4.zero_()
        except Exception as ex:  # pragma: no cover
            print("kernel-only graph failed: %r" % (ex,), file=sys.stderr)

    # ---- end-to-end through the reference-facing API with host buffers -----------------------------
    K2 = max(2, min(K, 20))
    # frames AND actions live in pinned host memory; the provider uploads step t+1's inputs on a copy stream while
    # step t computes (every byte still crosses PCIe inside the timed region)
    host_prov = HostReplayProvider({"root": torch.cat([f["root"] for f in frames[:2]]),
                                    "dof": torch.cat([f["dof"] for f in frames[:2]]),
                                    "actions": torch.cat([f["actions"] for f in frames[:2]])}, dev, extra_keys=("actions",),
                                   packed=os.environ.get("MMB_HOST_PACKED", "1") != "0")
    cfg2 = {"env": {"numEnvs": N, "env_name": "ten_ant"}, "sim": {"dt": 0.0166}, "seed": 7 + rank}
    task2 = TenAnt(cfg2, None, None, "cuda", local_rank, True, False, provider=host_prov)
    task2.keep_raw_obs = False
    env = VecTaskPython(task2, dev)
    st2 = RolloutStorage(N, T, (388,), (0,), (80,), dev, "sequential")
    st2.process_group = True if mode == "nccl" else None
    st2.stats_exchange = xchg
    h_rew = torch.empty(N, pin_memory=True); h_done = torch.empty(N, dtype=torch.int64, pin_memory=True)
    h_adv = torch.empty(T, N, 1, pin_memory=True)
    values = torch.randn(N, 1, device=dev); logp = torch.randn(N, device=dev)
    mu = torch.randn(N, 80, device=dev); sigma = torch.randn(N, 80, device=dev); states = torch.zeros(N, 0, device=dev)
    cur_obs = env.reset()

    def e2e_rollout(j):
        # the loop of the reference's PPO.run (ppo.py:127-139): step -> add_transitions(current_obs, ...) -> current_obs.copy_(next_obs)
        for t in range(T):
            a = host_prov.stage[(host_prov.cursor + 1) & 1]["actions"]               # uploaded with the frame (H2D)
            obs, rew, done, _ = env.step(a)                                          # waits for that upload only
            st2.add_transitions(cur_obs, states, a, rew, done, values, logp, mu, sigma)
            cur_obs.copy_(obs)
            h_rew.copy_(rew, non_blocking=True); h_done.copy_(done, non_blocking=True)  # D2H result of the step
        st2.compute_returns(last_values, GAMMA, LAM)
        h_adv.copy_(st2.advantages, non_blocking=True)                              # D2H result of the rollout
        st2.clear()

    for j in range(2):
        e2e_rollout(j)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(K2):
        e2e_rollout(2 + j)
    e1.record()
    barrier()
    e2e_ms = mdist.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = K2 * T * N * world / (e2e_ms * 1e-3)
    h2d = T * host_prov.h2d_bytes_per_frame
    d2h = T * (N * 4 + N * 8) + T * N * 4

    # ---- the one dense contraction on the path: the actor-critic MLP forward the rollout calls once per env step
    # (module.py:25-55,73-87: 388 -> 1024 -> 1024 -> 512 -> 80 and -> 1, ELU) on the tcgen05 kernels, at this config's batch ----
    mlp = None
    if world == 1 and not args.no_mlp:
        try:
            mlp = mlp_forward_bench(dev, N)
        except Exception as ex:  # pragma: no cover
            mlp = {"error": repr(ex)}
    ppo_ro = None
    if world == 1 and not args.no_mlp:
        try:
            ppo_ro = ppo_rollout_bench(dev, dev_frames[:2], N, T)
        except Exception as ex:  # pragma: no cover
            ppo_ro = {"error": repr(ex)}

    xchg_errors = int(mdist.sum_over_ranks(float(xchg.errors), dev)) if xchg is not None else 0
    chain_errors = int(mdist.sum_over_ranks(float(chain_errors), dev))
    if xchg_errors:
        raise RuntimeError("statistics exchange reported %d timed-out / overrun exchanges" % xchg_errors)
    if chain_errors:
        raise RuntimeError("the in-kernel progress / reset chain timed out %d times" % chain_errors)
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic, traffic_src = ncu_traffic_bytes()
    k_ms, k_n = prof.get("ten_ant", (0.0, 0))
    per_launch_bytes = (BYTES_PER_ENV_STEP_KERNEL if fused else BYTES_PER_ENV_STEP_KERNEL_UNFUSED) * T * N
    eager_launch_ms = k_ms / max(k_n, 1)
    launch_ms = sustained_ms if sustained_ms else eager_launch_ms
    achieved = per_launch_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms else None
    shares = {k: round(v[0] / max(1e-9, sum(x[0] for x in prof.values())), 4) for k, v in prof.items()}
    # the whole timed step against the same roofline: every algorithmic byte of the rollout (step kernel + normalise pass
    # 8 B per transition + the flag rows the reset compaction reads) over the driver-visible time per step
    step_bytes = per_launch_bytes + T * N * (8 + 1) + (0 if fused else T * N * 17)
    cpu = None
    if args.cpu_rollouts > 0 and world == 1:
        # only on a single-process run: under torchrun the other ranks would spin in a collective for the duration, and
        # torchrun's OMP_NUM_THREADS=1 starves the CPU arm (see --impl reference, which handles both)
        c_val, c_sec, c_thr, c_kind = cpu_rollouts(args.cpu_rollouts, 1)
        cpu = {"value": c_val, "unit": UNIT, "cores": c_thr, "kind": c_kind,
               "sample": "%d rollouts of %d steps x %d envs + GAE (%s)" % (
                   args.cpu_rollouts, T, N, "the reference's own TenAnt / VecTaskPython / RolloutStorage from baseline/_ref"
                   if c_kind == "reference" else "oracle port of the reference's torch pipeline")}
    elif world > 1:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "skipped",
               "sample": "rank 0 at N = 1 only (run `bench.py --impl reference --gpus %d` for the CPU arm of this job size)" % world}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "ms_per_step_by_rank": [round(x / K, 5) for x in rank_ms], "ms_per_step_eager_profiled": eager_ms_per_step,
        "cuda_graph": use_graph,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "num_envs_per_gpu": N, "num_agents": 10, "horizon": T, "env_steps_per_step": T * N,
                   "l2": "4 rotating frame/storage sets, ~225 MB of traffic per step each, > 126 MB L2",
                   "parallelism": "env-sharded dp%d" % world,
                   "timed_region": ("CUDA graphs of min(steps, %d) rollouts, side-stream tails joined once per graph; device time between two "
                                    "CUDA events%s") % (GROUP, "" if args.no_prime else ", launches enqueued behind a 0.2 ms delay kernel (host launch latency excluded)"),
                   "gae": "fused into the step kernel's chain executor" if fused else "mmb_gae_ppo on a side stream",
                   "stored_planes": "obs, rewards, dones, returns, advantages, forces (values are inputs; actions / mu / sigma / log-prob "
                                    "planes belong to the policy forward, not to this metric, and are not written)",
                   "stats_exchange": {"p2p": "NVLink peer-memory mailboxes written and awaited inside the normalise kernel (no collective launch)",
                                      "nccl": "NCCL all-reduce of 3 doubles", "none": "single shard"}[mode]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": K2, "numa_bound": numa_bound, "api": "VecTaskPython.step + RolloutStorage.add_transitions/compute_returns "
                "in the loop of the reference's PPO.run (current_obs.copy_(next_obs)), pinned host frames"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "ten_ant_split_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": launch_ms,
                     "timing": ("CUDA events around %d back-to-back launches of the kernel alone (graph of %d launches over the "
                                "rotating sets, replayed)" % (RK * GROUP, GROUP)) if sustained_ms else "event pair per launch, eager pass",
                     "avg_launch_ms_eager_event_pairs": eager_launch_ms,
                     "launches_timed": (RK * GROUP) if sustained_ms else k_n, "kernel_time_shares_eager": shares,
                     "whole_step": {"algorithmic_bytes": step_bytes, "ms": ms / K,
                                    "achieved": step_bytes / (ms / K * 1e-3) / 1e9, "frac": step_bytes / (ms / K * 1e-3) / 1e9 / peak}},
        "cpu_baseline": cpu,
        "mlp_forward": mlp,
        "ppo_rollout": ppo_ro,
        "clocks": sampler.summary(t_host0, t_host1),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rollouts", type=int, default=32, help="rollouts of the CPU baseline sample (~0.3 s each on 16 cores); N = 1 only")
    ap.add_argument("--stats-exchange", default="p2p", choices=["p2p", "nccl", "none"],
                    help="multi-GPU advantage statistics: NVLink peer-memory mailboxes (default) or NCCL all-reduce")
    ap.add_argument("--no-overlap", action="store_true", help="ordinary stream order between consecutive step kernels (no PDL)")
    ap.add_argument("--no-graph", action="store_true", help="launch the rollout eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-prime", action="store_true", help="no delay kernel in front of the timed region (host launch latency then counts)")
    ap.add_argument("--no-mlp", action="store_true", help="skip the secondary MLP-forward figure")
    ap.add_argument("--fused-gae", action="store_true",
                    help="GAE scan inside the step kernel's chain executor instead of its own launch (mmb_gae_ppo on a side stream, the "
                         "default: measured 1.5 us per rollout faster, DESIGN.md section 4)")
    ap.add_argument("--no-fused-gae", action="store_true", help="(default; kept for older command lines)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)      # rank 0 only; no process group, no GPU
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    if world > 1:
        from massive_marl_benchmark_b200 import dist as mdist
        mdist.init_from_env("nccl")
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

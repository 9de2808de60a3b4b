"""BASELINE configs[3]: TenAnt IPPO / MAPPO team update, 16384 envs sharded over the GPUs of one box, NCCL gradient
all-reduce overlapped with the next agent's backward.

    python tools/bench_team_update.py --algo ippo                                     (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/bench_team_update.py --algo ippo                                         (G GPUs, strong scaling)

Every rank owns envs/G environments with its own per-agent buffers (synthetic contents of TenAnt's widths: obs 46,
share_obs 388, act 8), `compute_returns` has run, and ONE update = `TeamUpdate.train()` = ppo_epoch x num_mini_batch
minibatch steps of all 10 agents (forward + fused loss + backward of 20 networks, per-agent all-reduce of 5 MB launched while
the next agent computes, one clip + Adam launch pair for the team).  Timed with CUDA events, max over ranks.  Also measured:
the same update without the all-reduce (what the overlap hides), a standalone all-reduce of the whole 51 MB gradient buffer
(algorithm / bus bandwidth), the reference's own per-agent trainer loop on the same data (1 GPU only), and the replica
consistency (bit-identical parameters on every rank after the updates).
Needs the reference's policies / trainers (baseline/_ref, installed by baseline/make_ref.py) under oracle/refshim.
Prints ONE JSON line on rank 0 and writes gpurun_out/team_update_<algo>_<G>gpu.json.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="ippo", choices=["ippo", "mappo"])
    ap.add_argument("--envs", type=int, default=16384, help="TOTAL environments of the job (sharded over the ranks)")
    ap.add_argument("--agents", type=int, default=10)
    ap.add_argument("--updates", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reference-loop", action="store_true", help="also time the reference's per-agent trainers (world 1)")
    args = ap.parse_args()

    from oracle import refshim
    refshim.install(REF)
    import yaml
    import torch.distributed as dist
    from massive_marl_benchmark_b200 import dist as mdist
    from massive_marl_benchmark_b200 import spaces
    from massive_marl_benchmark_b200.runner import resolve_algorithm
    from massive_marl_benchmark_b200.separated_buffer import SeparatedReplayBuffer
    from massive_marl_benchmark_b200.team_update import TeamUpdate

    rank, world, local_rank = mdist.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    A, N = args.agents, args.envs // world
    config = yaml.safe_load(open(os.path.join(REF, "cfg", args.algo, "config.yaml")))
    T = config["episode_length"]
    config.update(n_rollout_threads=N)
    TrainAlgo, Policy = resolve_algorithm(args.algo)
    ob = spaces.Box(low=-np.inf, high=np.inf, shape=(46,))
    sh = spaces.Box(low=-np.inf, high=np.inf, shape=(388,))
    ac = spaces.Box(low=-np.ones(8), high=np.ones(8))
    cent = sh if config["use_centralized_V"] else ob

    def team():
        torch.manual_seed(3)                              # identical initial replicas on every rank
        with contextlib.redirect_stdout(io.StringIO()):
            pols = [Policy(config, ob, cent, ac, device=dev) for _ in range(A)]
            trs = [TrainAlgo(config, p, device=dev) for p in pols]
        return pols, trs

    def buffers(trs):
        g = torch.Generator(device=dev).manual_seed(100 + rank)   # every rank its own env shard
        bufs = [SeparatedReplayBuffer(config, ob, cent, ac, dev) for _ in range(A)]
        for a, b in enumerate(bufs):
            for name, scale in (("share_obs", 1.0), ("obs", 1.0), ("actions", 0.3), ("value_preds", 0.5), ("rewards", 1.0)):
                t = getattr(b, name)
                t.copy_(torch.randn(t.shape, generator=g, device=dev) * scale)
            b.action_log_probs.copy_(-torch.rand(b.action_log_probs.shape, generator=g, device=dev))
            b.masks.copy_((torch.rand(b.masks.shape, generator=g, device=dev) > 0.02).float())
            b.compute_returns(torch.randn(N, 1, generator=g, device=dev), trs[a].value_normalizer)
        return bufs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return mdist.max_over_ranks(e0.elapsed_time(e1) / n, dev)

    pols, trs = team()
    bufs = buffers(trs)
    tu = TeamUpdate(trs, bufs, config, algorithm=args.algo)
    for _ in range(args.warmup):
        tu.train()
    ms = timed(tu.train, args.updates)
    checksum = tu.replica_checksum()
    grad_bytes = tu.opt.total * 4
    red_calls = tu.reducer.calls if tu.reducer is not None else 0
    red_bytes = tu.reducer.bytes_reduced if tu.reducer is not None else 0
    steps_per_update = config["ppo_epoch"] * config["num_mini_batch"]

    # the same update without the gradient all-reduce (replicas drift: timing only) = what the overlap has to hide
    ms_local = None
    allreduce_ms = None
    if world > 1:
        pols2, trs2 = team()
        bufs2 = buffers(trs2)
        tu2 = TeamUpdate(trs2, bufs2, config, algorithm=args.algo, data_parallel=False)
        for _ in range(args.warmup):
            tu2.train()
        ms_local = timed(tu2.train, args.updates)
        flat = tu.opt.flat_grads
        for _ in range(3):
            dist.all_reduce(flat)
        allreduce_ms = timed(lambda: dist.all_reduce(flat), 20)
        del tu2, bufs2, trs2, pols2

    ref_ms = None
    if args.reference_loop and world == 1:
        pols3, trs3 = team()
        bufs3 = buffers(trs3)

        def ref_update():
            for a in torch.randperm(A).tolist():
                trs3[a].prep_training()
                bufs3[a].update_factor(torch.ones(T, N, 1, device=dev))
                trs3[a].train(bufs3[a])
                bufs3[a].after_update()
        for _ in range(2):
            ref_update()
        ref_ms = timed(ref_update, max(2, args.updates // 3))

    if rank == 0:
        samples = steps_per_update * T * N * world * A                      # agent-transitions consumed per update
        out = {
            "metric": "TenAnt %s team update (10 agents, 20 networks), %d envs sharded over %d GPU(s)" % (args.algo.upper(), args.envs, world),
            "value": samples / (ms * 1e-3), "unit": "agent-transitions/s", "n_gpus": world, "ms_per_update": ms,
            "ms_per_minibatch_step": ms / steps_per_update, "scaling": "strong",
            "config": {"algo": args.algo, "envs_total": args.envs, "envs_per_gpu": N, "episode_length": T, "agents": A,
                       "ppo_epoch": config["ppo_epoch"], "num_mini_batch": config["num_mini_batch"], "hidden_size": config["hidden_size"]},
            "gradient_bytes": grad_bytes, "allreduce_calls_per_update": red_calls // max(1, args.warmup + args.updates),
            "allreduce_bytes_per_update": red_bytes // max(1, args.warmup + args.updates),
            "ms_per_update_without_allreduce": ms_local,
            "exposed_communication_ms": (ms - ms_local) if ms_local else None,
            "standalone_allreduce": None if allreduce_ms is None else {
                "bytes": grad_bytes, "ms": allreduce_ms, "algbw_gbs": grad_bytes / allreduce_ms / 1e6,
                "busbw_gbs": grad_bytes / allreduce_ms / 1e6 * 2 * (world - 1) / world},
            "replica_checksum_max_minus_min": checksum,
            "reference_per_agent_loop_ms": ref_ms, "speedup_vs_reference_loop": (ref_ms / ms) if ref_ms else None,
            "nccl_ranks": world, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        }
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "team_update_%s_%dgpu.json" % (args.algo, world)), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

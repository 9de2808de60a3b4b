"""BASELINE configs[4]: GAE + minibatch-shuffle sweep, 1 M - 64 M transitions, env-sharded over 1 / 2 / 4 / 8 GPUs, beside the
reference's CPU path.

    python tools/sweep_storage.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/sweep_storage.py

For every total size T x N (T = 16; N = 65,536 ... 4,194,304 envs, N / G per rank; payload per transition = OneAnt-width
observation 60 floats + actions / mu / sigma 8 floats each + 5 scalars = 356 B):
  gae        RolloutStorage.compute_returns: reverse-time scan + statistics + normalisation; with G > 1 the advantage
             statistics are exchanged over NVLink inside the normalise kernel (dist.StatsExchange).  25 B / transition.
  gather     one epoch of shuffle + gather of all fields in 4 minibatches:
               "indexed"   mini_batch_generator (device permutation -> int64 index tensor) + gather_minibatch
               "fused"     gather_epoch_minibatch: the bijection is evaluated inside the gather kernel, no index array,
                           shuffle_group 1 (single transitions) and 8 (groups of 8 consecutive envs travel together)
             algorithmic bytes / transition: index 8 + 8 (write + read, indexed only) + 2 x 356.
Times are CUDA events, max over ranks; rates are whole-job (all ranks).  The reference's own CPU RolloutStorage
(baseline/_ref) is timed at the two smallest sizes on rank 0 of the 1-GPU run.
Writes gpurun_out/sweep_storage_<G>gpu.json.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massive_marl_benchmark_b200 import dist as mdist  # noqa: E402
from massive_marl_benchmark_b200.storage import RolloutStorage  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ROW = 60 * 4 + 8 * 4 * 3 + 5 * 4


def main():
    import torch.distributed as dist
    rank, world, local_rank = mdist.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    xchg = mdist.StatsExchange() if world > 1 else None
    T = 16
    sizes = [int(s) for s in os.environ.get("SWEEP_ENVS", "65536,262144,1048576,4194304").split(",")]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timeit(fn, iters, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        return mdist.max_over_ranks(e0.elapsed_time(e1) / iters, dev)

    out = {"peak_gbs_per_gpu": PEAK, "n_gpus": world, "horizon": T, "payload_bytes_per_transition": ROW, "sizes": []}
    for n_total in sizes:
        N = n_total // world
        st = RolloutStorage(N, T, (60,), (0,), (8,), dev, "random")
        st.stats_exchange = xchg
        st.rewards.normal_(); st.values.normal_(); st.observations.normal_(); st.actions.normal_(); st.mu.normal_(); st.sigma.normal_()
        st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.01).to(torch.uint8))
        lv = torch.randn(N, 1, device=dev)
        rec = {"transitions": T * n_total, "envs_per_gpu": N}
        ms = timeit(lambda: st.compute_returns(lv, 0.96, 0.95), 20)
        rec["gae"] = {"ms": ms, "transitions_per_s": T * n_total / ms * 1e3, "gbs": 25 * T * n_total / ms / 1e6,
                      "frac_of_aggregate_peak": 25 * T * n_total / ms / 1e6 / (PEAK * world)}
        bufs = [None]

        def epoch_indexed():
            for idx in st.mini_batch_generator(4):
                bufs[0] = st.gather_minibatch(idx, bufs[0])

        def epoch_fused():
            st.new_epoch()
            for k in range(4):
                bufs[0] = st.gather_epoch_minibatch(k, 4, bufs[0])

        rec["gather"] = {}
        for name, fn, group, nbytes in (("indexed_group1", epoch_indexed, 1, 16 + 2 * ROW), ("fused_group1", epoch_fused, 1, 2 * ROW),
                                        ("indexed_group8", epoch_indexed, 8, 16 + 2 * ROW), ("fused_group8", epoch_fused, 8, 2 * ROW),
                                        ("fused_group16", epoch_fused, 16, 2 * ROW)):
            st.shuffle_group = group
            ms = timeit(fn, 5)
            rec["gather"][name] = {"ms_per_epoch": ms, "transitions_per_s": T * n_total / ms * 1e3,
                                   "gbs": nbytes * T * n_total / ms / 1e6,
                                   "frac_of_aggregate_peak": nbytes * T * n_total / ms / 1e6 / (PEAK * world)}
        if xchg is not None and xchg.errors:
            raise RuntimeError("statistics exchange errors: %d" % xchg.errors)
        out["sizes"].append(rec)
        del st, bufs
        torch.cuda.empty_cache()

    # ---- the reference's own CPU path at the smallest sizes (rank 0 of the single-GPU run only) ----
    if world == 1 and os.path.isdir(os.path.join(REF, "agents")) and os.environ.get("SWEEP_CPU", "1") != "0":
        from oracle import refshim
        refshim.install(REF)
        from agents.algorithms.rl.ppo.storage import RolloutStorage as RefStorage
        threads = len(os.sched_getaffinity(0))
        torch.set_num_threads(threads)
        out["cpu_reference"] = {"threads": threads, "sizes": []}
        for n_total in sizes[:2]:
            rs = RefStorage(n_total, T, (60,), (0,), (8,), "cpu", "random")
            rs.rewards.normal_(); rs.values.normal_(); rs.observations.normal_()
            rs.dones.copy_((torch.rand(T, n_total, 1) < 0.01).to(torch.uint8))
            lv = torch.randn(n_total, 1)
            t0 = time.perf_counter(); rs.compute_returns(lv, 0.96, 0.95); t_gae = time.perf_counter() - t0
            t0 = time.perf_counter()
            for indices in rs.mini_batch_generator(4):                       # storage.py:75-87 + the gathers of ppo.py:253-264
                for f in ("observations", "actions", "values", "returns", "actions_log_prob", "advantages", "mu", "sigma"):
                    src = getattr(rs, f)
                    src.view(-1, *src.size()[2:])[indices]
            t_gather = time.perf_counter() - t0
            out["cpu_reference"]["sizes"].append({"transitions": T * n_total, "gae_ms": t_gae * 1e3, "gather_epoch_ms": t_gather * 1e3,
                                                  "gae_transitions_per_s": T * n_total / t_gae,
                                                  "gather_transitions_per_s": T * n_total / t_gather})
    if rank == 0:
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_storage_%dgpu.json" % world), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

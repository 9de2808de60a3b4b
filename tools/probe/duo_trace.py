"""Timeline inside one dual-network chain launch (PPO actor + critic, M = 4096): %globaltimer stamps of the first cluster's
four CTAs per task (net, layer), relative to the first stamp.  Needs a -DMMB_CHAIN_TRACE_BUILD library:
  make -C massive_marl_benchmark_b200/csrc EXTRA=-DMMB_CHAIN_TRACE_BUILD BUILD=build_trace OUT=../../tools/probe/libmmb_b200_trace.so
  gpurun -- env MMB_LIB_PATH=tools/probe/libmmb_b200_trace.so MMB_CHAIN_TRACE=1 python tools/probe/duo_trace.py"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

os.environ.setdefault("MMB_CHAIN_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from massive_marl_benchmark_b200 import _lib as L  # noqa: E402
from massive_marl_benchmark_b200.mlp import FusedMLP, GroupedMLP  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def net(dims):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            mods.append(torch.nn.ELU())
    return torch.nn.Sequential(*mods).to(dev)


dims = [388, 1024, 1024, 512, 80]
pair = GroupedMLP([FusedMLP.from_sequential(net(dims), dev), FusedMLP.from_sequential(net(dims), dev)])
x = torch.randn(M, dims[0], device=dev)
for _ in range(20):
    pair([x, x])
torch.cuda.synchronize()
words = 8 * 6 * 24
buf = (C.c_uint64 * words)()
L.check(L.lib().mmb_mlp_chain_trace(buf, words), "mmb_mlp_chain_trace")
t = np.array(buf[:4 * 16 * 8], dtype=np.int64).reshape(4, 16, 8)
t0 = t[t > 0].min()
names = ["producer", "first_operands", "mma_issued", "acc_complete", "computed", "landed", "peers", "epi_begin"]
out = {}
for task in range(2 * (len(dims) - 1)):
    row = {n: [int(t[c, task, e] - t0) if t[c, task, e] > 0 else None for c in range(4)] for e, n in enumerate(names)}
    out["net%d_layer%d" % (task & 1, task >> 1)] = row
    print("net %d layer %d" % (task & 1, task >> 1), {k: v[0] for k, v in row.items()}, "| cta1", {k: v[1] for k, v in row.items() if k in ("first_operands", "mma_issued", "landed")})
kbt = np.array(buf[512:512 + 16 * 16], dtype=np.int64).reshape(16, 16)
req = np.array(buf[768:768 + 16 * 16], dtype=np.int64).reshape(16, 16)
for task in range(2, 2 * (len(dims) - 1)):
    row = [int(v - t0) for v in req[task] if v > 0]
    out["net%d_layer%d_request_ns" % (task & 1, task >> 1)] = row
    print("net %d layer %d producer requests, cta0:" % (task & 1, task >> 1), row)
for task in range(2 * (len(dims) - 1)):
    row = [int(v - t0) for v in kbt[task] if v > 0]
    out["net%d_layer%d_kblock_ns" % (task & 1, task >> 1)] = row
    print("net %d layer %d k-block operand arrival, cta0:" % (task & 1, task >> 1), row, "deltas", [b - a for a, b in zip(row, row[1:])])
print("end (last stamp)", int(t.max() - t0))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/duo_trace.json", "w"), indent=1)

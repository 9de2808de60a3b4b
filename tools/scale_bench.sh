#!/bin/bash
# The driver's scaling run in one gpurun --gpus 8 call: bench.py at N = 1, 2, 4, 8 back to back on the same box.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tools/scale_bench.sh'
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-rollouts 0 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/scale_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "failed", e); continue
    if n == 1:
        base = d["value"]
    print(n, "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "eff %.3f" % (d["value"] / (n * base) if base else 0),
          "by rank", d.get("ms_per_step_by_rank"), "e2e %.3g" % d["e2e"]["value"])
PY

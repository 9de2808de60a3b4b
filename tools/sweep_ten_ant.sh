#!/bin/bash
# usage: tools/sweep_ten_ant.sh "VAR=val VAR2=val" ... ; prints rollout us and main-kernel us for each environment setting
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 600 --warmup 20 --cpu-rollouts 0 2>/dev/null | tail -1)
  python - "$cfg" <<PY
import json, sys
d = json.loads('''$out''')
print("%-50s rollout %.2f us  kernel %.2f us  frac %.3f" % (sys.argv[1], d["ms_per_step"] * 1e3, d["roofline"]["avg_launch_ms"] * 1e3, d["roofline"]["frac"]))
PY
done

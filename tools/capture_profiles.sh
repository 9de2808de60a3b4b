#!/bin/bash
# One gpurun call that refreshes the profile artefacts of a round (recipe of /opt/skills/guides/B200_PROFILING.md):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/capture_profiles.sh r02'
# 1. bench.py without a profiler (the only numbers that count), 2. the launch list of the same command under ncu
# (cold-cache, serialised: compare SHARES), 3. one `--set full` capture of the dominant kernel.  Each ncu pass runs only
# after the plain command exited 0.  One GPU only - never run this under torchrun.  Afterwards, on the build box:
#   ncu -i gpurun_out/<tag>_ten_ant.ncu-rep --page raw --csv > profiles/<tag>_ten_ant_ncu_summary.csv
#   cp gpurun_out/<tag>_launches.csv gpurun_out/<tag>_bench_1gpu.json profiles/
set -o pipefail
tag=${1:-rXX}
kernel=${2:-ten_ant_split_kernel}
out=gpurun_out
mkdir -p $out
# --no-graph: eager launches, so that ncu sees ordinary kernel launches (graph replays need --graph-profiling node)
short="python bench.py --steps 8 --warmup 3 --cpu-rollouts 0 --no-graph --no-mlp"

python bench.py --steps 2000 --warmup 20 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
tail -c 600 $out/${tag}_bench_1gpu.json; echo

$short > $out/${tag}_plain.log 2>&1 || { echo "short bench failed"; tail -5 $out/${tag}_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    $short > $out/${tag}_ncu_launches.log 2>&1 || echo "launch-list pass failed (see ${tag}_ncu_launches.log)"

timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kernel -s 10 -c 1 \
    -o $out/${tag}_ten_ant -f $short > $out/${tag}_ncu_full.log 2>&1 || echo "--set full pass failed (see ${tag}_ncu_full.log)"
ls -la $out | tail -8

# the gather kernel (shuffle_group 8 then 1, 16 M transitions): launches 1-8 = group 8, 9-16 = group 1
gcase="python tools/probe/gather_case.py"
$gcase > $out/${tag}_gather_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 4 -c 1 -o $out/${tag}_gather_g8 -f $gcase > $out/${tag}_ncu_gather8.log 2>&1 || echo "gather g8 pass failed"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -s 12 -c 1 -o $out/${tag}_gather_g1 -f $gcase > $out/${tag}_ncu_gather1.log 2>&1 || echo "gather g1 pass failed"
ls -la $out | grep ${tag}

set -x
timeout 300 python -m pytest tests/test_gpu_ppo_rollout.py tests/test_gpu_mlp.py -x -q -m gpu > gpurun_out/r2s4_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_a.log
tail -5 gpurun_out/r2s4_pytest_a.log
timeout 200 python tools/bench_ppo_rollout.py > gpurun_out/r2s4_ppo_rollout.log 2>&1; tail -6 gpurun_out/r2s4_ppo_rollout.log
MMB_ACT_PDL=0 timeout 200 python bench.py --steps 20 --warmup 5 --cpu-rollouts 0 > gpurun_out/r2s4_bench_nopdl.json 2> gpurun_out/r2s4_bench_nopdl.err
timeout 200 python bench.py --steps 20 --warmup 5 --cpu-rollouts 0 > gpurun_out/r2s4_bench_pdl.json 2> gpurun_out/r2s4_bench_pdl.err
python - <<'PY'
import json
for f in ("nopdl","pdl"):
    try:
        d=json.load(open("gpurun_out/r2s4_bench_%s.json"%f)); print(f, d["ms_per_step"], d["mlp_forward"]["ms_eager"], d["mlp_forward"]["ms_graph_replay"], d["mlp_forward"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY

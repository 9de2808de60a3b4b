timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2s4_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_full.log
tail -4 gpurun_out/r2s4_pytest_full.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s4_bench20.json 2> gpurun_out/r2s4_bench20.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s4_bench20.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["mlp_forward"]["ms_graph_replay"], d["mlp_forward"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
PY

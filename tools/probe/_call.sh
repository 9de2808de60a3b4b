timeout 300 python -m pytest tests/test_gpu_ppo_rollout.py tests/test_gpu_storage.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/r2s4_pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s4_pytest_f.log
tail -12 gpurun_out/r2s4_pytest_f.log
timeout 200 python tools/bench_ppo_rollout.py 2>&1 | grep -v torch_fp32

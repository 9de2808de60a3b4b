#!/bin/bash
# Everything that is measured on N GPUs of one box, in one gpurun call:
#   tools/run_multi_gpu.sh N          (N = 2, 4, 8)
# bench.py (BASELINE configs[1] weak scaling, both arms), the TenAnt IPPO / MAPPO team update with the overlapped NCCL gradient
# all-reduce (configs[3]: 16384 envs sharded), the GAE + shuffle sweep (configs[4]).  Outputs under gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_${N}.json 2> gpurun_out/r2_scale_${N}.err
if [ "$N" -lt 8 ]; then
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/r2_scale_${N}_k2000.json 2> gpurun_out/r2_scale_${N}_k2000.err
fi
timeout 300 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_${N}_ref.json 2> gpurun_out/r2_scale_${N}_ref.err
timeout 400 $TR --master-port 29514 tools/bench_team_update.py --algo ippo --updates 5 > gpurun_out/r2_team_ippo_${N}.log 2>&1
if [ "$N" -lt 8 ]; then
timeout 400 $TR --master-port 29515 tools/bench_team_update.py --algo mappo --updates 5 > gpurun_out/r2_team_mappo_${N}.log 2>&1
fi
SWEEP_ENVS=65536,1048576,4194304 timeout 500 $TR --master-port 29516 tools/sweep_storage.py > gpurun_out/r2_sweep_${N}.log 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo_${N}.txt 2>&1
tail -c 600 gpurun_out/r2_scale_${N}.json; echo; tail -c 400 gpurun_out/r2_team_ippo_${N}.log

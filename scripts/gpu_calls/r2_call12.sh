#!/bin/bash
# 8-GPU box: the contract's multi-GPU bench launch at N=8 and N=2, then the configs[4] sweep at 2 / 4 / 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r2c12_bench8.json 2> gpurun_out/r2c12_bench8.err
echo "bench8 rc=$?"; tail -2 gpurun_out/r2c12_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --e2e-steps 0 --late-start 0 > gpurun_out/r2c12_bench2.json 2> gpurun_out/r2c12_bench2.err
echo "bench2 rc=$?"
python -c "
import json
for n in (8,2):
    d=json.load(open('gpurun_out/r2c12_bench%d.json'%n)); print(n, d['value'], d['ms_per_step'], d.get('late'), (d.get('e2e') or {}).get('value'), (d.get('e2e_packed') or {}).get('value'))
"
SWEEP_GPUS="2 4 8" SWEEP_STEPS=40 timeout 1500 bash scripts/sweep_cfg5.sh gpurun_out/sweep_cfg5.json 2>&1 | tail -20

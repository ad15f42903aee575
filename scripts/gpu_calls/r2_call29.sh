#!/bin/bash
# 8 GPUs: the full BASELINE configs[3] (4096 envs x 1024 ants) with the final build
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2c29_bench8.json 2> gpurun_out/r2c29_bench8.err
tail -1 gpurun_out/r2c29_bench8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('8 GPUs', '%.4e'%d['value'], '%.4f'%d['ms_per_step'], 'late', d.get('late',{}).get('ms_per_step'), 'e2e %.3e'%d['e2e']['value'], d['clocks'])"

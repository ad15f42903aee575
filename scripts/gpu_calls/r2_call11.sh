#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2c11_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c11_tests.log
tail -8 gpurun_out/r2c11_tests.log
timeout 900 python bench.py > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err
echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c11_bench.json'))
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['late'])
print({k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e_packed']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11_bench_driver.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2c11_bench_driver.json')); print('driver-style', d['value'], d['ms_per_step'], d['e2e']['value'])"
SWEEP_GPUS=1 SWEEP_STEPS=40 timeout 1200 bash scripts/sweep_cfg5.sh gpurun_out/sweep_cfg5.json 2>&1 | tail -8

#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2c8_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c8_tests.log
tail -14 gpurun_out/r2c8_tests.log
timeout 900 python bench.py > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2c8_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c8_bench.json'))
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['late'])
print({k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e_packed']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
"
timeout 600 python bench.py --workload cfg1 --steps 300 --warmup 20 > gpurun_out/r2c8_cfg1.json 2> gpurun_out/r2c8_cfg1.err; cat gpurun_out/r2c8_cfg1.json | cut -c1-700
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c8_ref.json 2> gpurun_out/r2c8_ref.err; cut -c1-300 gpurun_out/r2c8_ref.json
bash scripts/r2_ncu.sh r2_final > gpurun_out/r2c8_ncu.log 2>&1; tail -8 gpurun_out/r2c8_ncu.log

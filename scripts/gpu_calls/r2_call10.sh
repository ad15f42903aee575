#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2c10_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c10_tests.log
tail -12 gpurun_out/r2c10_tests.log
for g in 1 2 4 6 8; do
ANTS_ROLLOUT_GROUPS=$g timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c10_bench_g$g.json 2> gpurun_out/r2c10_bench_g$g.err
python -c "
import json; d=json.load(open('gpurun_out/r2c10_bench_g$g.json'))
print('groups $g', d['value'], d['ms_per_step'], d['gpu_launches'], d['late']['ms_per_step'], {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
"
done

#!/bin/bash
mkdir -p gpurun_out
for wl in cfg2 cfg3; do
timeout 600 python bench.py --workload $wl --steps 200 --warmup 20 --cpu-steps 60 > gpurun_out/r2c17_$wl.json 2> gpurun_out/r2c17_$wl.err
python -c "
import json; d=json.load(open('gpurun_out/r2c17_$wl.json'))
print('$wl', '%.4e' % d['value'], '%.4f' % d['ms_per_step'], 'late %.4f' % d['late']['ms_per_step'], {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()}, 'e2e %.3e' % d['e2e']['value'], 'cpu %.3e' % d['cpu_baseline']['value'])
"
done 2>&1 | tee gpurun_out/r2c17_workloads.txt
bash scripts/r2_ncu.sh r2_final > gpurun_out/r2c17_ncu.log 2>&1; tail -4 gpurun_out/r2c17_ncu.log

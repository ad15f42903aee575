#!/bin/bash
# final build: parity + bench lines of the three single-GPU configs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2c28_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/r2c28_bench_default.json 2> gpurun_out/r2c28_bench_default.err
timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/r2c28_bench_k200.json 2> gpurun_out/r2c28_bench_k200.err
timeout 900 python bench.py --workload cfg2 --steps 200 --warmup 20 > gpurun_out/r2c28_cfg2.json 2> gpurun_out/r2c28_cfg2.err
timeout 900 python bench.py --workload cfg3 --steps 200 --warmup 20 > gpurun_out/r2c28_cfg3.json 2> gpurun_out/r2c28_cfg3.err
python -c "
import json
for f in ('bench_default','bench_k200','cfg2','cfg3'):
    d=json.load(open('gpurun_out/r2c28_%s.json'%f))
    print(f, '%.4e'%d['value'], '%.4f'%d['ms_per_step'], 'late %.4f'%d['late']['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'cpu %.3e'%d['cpu_baseline']['value'], 'frac %.3f dram %.3f'%(d['roofline']['frac'], d['roofline']['dram_frac'] or 0), {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items()})
" > gpurun_out/r2c28_summary.txt 2>&1
cat gpurun_out/r2c28_tests.log gpurun_out/r2c28_summary.txt

#!/bin/bash
# perception block size 32 / k_env load batching: parity, variants, groups, ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2c20_tests.log 2>&1
export REC=compact8 ENVS=512 WARM=300
for rep in 1 2; do for v in antsrl_b200/lib/var_[123]*.so; do ANTS_LIB=$PWD/$v TAG=$(basename $v) timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done; done > gpurun_out/r2c20_vars.txt 2>&1
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1]))
print(sys.argv[2], '%.4e' % d['value'], '%.4f' % d['ms_per_step'], 'late %.4f' % d['late']['ms_per_step'], {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('perceive','env_update_move')})
" $1 "$2"; }
for v in var_2t32 var_4envold; do
ANTS_LIB=$PWD/antsrl_b200/lib/$v.so timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c20_$v.json 2> gpurun_out/r2c20_$v.err
summ gpurun_out/r2c20_$v.json "$v"
done > gpurun_out/r2c20_bench.txt 2>&1
for g in 2 3 5 6 8; do
ANTS_ROLLOUT_GROUPS=$g timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c20_g$g.json 2> gpurun_out/r2c20_g$g.err
summ gpurun_out/r2c20_g$g.json "groups $g"
done >> gpurun_out/r2c20_bench.txt 2>&1
WARM=300 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_perceive_rows -s 300 -c 1 -f -o gpurun_out/r2c20_k_perceive python scripts/r2_prof.py > gpurun_out/r2c20_ncu.log 2>&1
cat gpurun_out/r2c20_tests.log gpurun_out/r2c20_vars.txt gpurun_out/r2c20_bench.txt

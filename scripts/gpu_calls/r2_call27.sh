#!/bin/bash
mkdir -p gpurun_out
export REC=compact8 WARM=300
# (a) cfg3 (LAYOUT 1 kernel, 80 registers): shared memory / L1 split
for cv in default 58 72 86; do
if [ $cv = default ]; then unset ANTS_ROWS_CARVEOUT; else export ANTS_ROWS_CARVEOUT=$cv; fi
WL=cfg3 ENVS=2048 WARM=100 TAG="cfg3 carveout_$cv" timeout 600 python scripts/perceive_only.py 2>&1 | tail -1; done > gpurun_out/r2c27_cfg3.txt 2>&1
unset ANTS_ROWS_CARVEOUT
# (b) cache policy of the record gathers, cfg4
export ENVS=512
for v in libantsrl_b200 var_ld1 var_ld2 var_ld3 var_ld4; do ANTS_LIB=$PWD/antsrl_b200/lib/$v.so TAG=$v timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done > gpurun_out/r2c27_ld.txt 2>&1
# (c) k_env: state loads before the absorb pass
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1]))
print(sys.argv[2], '%.4e' % d['value'], '%.4f' % d['ms_per_step'], 'late %.4f' % d['late']['ms_per_step'], {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('perceive','env_update_move')})
" $1 "$2"; }
for v in libantsrl_b200 var_envhoist libantsrl_b200 var_envhoist; do
ANTS_LIB=$PWD/antsrl_b200/lib/$v.so timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c27_$v.json 2> gpurun_out/r2c27_$v.err
summ gpurun_out/r2c27_$v.json "$v"
done > gpurun_out/r2c27_env.txt 2>&1
cat gpurun_out/r2c27_cfg3.txt gpurun_out/r2c27_ld.txt gpurun_out/r2c27_env.txt

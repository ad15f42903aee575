#!/bin/bash
mkdir -p gpurun_out
for tpb in 256 512 1024; do for g in 1 4; do
ANTS_ENV_TPB=$tpb ANTS_ROLLOUT_GROUPS=$g timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --late-start 0 > gpurun_out/r2c14_t${tpb}_g$g.json 2> gpurun_out/r2c14_t${tpb}_g$g.err
python -c "
import json; d=json.load(open('gpurun_out/r2c14_t${tpb}_g$g.json'))
print('tpb $tpb groups $g', '%.4e' % d['value'], '%.4f' % d['ms_per_step'], {k: round(v['ms_per_step'],4) for k,v in d['kernels'].items() if k in ('perceive','env_update_move')})
"
done; done 2>&1 | tee gpurun_out/r2c14_shapes.txt

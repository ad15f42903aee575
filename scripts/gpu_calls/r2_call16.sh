#!/bin/bash
mkdir -p gpurun_out
for g in 2 3 4 5 6 8; do
ANTS_ROLLOUT_GROUPS=$g timeout 600 python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c16_g$g.json 2> gpurun_out/r2c16_g$g.err
python -c "
import json; d=json.load(open('gpurun_out/r2c16_g$g.json'))
print('groups $g', '%.4e' % d['value'], '%.4f' % d['ms_per_step'], 'late %.4f' % d['late']['ms_per_step'])
"
done 2>&1 | tee gpurun_out/r2c16_groups.txt

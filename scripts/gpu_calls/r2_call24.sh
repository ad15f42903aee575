#!/bin/bash
mkdir -p gpurun_out
export REC=compact8 ENVS=512 WARM=300
for rep in 1 2; do for cv in default 100 90 80 70 60 50; do
if [ $cv = default ]; then unset ANTS_ROWS_CARVEOUT; else export ANTS_ROWS_CARVEOUT=$cv; fi
TAG="carveout_$cv" timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done; done > gpurun_out/r2c24_carveout.txt 2>&1
cat gpurun_out/r2c24_carveout.txt

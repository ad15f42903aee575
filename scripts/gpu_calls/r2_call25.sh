#!/bin/bash
mkdir -p gpurun_out
export REC=compact8 ENVS=512 WARM=300
for v in t32 t64 t128; do for cv in default 45 50 58 65 72 86; do
if [ $cv = default ]; then unset ANTS_ROWS_CARVEOUT; else export ANTS_ROWS_CARVEOUT=$cv; fi
ANTS_LIB=$PWD/antsrl_b200/lib/var_$v.so TAG="$v carveout_$cv" timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done; done > gpurun_out/r2c25_carveout.txt 2>&1
cat gpurun_out/r2c25_carveout.txt

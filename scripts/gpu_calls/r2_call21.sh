#!/bin/bash
mkdir -p gpurun_out
export REC=compact8 ENVS=512 WARM=300
for rep in 1 2; do for v in antsrl_b200/lib/var_*.so; do ANTS_LIB=$PWD/$v TAG=$(basename $v) timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done; done > gpurun_out/r2c21_vars.txt 2>&1
cat gpurun_out/r2c21_vars.txt

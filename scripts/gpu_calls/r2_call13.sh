#!/bin/bash
mkdir -p gpurun_out
for v in antsrl_b200/lib/var_*.so; do echo "== $v"; ANTS_LIB=$PWD/$v ANTS_ROLLOUT_GROUPS=1 MODES=fused K=80 timeout 300 python scripts/r2_ab.py 2>&1 | tail -1; done > gpurun_out/r2c13_variants.txt 2>&1
cat gpurun_out/r2c13_variants.txt

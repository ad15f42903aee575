#!/bin/bash
mkdir -p gpurun_out
(
echo "== default (prefetch 4)"; MODES=fused,flat K=80 timeout 400 python scripts/r2_ab.py 2>&1 | tail -2
for v in antsrl_b200/lib/var_*.so; do echo "== $v"; ANTS_LIB=$PWD/$v MODES=fused K=80 timeout 300 python scripts/r2_ab.py 2>&1 | tail -1; done
) > gpurun_out/r2c7_variants.txt 2>&1
cat gpurun_out/r2c7_variants.txt
(
for bsz in 1 2 4 8; do ANTS_E2E_DENSE_FRACTION=0 ANTS_UNPACK_BATCH=$bsz python scripts/r2_e2e.py | sed "s/^/batch=$bsz /"; done
ANTS_E2E_DENSE_FRACTION=0 ANTS_NO_AVX512_TABLE=1 python scripts/r2_e2e.py | sed "s/^/per-sample avx512 /"
) > gpurun_out/r2c7_e2e.txt 2>&1
grep step_host gpurun_out/r2c7_e2e.txt

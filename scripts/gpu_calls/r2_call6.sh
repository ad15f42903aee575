#!/bin/bash
mkdir -p gpurun_out
for v in antsrl_b200/lib/var_*.so; do echo "== $v"; ANTS_LIB=$PWD/$v MODES=fused K=60 timeout 300 python scripts/r2_ab.py 2>&1 | tail -1; done > gpurun_out/r2c6_variants.txt 2>&1
cat gpurun_out/r2c6_variants.txt
(
python scripts/r2_e2e.py
ANTS_E2E_DENSE=1 python scripts/r2_e2e.py
ANTS_NO_AVX512=1 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0 ANTS_NO_AVX512=1 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0.15 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0.3 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0.45 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0 ANTS_HOST_THREADS=8 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0 ANTS_HOST_THREADS=12 python scripts/r2_e2e.py
ANTS_E2E_DENSE_FRACTION=0 ANTS_HOST_THREADS=24 python scripts/r2_e2e.py
PACKED_ONLY=1 python scripts/r2_e2e.py
) > gpurun_out/r2c6_e2e.txt 2>&1
cat gpurun_out/r2c6_e2e.txt | grep step_host

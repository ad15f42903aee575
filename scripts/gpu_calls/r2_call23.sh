#!/bin/bash
# perception kernel variants (rock channel by the whole warp, 2^23 age conversion, opaque record base, full-warp loop, block size)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2c23_tests.log 2>&1
export REC=compact8 ENVS=512 WARM=300
for v in antsrl_b200/lib/var_*.so; do ANTS_LIB=$PWD/$v TAG=$(basename $v) timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done > gpurun_out/r2c23_vars.txt 2>&1
for v in antsrl_b200/lib/var_*.so; do ANTS_LIB=$PWD/$v TAG=$(basename $v) timeout 300 python scripts/perceive_only.py 2>&1 | tail -1; done >> gpurun_out/r2c23_vars.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 20 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err
cat gpurun_out/r2c23_tests.log gpurun_out/r2c23_vars.txt

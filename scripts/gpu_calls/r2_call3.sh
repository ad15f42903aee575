#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2c3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c3_tests.log
tail -22 gpurun_out/r2c3_tests.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2c3_bench.err
MODES=fused timeout 600 python scripts/r2_ab.py > gpurun_out/r2c3_ab.txt 2>&1
tail -3 gpurun_out/r2c3_ab.txt

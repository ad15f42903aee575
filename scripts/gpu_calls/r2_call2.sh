#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 ) > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
tail -25 gpurun_out/r2c2_tests.log
timeout 600 python scripts/r2_ab.py > gpurun_out/r2c2_ab.txt 2>&1
cat gpurun_out/r2c2_ab.txt | tail -5

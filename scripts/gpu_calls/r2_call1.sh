#!/bin/bash
# round 2, GPU call 1: the whole GPU suite with the new bench-condition tests, the round-1 bench as this round's
# baseline, and the host-side bandwidth probe that bounds the end-to-end path
mkdir -p gpurun_out
nproc > gpurun_out/r2c1_host.txt; lscpu | head -25 >> gpurun_out/r2c1_host.txt; free -g >> gpurun_out/r2c1_host.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2c1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c1_tests.log
tail -30 gpurun_out/r2c1_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?"
timeout 300 scripts/microbench/host_bw 728 > gpurun_out/r2c1_hostbw.txt 2>&1
cat gpurun_out/r2c1_hostbw.txt

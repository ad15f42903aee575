#!/bin/bash
# usage: scripts/ncu_capture.sh <tag> [kernel-regex]  -- run under gpurun; writes gpurun_out/<tag>_*
TAG=${1:-r1}
KREGEX=${2:-k_perceive}
CMD="python bench.py --envs 128 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 4 -c 2 -f -o gpurun_out/${TAG}_${KREGEX} $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out/

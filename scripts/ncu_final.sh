#!/bin/bash
# Final round-1 evidence: launch list of the bench command + full capture of k_perceive after 300 steps (dispersed ants).
TAG=${1:-r1_final}
mkdir -p gpurun_out
CMD="python bench.py --envs 128 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
export REC=compact8 ENVS=128 WARM=300 TAG=$TAG
CMD2="python scripts/perceive_only.py"
$CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_perceive -s 305 -c 1 -f -o gpurun_out/${TAG}_k_perceive $CMD2 > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_plain2.log

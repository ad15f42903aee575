#!/bin/bash
# round-2 evidence: launch list of the bench command + full captures of the two step kernels at the bench batch
TAG=${1:-r2_final}
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --late-start 0"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu0.log 2>&1
WARM=300 python scripts/r2_prof.py > gpurun_out/${TAG}_plain2.log 2>&1 &&
WARM=300 ncu --set full --clock-control none --import-source on -k regex:k_env -s 300 -c 1 -f -o gpurun_out/${TAG}_k_env python scripts/r2_prof.py > gpurun_out/${TAG}_ncu1.log 2>&1
WARM=300 ncu --set full --clock-control none --import-source on -k regex:k_perceive_rows -s 300 -c 1 -f -o gpurun_out/${TAG}_k_perceive python scripts/r2_prof.py > gpurun_out/${TAG}_ncu2.log 2>&1
ANTS_ROLLOUT_GROUPS=1 WARM=300 ncu --set full --clock-control none --import-source on -k regex:k_perceive_rows -s 300 -c 1 -f -o gpurun_out/${TAG}_k_perceive_whole python scripts/r2_prof.py > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/${TAG}_*

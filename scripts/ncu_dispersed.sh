#!/bin/bash
# ncu capture of k_perceive in the dispersed regime (after 300 steps), both record formats. Run under gpurun.
mkdir -p gpurun_out
for REC in f64 compact; do
  CMD="python scripts/perceive_only.py"
  export REC ENVS=128 WARM=300 TAG=disp_$REC
  $CMD > gpurun_out/disp_${REC}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_perceive -s 305 -c 1 -f -o gpurun_out/disp_${REC} $CMD > gpurun_out/disp_${REC}_ncu.log 2>&1
  tail -2 gpurun_out/disp_${REC}_plain.log
done
ls -la gpurun_out | tail -5

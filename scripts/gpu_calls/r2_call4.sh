#!/bin/bash
mkdir -p gpurun_out
ls -la oracle/_ref | head -8 > gpurun_out/r2c4_ref.txt 2>&1
python -c "from oracle import ref_harness as r; print(r.REFERENCE_ROOT, r.reference_available())" >> gpurun_out/r2c4_ref.txt 2>&1
cat gpurun_out/r2c4_ref.txt
WARM=300 python scripts/r2_prof.py > gpurun_out/r2c4_plain.log 2>&1 &&
WARM=300 ncu --set full --clock-control none --import-source on -k regex:k_env -s 300 -c 1 -f -o gpurun_out/r2_k_env python scripts/r2_prof.py > gpurun_out/r2c4_ncu1.log 2>&1
tail -2 gpurun_out/r2c4_ncu1.log
WARM=300 ncu --set full --clock-control none --import-source on -k regex:k_perceive_rows -s 300 -c 1 -f -o gpurun_out/r2_k_perceive python scripts/r2_prof.py > gpurun_out/r2c4_ncu2.log 2>&1
tail -2 gpurun_out/r2c4_ncu2.log

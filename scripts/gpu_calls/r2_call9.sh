#!/bin/bash
mkdir -p gpurun_out
(GROUPS=1,2,4 CHUNKS=8,32 python scripts/r2_groups.py; ANTS_NO_FUSED=1 GROUPS=1,2 CHUNKS=8 python scripts/r2_groups.py) > gpurun_out/r2c9_groups.txt 2>&1
grep groups gpurun_out/r2c9_groups.txt

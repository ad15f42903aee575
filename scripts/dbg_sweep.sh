#!/bin/bash
# perceive kernel with parts disabled (ANTS_DBG bit 0: no obs store, 1: no exploration stamps, 2: all samples in a 16x16 corner)
export REC=compact ENVS=512 WARM=100
for d in 0 1 2 3 4 5 7; do ANTS_DBG=$d TAG=dbg$d python scripts/perceive_only.py 2>&1 | tail -1; done
ANTS_PERCEIVE_GENERIC=1 TAG=generic python scripts/perceive_only.py 2>&1 | tail -1

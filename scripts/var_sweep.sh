#!/bin/bash
# time the step-loop kernels with alternative builds of the library (antsrl_b200/lib/var_*.so)
export REC=${REC:-compact8} ENVS=512 WARM=100
for v in antsrl_b200/lib/var_*.so; do ANTS_LIB=$PWD/$v TAG=$(basename $v) python scripts/perceive_only.py 2>&1 | tail -1; done

#!/bin/bash
mkdir -p gpurun_out
python -c "from oracle import ref_harness as r; print(r.REFERENCE_ROOT, r.reference_available())" > gpurun_out/r2c5_ref.txt 2>&1; cat gpurun_out/r2c5_ref.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2c5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c5_tests.log
tail -18 gpurun_out/r2c5_tests.log
timeout 600 python scripts/r2_ab.py > gpurun_out/r2c5_ab.txt 2>&1
tail -3 gpurun_out/r2c5_ab.txt
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2c5_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c5_bench.json'))
print(d['value'], d['ms_per_step'], d['late'])
print(d['e2e']); print(d['e2e_packed']); print(d['cpu_baseline'])
"

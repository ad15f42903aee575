#!/bin/bash
# BASELINE configs[4] sweep on every GPU count the box offers (1, 2, 4, 8); merges the runs into profiles/sweep_cfg5.json
# usage (GPU box): bash scripts/sweep_cfg5.sh [out.json]
# SWEEP_GPUS="1 2 4 8" selects the GPU counts of this call (earlier runs found in gpurun_out/ are merged in)
OUT=${1:-gpurun_out/sweep_cfg5.json}
NG=$(nvidia-smi -L | wc -l)
GPUS=${SWEEP_GPUS:-"1 2 4 8"}
mkdir -p gpurun_out
for n in $GPUS; do
  if [ $n -eq 1 ]; then
    python scripts/sweep_cfg5.py > gpurun_out/sweep_1.json 2> gpurun_out/sweep_1.err || exit 1
  elif [ $NG -ge $n ]; then
    SWEEP_NO_CPU=1 SWEEP_ENVS=${SWEEP_ENVS_MULTI:-64,256,1024,4096,16384,65536} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      scripts/sweep_cfg5.py > gpurun_out/sweep_$n.json 2> gpurun_out/sweep_$n.err || echo "N=$n failed"
  fi
done
python - "$OUT" <<'PY'
import json, sys, glob
runs = []
for f in sorted(glob.glob("gpurun_out/sweep_[1248].json")):
    try:
        runs.append(json.loads(open(f).read().strip().splitlines()[-1]))
    except Exception as e:
        print("skipping", f, e)
json.dump({"config": "BASELINE.json configs[4]: env-count scaling sweep 64 -> 65536 envs x 256 ants (256x256 maps)", "runs": runs}, open(sys.argv[1], "w"), indent=1)
for r in runs:
    for row in r["rows"]:
        print(r["n_gpus"], row["envs_total"], "%.3f ms" % row["ms_per_step"], "%.3e" % row["ant_steps_per_s"])
PY

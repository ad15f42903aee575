#!/usr/bin/env python
"""Record one environment of a device-resident batch for the reference's viewer.

main.py:136-147 appends `env.save_state()` after every update and pickles the list into `saved/<name>.arl`;
`gui/visualize.py` replays that file.  Here E environments step on the GPU under the agents' exploration policy
(random actions drawn on the device, collect_agent.py:172-177) and environment `--env` is recorded with
antsrl_b200.snapshot.EpisodeRecorder: one ants_export_env_state per step, the other E - 1 environments stay in HBM.

    python examples/record_episode.py --out saved/gpu_episode.arl [--envs 64] [--steps 200] [--workload cfg2]
    # then, in the reference checkout:  python -c "from gui.visualize import Visualizer; ..."  (it lists saved/)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench                                             # noqa: E402
from antsrl_b200.snapshot import EpisodeRecorder         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--env", type=int, default=0, help="which environment of the batch to record")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default="saved/gpu_episode.arl")
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    batch = bench.make_generator(wl, a.steps).generate(a.envs)      # env e == the reference generator with seed 1000 + e
    rec = EpisodeRecorder(batch, env_index=a.env)
    batch.observe()                                                 # main.py:88
    for t in range(a.steps):
        rot, ph = batch.sample_actions(seed=2026)                   # keyed by (seed, env, timestep, ant)
        batch.step(rot, ph)                                         # main.py:98
        batch.update()                                              # main.py:131
        rec.record()                                                # main.py:136-137
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    rec.save(a.out)                                                 # main.py:139-147
    print("wrote %s: %d states of env %d (%dx%d map, %d ants)" % (a.out, a.steps, a.env, wl["w"], wl["h"], wl["n_ants"]))
    batch.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json configs[4]: env-count scaling sweep, 64 -> 65 536 environments (x4 steps) of 256 ants on 256x256 maps
(SURVEY 8-d cfg5 = the cfg3 map), at 1 / 2 / 4 / 8 GPUs (torchrun: the envs are sharded contiguously, no per-step
communication), with the unmodified reference's CPU loop on the host cores beside it.

    python scripts/sweep_cfg5.py                      # 1 GPU
    torchrun --nproc-per-node N ... scripts/sweep_cfg5.py
Writes / merges profiles/sweep_cfg5.json (one entry per GPU count) when run from scripts/sweep_cfg5.sh."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench

ENV_COUNTS = [int(x) for x in os.environ.get("SWEEP_ENVS", "64,256,1024,4096,16384,65536").split(",")]
STEPS, WARM, SLICE = int(os.environ.get("SWEEP_STEPS", "50")), 10, 2048


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from antsrl_b200 import BatchedAnts
    from antsrl_b200.generator import stack_states
    wl = bench.WORKLOADS["cfg3"]
    N = wl["n_ants"]
    gen = bench.make_generator(wl, 10000)
    rows = []
    for total in ENV_COUNTS:
        if total % world:
            continue
        E = total // world
        t0 = time.perf_counter()
        batch = BatchedAnts(gen.cfg, E, device=local_rank, evap_mode="lazy", record="compact8", rng_seed=5, env_id_base=rank * E)
        # upload in slices (65 536 maps never sit in host memory at once).  Up to 2048 distinct generated maps per rank;
        # larger batches reuse them cyclically (the environments still diverge: actions and Philox noise are per env id)
        stacked = stack_states(bench.generate_states_parallel(wl, 10000, rank * E, min(SLICE, E)), "all")
        for s0 in range(0, E, SLICE):
            n = min(SLICE, E - s0)
            part = stacked if n == min(SLICE, E) else {k: (v[:n] if hasattr(v, "shape") and v.ndim else v) for k, v in stacked.items()}
            batch.import_state(part, envs=(s0, n))
        del stacked
        batch.activate_all_pheromones(np.ones((E, N, 2)) * 10.0)
        rs = np.random.RandomState(7 + rank)
        T = 16
        rot = torch.from_numpy((rs.randint(0, 3, size=(T, E, N)) - 1).astype(np.int8)).cuda()
        ph = torch.from_numpy(rs.randint(0, 3, size=(T, E, N)).astype(np.int8)).cuda()
        batch.observe()
        setup_s = time.perf_counter() - t0

        def run(n):
            done = 0
            while done < n:
                m = min(T, n - done)
                batch.rollout(rot[:m], ph[:m]); done += m
        run(WARM)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record(); run(STEPS); ev1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        dev_bytes = batch.stats()["device_bytes"]
        batch.close()
        del rot, ph
        torch.cuda.empty_cache()
        if rank == 0:
            row = {"envs_total": total, "envs_per_gpu": E, "n_gpus": world, "ants_per_step": total * N, "steps": STEPS,
                   "ms_per_step": ms / STEPS, "ant_steps_per_s": total * N * STEPS / (ms / 1000.0),
                   "device_gb_per_gpu": dev_bytes / 1e9, "setup_s": setup_s}
            rows.append(row)
            sys.stderr.write(json.dumps(row) + "\n")
    out = {"n_gpus": world, "maps": "up to %d distinct generated maps per rank, reused cyclically beyond that" % SLICE, "workload": "cfg5 = cfg3 map: " + wl["desc"].split(",")[0] + ", 256 ants/env, compact8 records, lazy field, "
           "one ants_rollout call per %d steps" % 16, "rows": rows}
    if rank == 0 and world == 1 and not os.environ.get("SWEEP_NO_CPU"):
        out["cpu_reference"] = bench.run_cpu_baseline(wl, 3, int(os.environ.get("SWEEP_CPU_STEPS", "60")))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Experiment: the cfg4 shard as G independent groups of envs, each with its own handle and stream, driven by
ants_rollout calls of CHUNK steps issued round-robin (the kernels of one group overlap the other groups')."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS["cfg4"])
E = int(os.environ.get("ENVS", "512")); N = wl["n_ants"]
gen = bench.make_generator(wl, 5000)
states = bench.generate_states_parallel(wl, 5000, 0, E)
rs = np.random.RandomState(1)
T = 32
rot = torch.from_numpy((rs.randint(0, 3, size=(T, E, N)) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(T, E, N)).astype(np.int8)).cuda()
for G in [int(x) for x in os.environ.get("GROUPS", "1,2,3,4").split(",")]:
    for CHUNK in [int(x) for x in os.environ.get("CHUNKS", "8").split(",")]:
        Eg = E // G
        streams = [torch.cuda.Stream() for g in range(G)]
        groups, tapes = [], []
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                b = BatchedAnts(gen.cfg, Eg, evap_mode="lazy", record="compact8", env_id_base=g * Eg)
                b.import_state(stack_states(states[g * Eg:(g + 1) * Eg], "all"))
                b.activate_all_pheromones(np.ones((Eg, N, 2)) * 10.0)
                b.observe()
            groups.append(b)
            tapes.append((rot[:, g * Eg:(g + 1) * Eg].contiguous(), ph[:, g * Eg:(g + 1) * Eg].contiguous()))
        torch.cuda.synchronize()
        def run(n):
            done = 0
            while done < n:
                m = min(CHUNK, n - done)
                for g in range(G):
                    with torch.cuda.stream(streams[g]):
                        groups[g].rollout(tapes[g][0][:m], tapes[g][1][:m])
                done += m
        run(40)
        torch.cuda.synchronize()
        K = 160
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print("groups %d chunk %d fused %s: ms/step %.4f  ant-steps/s %.3e" % (G, CHUNK, "no" if os.environ.get("ANTS_NO_FUSED") else "yes", dt * 1e3, G * Eg * N / dt), flush=True)
        for b in groups:
            b.close()
        del groups, tapes
        torch.cuda.empty_cache()

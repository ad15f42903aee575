"""Experiment: the batch as G independent half-batches on their own streams (kernels of one group overlap the other's)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS[os.environ.get("WL", "cfg4")])
E = int(os.environ.get("ENVS", "512")); G = int(os.environ.get("NGROUPS", "2")); N = wl["n_ants"]
gen = bench.make_generator(wl, 1000)
states = bench.generate_states_parallel(wl, 1000, 0, E)
Eg = E // G
prio = os.environ.get("PRIO", "0") == "1"
streams = [torch.cuda.Stream(priority=(-1 if (prio and g % 2) else 0)) for g in range(G)]
groups = []
for g in range(G):
    with torch.cuda.stream(streams[g]):
        b = BatchedAnts(gen.cfg, Eg, evap_mode="lazy", record="compact", env_id_base=g * Eg)
        b.import_state(stack_states(states[g * Eg:(g + 1) * Eg], "all"))
        b.activate_all_pheromones(np.ones((Eg, N, 2)) * 10.0)
        b.observe()
    groups.append(b)
rs = np.random.RandomState(1)
rot = torch.from_numpy((rs.randint(0, 3, size=(16, E, N)) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(16, E, N)).astype(np.int8)).cuda()
torch.cuda.synchronize()
stagger = os.environ.get("STAGGER", "0") == "1"
def one_step(t):
    k = t % 16
    if stagger:
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                groups[g].step(rot[k, g * Eg:(g + 1) * Eg], ph[k, g * Eg:(g + 1) * Eg])
                groups[g].update(None)
    else:
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                groups[g].step(rot[k, g * Eg:(g + 1) * Eg], ph[k, g * Eg:(g + 1) * Eg])
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                groups[g].update(None)
for t in range(100): one_step(t)
torch.cuda.synchronize()
K = 200
t0 = time.perf_counter()
for t in range(K): one_step(100 + t)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print("groups", G, "stagger", stagger, "prio", prio, "ms/step %.4f" % (dt * 1e3), "ant-steps/s %.3e" % (E * N / dt))

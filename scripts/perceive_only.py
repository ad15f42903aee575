"""Time the step-loop kernels at a fixed workload (quick experiments)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS[os.environ.get("WL", "cfg4")])
if os.environ.get("N_ANTS"): wl["n_ants"] = int(os.environ["N_ANTS"])
E = int(os.environ.get("ENVS", "256"))
gen = bench.make_generator(wl, 1000)
states = bench.generate_states_parallel(wl, 1000, 0, E)
b = BatchedAnts(gen.cfg, E, evap_mode=os.environ.get("EVAP", "lazy"), record=os.environ.get("REC", "f64"))
b.import_state(stack_states(states, "all"))
b.activate_all_pheromones(np.ones((E, wl["n_ants"], 2)) * 10.0)
rs = np.random.RandomState(1)
rot = torch.from_numpy((rs.randint(0, 3, size=(16, E, wl["n_ants"])) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(16, E, wl["n_ants"])).astype(np.int8)).cuda()
b.observe()
for t in range(int(os.environ.get("WARM", "30"))):
    b.step(rot[t % 16], ph[t % 16]); b.update(None)
b.set_profiling(True); b.reset_kernel_ms()
K = 20
for t in range(K):
    b.step(rot[t % 16], ph[t % 16]); b.update(None)
km = b.kernel_ms()
print(os.environ.get("TAG", ""), {k: round(v[0] / K, 4) for k, v in km.items() if v[1]}, "total", round(sum(v[0] for v in km.values()) / K, 4))

"""Time Environment.update() with DIFFUSE_FACTOR != 0 (the dense 3x3 stencil path, pheromone.py:43-45)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS[os.environ.get("WL", "cfg3")])
E = int(os.environ.get("ENVS", "1024"))
gen = bench.make_generator(wl, 1000)
cfg = dict(gen.cfg); cfg["diffuse_factor"] = float(os.environ.get("DF", "0.02"))
states = bench.generate_states_parallel(wl, 1000, 0, E)
b = BatchedAnts(cfg, E, evap_mode="dense", record="f64")
b.import_state(stack_states(states, "all"))
b.activate_all_pheromones(np.ones((E, wl["n_ants"], 2)) * 10.0)
rs = np.random.RandomState(1)
rot = torch.from_numpy((rs.randint(0, 3, size=(16, E, wl["n_ants"])) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(16, E, wl["n_ants"])).astype(np.int8)).cuda()
b.observe()
for t in range(20):
    b.step(rot[t % 16], ph[t % 16]); b.update(None)
b.set_profiling(True); b.reset_kernel_ms()
K = 20
for t in range(K):
    b.step(rot[t % 16], ph[t % 16]); b.update(None)
km = b.kernel_ms()
ev = km["evaporate"][0] / K
cells = E * wl["w"] * wl["h"]
alg = cells * (16 * 2 + 1)
print({k: round(v[0] / K, 4) for k, v in km.items() if v[1]})
print("stencil: %.4f ms per update, %d cells, algorithmic %.1f MB -> %.0f GB/s (%.2f of 6546)" % (ev, cells, alg / 1e6, alg / ev / 1e6, alg / ev / 1e6 / 6546))

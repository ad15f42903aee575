"""End-to-end (host buffer) timing of ants_step_host on the cfg4 shard under the environment's settings
(ANTS_NO_AVX512, ANTS_E2E_DENSE_FRACTION, ANTS_HOST_THREADS, ANTS_E2E_DENSE)."""
import sys, os, json, time
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
b = BatchedAnts(gen.cfg, E, evap_mode="lazy", record="compact8", rng_seed=3)
b.import_state(stack_states(states, "all")); del states
b.activate_all_pheromones(np.ones((E, N, 2)) * 10.0)
rs = np.random.RandomState(1)
h_rot = b.pinned("rot", (E, N), np.int8); h_ph = b.pinned("ph", (E, N), np.int8)
h_rot[:] = (rs.randint(0, 3, size=(E, N)) - 1); h_ph[:] = rs.randint(0, 3, size=(E, N))
b.observe()
fn = b.step_host_packed if os.environ.get("PACKED_ONLY") else b.step_host
for _ in range(int(os.environ.get("WARM", "6"))):
    fn(h_rot, h_ph); b.update_host(None)
torch.cuda.synchronize()
K = int(os.environ.get("K", "15"))
ts = []
for _ in range(K):
    t0 = time.perf_counter(); fn(h_rot, h_ph); t1 = time.perf_counter(); b.update_host(None); torch.cuda.synchronize()
    ts.append((t1 - t0) * 1e3)
tag = " ".join("%s=%s" % (k, os.environ[k]) for k in ("ANTS_NO_AVX512", "ANTS_E2E_DENSE_FRACTION", "ANTS_HOST_THREADS", "ANTS_E2E_DENSE", "PACKED_ONLY") if k in os.environ)
print("%-60s step_host ms: median %.2f min %.2f  dense_permille %d" % (tag or "default", sorted(ts)[len(ts) // 2], min(ts), b.stats()["e2e_dense_permille"]), flush=True)

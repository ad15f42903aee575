"""Profiling driver: the bench workload (cfg4 shard, 512 envs x 1024 ants) fast-forwarded WARM steps through ants_rollout,
then a few more steps for ncu to capture (-k regex:... -s WARM -c 1)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS[os.environ.get("WL", "cfg4")])
E = int(os.environ.get("ENVS", str(wl["envs_per_gpu"])))
WARM = int(os.environ.get("WARM", "300"))
N = wl["n_ants"]
gen = bench.make_generator(wl, 5000)
states = bench.generate_states_parallel(wl, 5000, 0, E)
b = BatchedAnts(gen.cfg, E, evap_mode="lazy", record=os.environ.get("REC", "compact8"), rng_seed=3)
b.import_state(stack_states(states, "all"))
del states
b.activate_all_pheromones(np.ones((E, N, 2)) * 10.0)
rs = np.random.RandomState(1)
T = 32
rot = torch.from_numpy((rs.randint(0, 3, size=(T, E, N)) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(T, E, N)).astype(np.int8)).cuda()
b.observe()
done = 0
while done < WARM + 4:
    n = min(T, WARM + 4 - done)
    b.rollout(rot[:n], ph[:n]); done += n
torch.cuda.synchronize()
print("ok", done)

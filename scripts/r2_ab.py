"""A/B timing on the cfg4 shard: flat kernels vs block-per-env kernels, Python step loop vs ants_rollout."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = dict(bench.WORKLOADS[os.environ.get("WL", "cfg4")])
E = int(os.environ.get("ENVS", str(wl["envs_per_gpu"])))
K = int(os.environ.get("K", "100"))
WARM = int(os.environ.get("WARM", "20"))
gen = bench.make_generator(wl, 5000)
states = bench.generate_states_parallel(wl, 5000, 0, E)
st = stack_states(states, "all")
del states
N = wl["n_ants"]
rs = np.random.RandomState(1)
T = 64
rot = torch.from_numpy((rs.randint(0, 3, size=(T, E, N)) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(T, E, N)).astype(np.int8)).cuda()
out = {}
for mode in os.environ.get("MODES", "fused,flat").split(","):
    if mode == "flat":
        os.environ["ANTS_NO_FUSED"] = "1"
    else:
        os.environ.pop("ANTS_NO_FUSED", None)
    b = BatchedAnts(gen.cfg, E, evap_mode="lazy", record=os.environ.get("REC", "compact8"), rng_seed=3)
    b.import_state(st)
    b.activate_all_pheromones(np.ones((E, N, 2)) * 10.0)
    b.observe()
    for t in range(WARM):
        b.step(rot[t % T], ph[t % T]); b.update(None)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for t in range(K):
        b.step(rot[t % T], ph[t % T]); b.update(None)
    ev1.record(); torch.cuda.synchronize()
    out[mode + "_loop_ms"] = ev0.elapsed_time(ev1) / K
    ev0.record()
    done = 0
    while done < K:
        n = min(T, K - done)
        b.rollout(rot[:n], ph[:n]); done += n
    ev1.record(); torch.cuda.synchronize()
    out[mode + "_rollout_ms"] = ev0.elapsed_time(ev1) / K
    b.set_profiling(True); b.reset_kernel_ms()
    b.rollout(rot[:20], ph[:20])
    km = b.kernel_ms()
    out[mode + "_kernels"] = {k: round(v[0] / 20, 4) for k, v in km.items() if v[1]}
    b.set_profiling(False)
    b.close()
    print(mode, json.dumps({k: v for k, v in out.items() if k.startswith(mode)}), flush=True)

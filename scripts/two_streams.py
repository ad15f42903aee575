"""Experiment: S handles (sub-batches) on S CUDA streams vs one handle, same total envs."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from antsrl_b200 import BatchedAnts
from antsrl_b200.generator import stack_states
wl = bench.WORKLOADS["cfg4"]
E = int(os.environ.get("ENVS", "512")); S = int(os.environ.get("STREAMS", "2")); K = 100; WARM = int(os.environ.get("WARM", "50"))
gen = bench.make_generator(wl, 2000)
states = bench.generate_states_parallel(wl, 2000, 0, E)
per = E // S
streams = [torch.cuda.Stream() for _ in range(S)]
handles = []
N = wl["n_ants"]
rs = np.random.RandomState(1)
rot = torch.from_numpy((rs.randint(0, 3, size=(16, E, N)) - 1).astype(np.int8)).cuda()
ph = torch.from_numpy(rs.randint(0, 3, size=(16, E, N)).astype(np.int8)).cuda()
for s in range(S):
    with torch.cuda.stream(streams[s]):
        b = BatchedAnts(gen.cfg, per, evap_mode="lazy", record="compact", env_id_base=s * per)
        b.import_state(stack_states(states[s * per:(s + 1) * per], "all"))
        b.activate_all_pheromones(np.ones((per, N, 2)) * 10.0)
        b.observe()
        handles.append(b)
torch.cuda.synchronize()
def one_step(t):
    for s, b in enumerate(handles):
        with torch.cuda.stream(streams[s]):
            b.step(rot[t % 16, s * per:(s + 1) * per].contiguous() if False else rot[t % 16][s * per:(s + 1) * per], ph[t % 16][s * per:(s + 1) * per])
    for s, b in enumerate(handles):
        with torch.cuda.stream(streams[s]):
            b.update(None)
# note: BatchedAnts binds the stream that was current at construction
for t in range(WARM): one_step(t)
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(K): one_step(t)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("streams=%d envs=%d: %.4f ms/step, %.3e ant-steps/s" % (S, E, dt / K * 1e3, E * N * K / dt))

import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from scenarios import make_scenario
from parity_util import stack_init
from oracle.antsrl_oracle import OracleEnv
from antsrl_b200 import BatchedAnts

scen = [make_scenario(seed=1100 + e, w=64, h=64, n_ants=24, steps=10) for e in range(2)]
cfg = scen[0][0]
orc = [OracleEnv(c, i) for c, i, _ in scen]
b = BatchedAnts(cfg, 2)
b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
b.observe(); [o.observation() for o in orc]
for t in range(3):
    rot = np.stack([s[2]["rot"][t] for s in scen]); ph = np.stack([s[2]["ph"][t] for s in scen])
    ref = [o.step(rot[e].astype(np.int64), ph[e].astype(np.int64)) for e, o in enumerate(orc)]
    obs, ast, rew, done = b.step(torch.from_numpy(rot).cuda(), torch.from_numpy(ph).cuda())
    obs = obs.cpu().numpy().astype(np.float64)
    ro = np.stack([r[0] for r in ref])
    bad = np.abs(obs - ro) > 1e-5
    print("t", t, "obs mismatches per channel", bad.sum(axis=(0, 1, 2, 3)), cfg["channels"])
    print("   per ant (env0):", bad[0].sum(axis=(1, 2, 3)))
    st = b.export_state()
    for k in ("x", "y", "theta", "holding"):
        d = np.abs(st[k] - np.stack([o.s[k] for o in orc])).max()
        print("   ", k, "max abs diff", d)
    print("   reward diff", np.abs(rew.cpu().numpy() - np.stack([r[2] for r in ref])).max())
    noise = np.stack([s[2]["noise"][t] for s in scen])
    [o.update(noise[e]) for e, o in enumerate(orc)]
    b.update(torch.from_numpy(noise).cuda())
    st = b.export_state()
    for k in ("x", "y", "theta", "holding", "phero", "food"):
        d = np.abs(st[k] - np.stack([o.s[k] for o in orc])).max()
        print("   upd", k, "max abs diff", d)

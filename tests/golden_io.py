"""Load a committed reference fixture (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith(("gen_", "mainloop_", "long_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg_json"]))
    if cfg["mask"] is not None:
        cfg["mask"] = np.array(cfg["mask"], dtype=bool)
    cfg["reward_factors"] = tuple(cfg["reward_factors"])
    init, tape, rec = {}, {}, {}
    for k in z.files:
        if k.startswith("init_"):
            init[k[5:]] = z[k]
        elif k.startswith("tape_"):
            tape[k[5:]] = z[k]
        elif k not in ("cfg_json", "scenario_json"):
            rec[k] = z[k]
    init["act_bool"] = bool(init["act_bool"])
    return cfg, init, tape, rec

"""Seeded synthetic scenarios (maps, ants, action tapes, collision-noise tapes) shared by the golden
generator, the oracle tests and the GPU parity tests.  Pure numpy, independent of the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.antsrl_oracle import make_config  # noqa: E402


def disc_union(rng, w, h, n, rmin, rmax):
    """Union of n discs (dist <= r on integer cells), the map family of map_generators.py:28-46."""
    m = np.zeros((w, h), dtype=bool)
    xs = np.arange(w)[:, None]
    ys = np.arange(h)[None, :]
    for _ in range(n):
        r = int(rng.random_sample() * (rmax - rmin) + rmin)
        r = min(r, (min(w, h) - 1) // 2)
        xc = int(rng.random_sample() * (w - 2 * r) + r)
        yc = int(rng.random_sample() * (h - 2 * r) + r)
        m |= ((xc - xs) ** 2 + (yc - ys) ** 2) ** 0.5 <= r
    return m


def make_scenario(seed, w=64, h=64, n_ants=24, n_phero=2, n_rocks=0, steps=40, n_walls=4, n_food=6,
                  wall_r=(3, 8), food_r=(3, 6), none_rot_every=0, none_ph_every=0, float_food=False,
                  act_float=True, **cfg_kw):
    """Returns (cfg, init_state, tape).  Follows the construction order of environment_generator.py:52-106
    (anthill -> walls cleared in hill -> food cleared in walls -> rocks -> ants in 0.8*radius disc) with its
    own RandomState."""
    rng = np.random.RandomState(seed)
    cfg = make_config(w, h, n_ants, n_phero=n_phero, n_rocks=n_rocks, max_time=cfg_kw.pop("max_time", steps),
                      **cfg_kw)
    m = min(w, h)
    ax = int(rng.random_sample() * w * 0.5 + w * 0.25)
    ay = int(rng.random_sample() * h * 0.5 + h * 0.25)
    ar = int(rng.random_sample() * m * 0.05 + m * 0.05)
    ar = max(ar, 2)
    xs = np.arange(w)[:, None]
    ys = np.arange(h)[None, :]
    area = ((ax - xs) ** 2 + (ay - ys) ** 2) ** 0.5 <= ar
    walls = disc_union(rng, w, h, n_walls, *wall_r)
    walls[area] = False
    food = disc_union(rng, w, h, n_food, *food_r).astype(float)
    if float_food:
        food *= np.round(rng.random_sample((w, h)) * 4, 2)
    food *= (1 - walls)
    init = {"walls": walls.astype(np.uint8), "food": food,
            "anthill_xyr": np.array([ax, ay, ar], dtype=np.int32)}
    if n_rocks > 0:
        c = rng.random_sample((n_rocks, 2))
        c[:, 0] = c[:, 0] * w * 0.5 + w * 0.25
        c[:, 1] = c[:, 1] * h * 0.5 + h * 0.25
        init["rock_centers"] = c
        init["rock_radii"] = rng.random_sample(n_rocks) * 3 + 2
        init["rock_weights"] = rng.random_sample(n_rocks) * 50 + 50
    ang = rng.random_sample(n_ants) * 2 * np.pi
    dist = rng.random_sample(n_ants) * ar * 0.8
    init["x"] = np.mod(np.cos(ang) * dist + ax, w)
    init["y"] = np.mod(np.sin(ang) * dist + ay, h)
    init["theta"] = rng.random_sample(n_ants) * 2 * np.pi
    init["seed"] = rng.random_sample(n_ants)
    init["act_bool"] = not act_float
    if act_float:   # what agent.initialize does (collect_agent.py:100-102)
        init["activation"] = np.ones((n_ants, n_phero)) * 10.0
    rot = rng.randint(0, 3, size=(steps, n_ants)).astype(np.int8) - 1
    ph = rng.randint(0, 3, size=(steps, n_ants)).astype(np.int8)
    noise = rng.random_sample((steps, n_ants))
    tape = {"rot": rot, "ph": ph, "noise": noise,
            "rot_none": np.array([none_rot_every > 0 and (t % none_rot_every) == none_rot_every - 1
                                  for t in range(steps)]),
            "ph_none": np.array([none_ph_every > 0 and (t % none_ph_every) == none_ph_every - 1
                                 for t in range(steps)])}
    return cfg, init, tape


# Named scenario table: (name, kwargs).  Small enough that the oracle replays each in well under a second.
GOLDEN_SCENARIOS = [
    ("default_small", dict(seed=11, w=64, h=64, n_ants=24, steps=60)),
    ("default_200", dict(seed=1000, w=200, h=200, n_ants=50, steps=120, n_walls=10, n_food=20,
                         wall_r=(5, 15), food_r=(5, 10))),
    ("crowded", dict(seed=12, w=32, h=32, n_ants=160, steps=50, n_walls=2, n_food=8)),          # Q1 duplicates
    ("bool_activation", dict(seed=13, w=48, h=40, n_ants=20, steps=40, act_float=False)),
    ("rocks", dict(seed=14, w=64, h=64, n_ants=48, n_rocks=6, steps=60)),
    ("rect_float_food", dict(seed=15, w=40, h=72, n_ants=30, steps=50, float_food=True)),
    ("none_actions", dict(seed=16, w=48, h=48, n_ants=20, steps=40, none_rot_every=3, none_ph_every=4)),
    ("diffuse", dict(seed=17, w=48, h=48, n_ants=30, steps=40, diffuse_factor=0.02, evap_factor=0.01)),
    ("explore_reward", dict(seed=18, w=48, h=48, n_ants=20, steps=30, reward_kind="explore")),
    ("food_reward", dict(seed=19, w=48, h=48, n_ants=40, steps=40, reward_kind="food", n_food=10)),
    ("no_mask_r2", dict(seed=20, w=48, h=48, n_ants=16, steps=30, radius=2, mask=None, fwd_delta=0)),
    ("channel_order", dict(seed=21, w=48, h=48, n_ants=30, steps=40,
                           channels=["food", "walls", "phero1", "anthill", "ants", "phero0"])),
    ("fast_backward", dict(seed=22, w=48, h=48, n_ants=40, steps=60, carry_speed_reduction=0.3,
                           max_speed=1.7, n_food=12)),
]


def conservation_scenario(steps=150):
    """A scenario in which food is conserved EXACTLY (food plane + carried + delivered == initial amount after every
    step): 64 ants start on a 32-cell lattice of a 256x256 map, so no two of them act on the food of the same cell in
    the same step (the reference's last-writer scatter, quirk Q1, would otherwise duplicate or lose units).  That
    property of this seed is checked on the oracle by tests/test_oracle_golden.py::test_conservation_scenario."""
    cfg, init, tape = make_scenario(seed=4242, w=256, h=256, n_ants=64, steps=steps, n_walls=10, n_food=60,
                                    wall_r=(5, 12), food_r=(6, 12))
    gx, gy = np.meshgrid(16 + 32 * np.arange(8), 16 + 32 * np.arange(8), indexing="ij")
    x, y = gx.reshape(-1).astype(float) + 0.5, gy.reshape(-1).astype(float) + 0.5
    walls = init["walls"].astype(bool)
    for k in range(64):                      # lattice points inside walls slide along x to the next free cell
        while walls[int(x[k]) % 256, int(y[k])]:
            x[k] = (x[k] + 1.0) % 256
    init["x"], init["y"] = x, y
    return cfg, init, tape


def food_total(state):
    """food plane + carried + delivered, for one env's state dict or a batched one."""
    return (np.asarray(state["food"], dtype=float).sum() + np.asarray(state["holding"], dtype=float).sum() +
            np.asarray(state["anthill_food"], dtype=float).sum())

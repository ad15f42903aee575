"""Host-side map interface: the reference's EnvironmentGenerator / CirclesGenerator semantics
(generator/environment_generator.py:19-106, generator/map_generators.py:28-46) producing plain state dicts that
`BatchedAnts.import_state` uploads.  Runs once per episode on the host; nothing here is on the step path.

Seeding follows the reference exactly (`random.seed(seed); np.random.seed(seed * 5)`, then the same draws in the
same order from the *global* generators), so the same seed yields the same anthill, walls, food, rocks and ants
as the reference, also with user-supplied duck-typed generators that draw from the global RNGs themselves."""
import random

import numpy as np

from .batch import DEFAULT_MASK, make_config


def disc_area(w, h, cx, cy, r):
    """bool (w, h): ((cx - x)^2 + (cy - y)^2)^0.5 <= r on integers -- anthill.py:28-33 without the Python loop
    (exact: all quantities are integers, so the comparison can be done on squares)."""
    xs = np.arange(w, dtype=np.int64)[:, None]
    ys = np.arange(h, dtype=np.int64)[None, :]
    return (cx - xs) ** 2 + (cy - ys) ** 2 <= int(r) * int(r) if r >= 0 else np.zeros((w, h), dtype=bool)


class CirclesGenerator:
    """map_generators.py:28-46: union of n discs; three `random.random()` draws per disc (radius, xc, yc)."""

    def __init__(self, n_circles, min_radius, max_radius):
        self.n_circles = n_circles
        self.min_radius = min_radius
        self.max_radius = max_radius
        self._stamps = {}

    def _stamp(self, r):
        if r not in self._stamps:
            d = np.arange(-r, r + 1, dtype=np.int64)
            self._stamps[r] = d[:, None] ** 2 + d[None, :] ** 2 <= r * r
        return self._stamps[r]

    def generate(self, w, h):
        gen = np.zeros((w, h), dtype=bool)
        for _ in range(self.n_circles):
            radius = int(random.random() * (self.max_radius - self.min_radius) + self.min_radius)
            xc = int(random.random() * (w - 2 * radius) + radius)
            yc = int(random.random() * (h - 2 * radius) + radius)
            # the reference indexes gen[x, y] for x in [xc - r, xc + r]; numpy wraps negative indices
            xi = np.arange(xc - radius, xc + radius + 1)
            yi = np.arange(yc - radius, yc + radius + 1)
            if xi[0] >= 0 and xi[-1] < w and yi[0] >= 0 and yi[-1] < h:
                gen[xi[0]:xi[-1] + 1, yi[0]:yi[-1] + 1] |= self._stamp(radius)
            else:   # degenerate maps smaller than a disc: same wrap-around / IndexError behaviour as numpy
                st = self._stamp(radius)
                for a, x in enumerate(xi):
                    for b, y in enumerate(yi):
                        if st[a, b]:
                            gen[x, y] = True
        return gen


from .perlin import PerlinGenerator   # noqa: E402,F401  (map_generators.py:9-25; parity unpinned, see perlin.py)


def generate_state(w, h, n_ants, n_pheromones, n_rocks, food_generator, walls_generator, seed=None, max_hold=5,
                   draw_ant_seed=True):
    """environment_generator.py:52-99 -> one env's initial state dict (shared schema, no env axis)."""
    if seed is not None:
        random.seed(seed)
        np.random.seed(seed * 5)
    m = min(w, h)
    ax = int(random.random() * w * 0.5 + w * 0.25)
    ay = int(random.random() * h * 0.5 + h * 0.25)
    ar = int(random.random() * m * 0.05 + m * 0.05)
    area = disc_area(w, h, ax, ay, ar)
    walls = np.asarray(walls_generator.generate(w, h)).copy()
    walls[area] = False
    walls = walls.astype(bool)
    food = np.asarray(food_generator.generate(w, h)).astype(float)
    food *= (1 - walls)
    st = {"walls": walls.astype(np.uint8), "food": food, "anthill_xyr": np.array([ax, ay, ar], dtype=np.int32)}
    if n_rocks > 0:   # environment_generator.py:76-85 with the evident intent `self.n_rocks` (Q17)
        c = np.random.random((n_rocks, 2))
        c[:, 0] *= w * 0.75
        c[:, 1] *= h * 0.25
        c[:, 0] += w * 0.25
        c[:, 1] += h * 0.25
        st["rock_centers"] = c
        st["rock_radii"] = np.random.random(n_rocks) * 5 + 5
        st["rock_weights"] = np.random.random(n_rocks) * 50 + 50
    ang = np.random.random(n_ants) * 2 * np.pi
    dist = np.random.random(n_ants) * ar * 0.8
    x = np.cos(ang) * dist + ax
    y = np.sin(ang) * dist + ay
    t = np.random.random(n_ants) * 2 * np.pi
    st["x"] = np.mod(x, w)                      # Ants.__init__ -> warp_xy, ants.py:28
    st["y"] = np.mod(y, h)
    st["theta"] = t
    if draw_ant_seed:
        st["seed"] = np.random.random(n_ants)   # ants.py:41
    st["activation"] = np.zeros((n_ants, n_pheromones))
    st["act_bool"] = True                       # ants.py:83
    st["rw_alias"] = True
    st["timestep"] = 1
    return st


def stack_states(states, reward_kind="all"):
    """list of per-env dicts -> batched dict with a leading env axis (+ All_Rewards' initial distance)."""
    out = {}
    for k in states[0]:
        if k in ("act_bool", "rw_alias", "timestep"):
            out[k] = states[0][k]
        else:
            out[k] = np.stack([np.asarray(s[k]) for s in states])
    if reward_kind == "all" and "rw_prev_dist" not in out:      # reward_custom.py:77
        ax = out["anthill_xyr"][:, 0:1].astype(float)
        ay = out["anthill_xyr"][:, 1:2].astype(float)
        out["rw_prev_dist"] = ((out["x"] - ax) ** 2 + (out["y"] - ay) ** 2) ** 0.5
    return out


class BatchedEnvironmentGenerator:
    """E environments, env e generated exactly like the reference generator with seed = seed_base + e."""

    def __init__(self, w, h, n_ants, n_pheromones, n_rocks, food_generator, walls_generator, max_steps,
                 seed_base=1000, perception_mask=None, perception_shift=4, **cfg_kw):
        self.w, self.h, self.n_ants, self.n_pheromones, self.n_rocks = w, h, n_ants, n_pheromones, n_rocks
        self.food_generator, self.walls_generator = food_generator, walls_generator
        self.max_steps, self.seed_base = max_steps, seed_base
        mask = DEFAULT_MASK.copy() if perception_mask is None else np.asarray(perception_mask).astype(bool)
        self.cfg = make_config(w, h, n_ants, n_phero=n_pheromones, n_rocks=n_rocks, max_time=max_steps,
                               radius=mask.shape[0] // 2, mask=mask, fwd_delta=perception_shift, **cfg_kw)

    def generate_states(self, n_envs, first_env=0):
        return [generate_state(self.w, self.h, self.n_ants, self.n_pheromones, self.n_rocks, self.food_generator,
                               self.walls_generator, seed=self.seed_base + first_env + e) for e in range(n_envs)]

    def generate(self, n_envs, device=0, first_env=0, float_activation=True, **batch_kw):
        """-> BatchedAnts holding envs [first_env, first_env + n_envs) (global ids, for sharding).  Defaults to the
        fast field representation the configuration allows: lazy decay and, for one or two pheromones without
        diffusion, the compact 8-byte cell records."""
        from .batch import BatchedAnts
        batch_kw.setdefault("evap_mode", "lazy")
        if self.cfg.get("diffuse_factor", 0.0) == 0.0 and 1 <= self.n_pheromones <= 2 and batch_kw["evap_mode"] == "lazy":
            batch_kw.setdefault("record", "compact8")
        states = self.generate_states(n_envs, first_env)
        batch = BatchedAnts(self.cfg, n_envs, device=device, env_id_base=first_env, **batch_kw)
        batch.import_state(stack_states(states, self.cfg["reward_kind"]))
        if float_activation and self.n_pheromones > 0:   # what agent.initialize does, collect_agent.py:100-102
            batch.activate_all_pheromones(np.ones((n_envs, self.n_ants, self.n_pheromones)) * 10.0)
        return batch

"""Harness that runs the UNMODIFIED reference (SelennLamson/AntsRL) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imported by ``tests/`` (golden generation, live-reference checks), by ``bench.py --impl reference`` and by bench's
``cpu_baseline`` leg; never by the product path.  It copies no reference source; it imports the reference's own
modules and drives them through their public API.  Where the reference comes from, in this order:
  1. ``$ANTSRL_REFERENCE`` if set,
  2. the checkout ``/root/reference`` (build container),
  3. ``oracle/_ref/reference.zip`` -- the same modules byte-compiled from that checkout by the committed recipe
     ``oracle/build_ref.py`` (an archive of sourceless ``.pyc``, git-ignored, shipped to the GPU box with the snapshot).

Shims applied from outside the reference tree:
  1. ``noise`` / ``matplotlib`` stubs (imported at ``utils.py:2-3``, ``anthill.py:2-3``, ``food.py:2-3``,
     unused on the step path).
  2. the name ``np`` inside module ``environment.walls`` only is replaced by a proxy whose
     ``random.random(n)`` returns the taped collision noise of the colliding ants (``walls.py:28``), so the
     hidden global-RNG draw becomes an explicit input shared with the oracle and the CUDA path.
  3. rocks are built directly with ``CircleObstacles`` because ``environment_generator.py:83-84`` raises
     ``NameError`` (bare ``n_rocks``).
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
COMPILED_ROOT = os.path.join(_HERE, "_ref", "reference.zip")


def _find_reference():
    env = os.environ.get("ANTSRL_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return COMPILED_ROOT


REFERENCE_ROOT = _find_reference()


def reference_available():
    """True if the reference can be imported here (checkout or the byte-compiled oracle/_ref)."""
    if os.path.isfile(REFERENCE_ROOT):                  # the byte-compiled archive: usable by the interpreter that wrote it
        try:
            import importlib.util, json
            with open(os.path.join(os.path.dirname(REFERENCE_ROOT), "MANIFEST.json")) as f:
                return json.load(f).get("magic") == importlib.util.MAGIC_NUMBER.hex()
        except Exception:
            return False
    return os.path.exists(os.path.join(REFERENCE_ROOT, "environment", "RL_api.py"))


def reference_kind():
    return "compiled (oracle/_ref)" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(COMPILED_ROOT) else "checkout"


class _WallsNumpyProxy:
    """Stands in for ``np`` inside ``environment.walls`` (walls.py:22-30)."""

    class _Random:
        def __init__(self, outer):
            self._outer = outer

        def random(self, n):
            o = self._outer
            if o.noise_row is None:          # untaped: the reference's own draw from the global RNG (walls.py:28)
                return np.random.random(n)
            mask = o.last_mask
            assert mask is not None and int(mask.sum()) == int(n)
            vals = o.noise_row[mask]
            o.hits += int(n)
            return vals

    def __init__(self):
        self.noise_row = None   # (N,) f64: noise of every ant for the current update
        self.last_mask = None
        self.hits = 0
        self.random = _WallsNumpyProxy._Random(self)

    def sum(self, a, *args, **kw):
        # walls.py:28 calls np.sum(colliding_ants) right before drawing; remember the mask.
        arr = np.asarray(a)
        if arr.dtype == bool and arr.ndim == 1:
            self.last_mask = arr
        return np.sum(a, *args, **kw)

    def __getattr__(self, name):
        return getattr(np, name)


_loaded = None


def load_reference():
    """Import the reference packages (environment, generator, utils) and return a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not found at %s (build oracle/_ref with `python oracle/build_ref.py` where "
                           "/root/reference exists)" % REFERENCE_ROOT)
    for name in ("noise", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import environment.RL_api as m_api
    import environment.environment as m_env
    import environment.ants as m_ants
    import environment.pheromone as m_phero
    import environment.walls as m_walls
    import environment.food as m_food
    import environment.anthill as m_anthill
    import environment.circle_obstacles as m_rocks
    import environment.rewards.reward_custom as m_rewards
    import generator.environment_generator as m_gen
    import generator.map_generators as m_maps
    assert m_api.__file__.startswith(REFERENCE_ROOT), m_api.__file__
    proxy = _WallsNumpyProxy()
    m_walls.np = proxy
    ns = types.SimpleNamespace(api=m_api, env=m_env, ants=m_ants, phero=m_phero, walls=m_walls,
                               food=m_food, anthill=m_anthill, rocks=m_rocks, rewards=m_rewards,
                               gen=m_gen, maps=m_maps, walls_proxy=proxy)
    _loaded = ns
    return ns


def set_diffuse(ref, diffuse_factor, evap_factor):
    """Rebuild the module-level filter of pheromone.py:5-10 for another (DIFFUSE, EVAP) pair."""
    f = np.ones((3, 3)) * diffuse_factor
    f[1, 1] = 1 - 8 * diffuse_factor
    f *= 1 - evap_factor
    ref.phero.DIFFUSE_FACTOR = diffuse_factor
    ref.phero.EVAP_FACTOR = evap_factor
    ref.phero.DIFFUSE_FILTER = f


def make_reward(ref, kind, factors):
    if kind == "all":
        return ref.rewards.All_Rewards(*factors)
    if kind == "explore":
        return ref.rewards.ExplorationReward()
    if kind == "food":
        return ref.rewards.Food_Reward()
    raise ValueError(kind)


def build_env(ref, cfg, init):
    """Build a reference Environment + RLApi from explicit initial arrays (no generator).

    cfg: dict (see oracle.antsrl_oracle.make_config); init: dict of initial arrays
    (walls, food, x, y, theta, seed, anthill_xyr, rock_*).
    Object insertion order follows environment_generator.py:60-101.
    """
    reward = make_reward(ref, cfg["reward_kind"], cfg["reward_factors"])
    api = ref.api.RLApi(reward, cfg["reward_threshold"], cfg["max_speed"], cfg["max_rot_speed"],
                        cfg["carry_speed_reduction"], cfg["backward_speed_reduction"])
    env = ref.env.Environment(cfg["w"], cfg["h"], cfg["max_time"])
    ax, ay, ar = [int(v) for v in init["anthill_xyr"]]
    anthill = ref.anthill.Anthill(env, ax, ay, ar)
    walls = ref.walls.Walls(env, np.asarray(init["walls"]).astype(bool))
    food = ref.food.Food(env, np.asarray(init["food"], dtype=float))
    objs = {"anthill": anthill, "walls": walls, "food": food}
    rocks = None
    if cfg["n_rocks"] > 0:
        rocks = ref.rocks.CircleObstacles(env, centers=np.array(init["rock_centers"], dtype=float),
                                          radiuses=np.array(init["rock_radii"], dtype=float),
                                          weights=np.array(init["rock_weights"], dtype=float))
        objs["rocks"] = rocks
    n = cfg["n_ants"]
    xyt = np.stack([init["x"], init["y"], init["theta"]], axis=1).astype(float)
    ants = ref.ants.Ants(env, n, cfg["max_hold"], xyt=xyt)
    ants.seed = np.array(init["seed"], dtype=float)
    objs["ants"] = ants
    pheros = []
    for p in range(cfg["n_phero"]):
        ph = ref.phero.Pheromone(env, max_val=cfg["phero_max_val"])
        ants.register_pheromone(ph)
        pheros.append(ph)
    objs["pheros"] = pheros
    api.register_ants(ants)
    perceived = []
    for ch in cfg["channels"]:
        if ch == "ants":
            perceived.append(ants)
        elif ch.startswith("phero"):
            perceived.append(pheros[int(ch[5:])])
        elif ch == "anthill":
            perceived.append(anthill)
        elif ch == "walls":
            perceived.append(walls)
        elif ch == "food":
            perceived.append(food)
        elif ch == "rocks":
            perceived.append(rocks)
        else:
            raise ValueError(ch)
    mask = None if cfg["mask"] is None else np.asarray(cfg["mask"]).astype(bool)
    api.setup_perception(cfg["radius"], perceived, mask, cfg["fwd_delta"])
    return env, api, objs


def export_state(env, api, objs):
    """Every array of the path as plain numpy, in the shared state schema."""
    ants = objs["ants"]
    rw = api.reward
    n = ants.n_ants
    st = {
        "x": ants.ants[:, 0].copy(), "y": ants.ants[:, 1].copy(), "theta": ants.ants[:, 2].copy(),
        "prev_x": ants.prev_ants[:, 0].copy(), "prev_y": ants.prev_ants[:, 1].copy(),
        "prev_theta": ants.prev_ants[:, 2].copy(),
        "holding": np.asarray(ants.holding, dtype=float).copy(),
        "mandibles": np.asarray(ants.mandibles).astype(np.uint8),
        "reward_state": np.asarray(ants.reward_state).astype(np.uint8),
        "activation": np.asarray(ants.phero_activation, dtype=float).reshape(n, -1).copy(),
        "seed": np.asarray(ants.seed, dtype=float).copy(),
        "phero": np.stack([p.phero for p in objs["pheros"]]).astype(float) if objs["pheros"]
        else np.zeros((0, env.w, env.h)),
        "food": objs["food"].qte.astype(float).copy(),
        "walls": objs["walls"].map.astype(np.uint8),
        "anthill_xyr": np.array([objs["anthill"].x, objs["anthill"].y, objs["anthill"].radius], dtype=np.int32),
        "anthill_food": np.float64(objs["anthill"].food),
        "timestep": np.int64(env.timestep),
    }
    if "rocks" in objs:
        r = objs["rocks"]
        st["rock_centers"] = r.centers.astype(float).copy()
        st["rock_radii"] = r.radiuses.astype(float).copy()
        st["rock_weights"] = r.weights.astype(float).copy()
    else:
        st["rock_centers"] = np.zeros((0, 2)); st["rock_radii"] = np.zeros(0); st["rock_weights"] = np.zeros(0)
    st["rewards"] = np.asarray(rw.rewards, dtype=float).copy()
    if hasattr(rw, "explored_map") and rw.explored_map is not None:
        st["explored"] = rw.explored_map.astype(np.uint8)
    else:
        st["explored"] = np.zeros((env.w, env.h), dtype=np.uint8)
    st["rw_holding_prev"] = (np.asarray(rw.ants_holding, dtype=float).copy()
                             if getattr(rw, "ants_holding", None) is not None else np.zeros(n))
    st["rw_prev_dist"] = (np.asarray(rw.previous_dist, dtype=float).copy()
                          if getattr(rw, "previous_dist", None) is not None else np.zeros(n))
    return st


def run_update(ref, env, noise_row):
    """env.update() (environment.py:42-47) with the collision noise of this tick taped."""
    ref.walls_proxy.noise_row = None if noise_row is None else np.asarray(noise_row, dtype=float)
    ref.walls_proxy.last_mask = None
    env.update()

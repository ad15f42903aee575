"""Glue between the reference-shaped host objects (Environment, Ants, Pheromone, ...) and one antsrl_b200 handle
(a batch of E = 1 environments).  Created lazily on the first observation / step / update; from then on the device
holds the authoritative state and the host objects' arrays are refreshed on demand (``pull``)."""
import numpy as np


def _classes():
    from .ants import Ants
    from .pheromone import Pheromone
    from .walls import Walls
    from .food import Food
    from .anthill import Anthill
    from .circle_obstacles import CircleObstacles
    from .RL_api import RLApi
    return Ants, Pheromone, Walls, Food, Anthill, CircleObstacles, RLApi


class DeviceBridge:
    def __init__(self, env):
        from antsrl_b200.batch import BatchedAnts, make_config
        from . import pheromone as phero_mod
        from .rewards.reward_custom import All_Rewards, ExplorationReward, Food_Reward
        Ants, Pheromone, Walls, Food, Anthill, CircleObstacles, RLApi = _classes()
        self.env = env
        pick = lambda cls: [o for o in env.objects if isinstance(o, cls)]
        ants_l, walls_l, food_l, hill_l = pick(Ants), pick(Walls), pick(Food), pick(Anthill)
        if len(ants_l) != 1 or len(walls_l) != 1 or len(food_l) != 1 or len(hill_l) != 1:
            raise ValueError("the CUDA backend needs exactly one Ants, Walls, Food and Anthill object per Environment")
        self.ants, self.walls, self.food, self.hill = ants_l[0], walls_l[0], food_l[0], hill_l[0]
        rocks_l, api_l = pick(CircleObstacles), pick(RLApi)
        self.rocks = rocks_l[0] if rocks_l else None
        self.api = api_l[0] if api_l else None
        self.pheros = list(self.ants.pheromones)
        n, P = self.ants.n_ants, len(self.pheros)
        R = self.rocks.n_obst if self.rocks is not None else 0
        max_vals = {ph.max_val for ph in self.pheros}
        if len(max_vals) > 1:
            raise ValueError("all pheromones must share one max_val on the CUDA backend")
        max_val = max_vals.pop() if max_vals else 255.0
        kw = dict(n_phero=P, n_rocks=R, max_time=env.max_time, phero_max_val=max_val, max_hold=self.ants.max_hold,
                  diffuse_factor=phero_mod.DIFFUSE_FACTOR, evap_factor=phero_mod.EVAP_FACTOR)
        api = self.api
        if api is not None:
            channels = []
            for obj in api.perceived_objects:                     # RL_api.py:123-142
                if isinstance(obj, Pheromone):
                    channels.append("phero%d" % self.pheros.index(obj))
                elif isinstance(obj, Food):
                    channels.append("food")
                elif isinstance(obj, Walls):
                    channels.append("walls")
                elif isinstance(obj, Anthill):
                    channels.append("anthill")
                elif isinstance(obj, CircleObstacles):
                    channels.append("rocks")
                elif isinstance(obj, Ants):
                    channels.append("ants")
                else:
                    raise ValueError("unsupported perceived object %r" % (obj,))
            rw = api.reward
            if isinstance(rw, All_Rewards):
                kind, factors = "all", (rw.fct_explore, rw.fct_food, rw.fct_anthill, rw.fct_explore_holding,
                                        rw.fct_headinganthill)
            elif isinstance(rw, ExplorationReward):
                kind, factors = "explore", (0, 0, 0, 0, 0)
            elif isinstance(rw, Food_Reward):
                kind, factors = "food", (0, 0, 0, 0, 0)
            else:
                raise NotImplementedError("only All_Rewards, ExplorationReward and Food_Reward run on the CUDA backend "
                                          "(custom Python Reward subclasses would need a host round trip per step)")
            kw.update(radius=api.perception_radius, mask=api.perception_mask, fwd_delta=api.perception_fwd_delta,
                      channels=channels or ["walls"], reward_kind=kind, reward_factors=factors,
                      reward_threshold=api.reward_threshold, max_speed=api.max_speed, max_rot_speed=api.max_rot_speed,
                      carry_speed_reduction=api.carry_speed_reduction,
                      backward_speed_reduction=api.backward_speed_reduction)
        else:
            kw.update(radius=0, mask=None, fwd_delta=0, channels=["walls"], reward_kind="food")
        self.cfg = make_config(env.w, env.h, n, **kw)
        self.batch = BatchedAnts(self.cfg, 1)
        self.version = 0          # bumped by every device operation
        self.pulled = -1
        self.push()

    # ------------------------------------------------------------------ host -> device
    def push(self):
        a, env = self.ants, self.env
        st = {"x": a._ants[None, :, 0], "y": a._ants[None, :, 1], "theta": a._ants[None, :, 2],
              "prev_x": a._prev_ants[None, :, 0], "prev_y": a._prev_ants[None, :, 1],
              "prev_theta": a._prev_ants[None, :, 2],
              "holding": np.asarray(a._holding, dtype=float)[None], "seed": np.asarray(a._seed, dtype=float)[None],
              "mandibles": np.asarray(a._mandibles).astype(np.uint8)[None],
              "reward_state": np.asarray(a._reward_state).astype(np.uint8)[None],
              "activation": np.asarray(a._phero_activation, dtype=float).reshape(1, a.n_ants, -1),
              "walls": self.walls.map.astype(np.uint8)[None], "food": np.asarray(self.food._qte, dtype=float)[None],
              "anthill_xyr": np.array([[self.hill.x, self.hill.y, self.hill.radius]], dtype=np.int32),
              "anthill_food": np.array([float(self.hill._food)]),
              "timestep": env._timestep,
              "act_bool": bool(np.asarray(a._phero_activation).dtype == bool)}
        if self.pheros:
            st["phero"] = np.stack([np.asarray(ph._phero, dtype=float) for ph in self.pheros])[None]
        if self.rocks is not None:
            st["rock_centers"] = np.asarray(self.rocks._centers, dtype=float)[None]
            st["rock_radii"] = np.asarray(self.rocks.radiuses, dtype=float)[None]
            st["rock_weights"] = np.asarray(self.rocks.weights, dtype=float)[None]
        rw = self.api.reward if self.api is not None else None
        st["rw_alias"] = True
        if rw is not None:
            st["rw_alias"] = bool(getattr(rw, "_aliased", True))
            if getattr(rw, "_explored_map", None) is not None:
                st["explored"] = np.asarray(rw._explored_map).astype(np.uint8)[None]
            if getattr(rw, "_ants_holding", None) is not None:
                st["rw_holding_prev"] = np.asarray(rw._ants_holding, dtype=float)[None]
            if getattr(rw, "_previous_dist", None) is not None:
                st["rw_prev_dist"] = np.asarray(rw._previous_dist, dtype=float)[None]
            if getattr(rw, "_rewards", None) is not None:
                st["rewards"] = np.asarray(rw._rewards, dtype=float)[None]
        self.batch.import_state(st)
        self.version += 1
        self.pulled = self.version      # host and device agree

    # ------------------------------------------------------------------ device -> host
    def pull(self):
        if self.pulled == self.version:
            return
        st = self.batch.export_state()
        a = self.ants
        a._ants = np.stack([st["x"][0], st["y"][0], st["theta"][0]], axis=1)
        a._prev_ants = np.stack([st["prev_x"][0], st["prev_y"][0], st["prev_theta"][0]], axis=1)
        a._holding = st["holding"][0]
        a._mandibles = st["mandibles"][0].astype(np.int64)
        a._reward_state = st["reward_state"][0]
        act = st["activation"][0]
        a._phero_activation = (act != 0) if st["act_bool"] else act
        for k, ph in enumerate(self.pheros):
            ph._phero = st["phero"][0, k]
        self.food._qte = st["food"][0]
        self.hill._food = np.float64(st["anthill_food"][0])       # anthill.py:46: int 0 + np.sum(...)
        if self.rocks is not None:
            self.rocks._centers = st["rock_centers"][0]
        rw = self.api.reward if self.api is not None else None
        if rw is not None:
            rw._rewards = st["rewards"][0]
            rw._aliased = st["rw_alias"]
            if hasattr(rw, "_explored_map"):
                rw._explored_map = st["explored"][0].astype(bool)
            if hasattr(rw, "_ants_holding"):
                rw._ants_holding = st["rw_holding_prev"][0]
            if hasattr(rw, "_previous_dist"):
                rw._previous_dist = st["rw_prev_dist"][0]
        self.env._timestep = st["timestep"]
        self.pulled = self.version

    # ------------------------------------------------------------------ the step loop
    def observe(self):
        obs, ast, state, rew = self.batch.observe_host()
        self.version += 1
        return obs[0].astype(np.float64), ast[0].astype(np.float64), state[0].astype(np.float64), rew[0].copy()

    def step(self, rotation, pheromone):
        n = self.ants.n_ants

        def as_rot(a):
            # RL_api.py:191 multiplies whatever array it gets by max_rot_speed; the kernels take whole turns in int8
            # (what the agents produce: argmax / randint minus n // 2).  Anything else is refused, not truncated.
            if a is None:
                return None
            a = np.asarray(a)
            if a.shape != (n,):
                raise ValueError("rotation must have shape (%d,)" % n)
            r = np.rint(a)
            if not (np.array_equal(r, a) and np.all(np.abs(r) <= 127)):
                raise ValueError("rotation must hold whole numbers in [-127, 127] (units of max_rot_speed)")
            return np.ascontiguousarray(r.astype(np.int8)[None])

        def as_ph(a):
            # ants.py:89-96: 0 -> off, 1 -> first pheromone, ANY other value -> second pheromone
            if a is None:
                return None
            a = np.asarray(a)
            if a.shape != (n,):
                raise ValueError("on_off_pheromones must have shape (%d,)" % n)
            return np.ascontiguousarray(np.where(a == 0, 0, np.where(a == 1, 1, 2)).astype(np.int8)[None])
        obs, ast, rew, done = self.batch.step_host(as_rot(rotation), as_ph(pheromone))
        self.version += 1
        return obs[0].astype(np.float64), ast[0].astype(np.float64), rew[0].copy(), done

    def update(self):
        mode = self.env.collision_noise
        if isinstance(mode, str) and mode == "philox":
            self.batch.update_host(None)
        else:
            # walls.py:24-28: one draw per colliding ant, in ant order, from the global numpy RNG (or the callable)
            st = self.batch.export_state(keys=("x", "y"))
            w, h = self.env.w, self.env.h
            cx = st["x"][0].astype(int); cy = st["y"][0].astype(int)
            cx[cx >= w] -= w; cy[cy >= h] -= h
            hit = self.walls.map[cx, cy]
            noise = np.zeros(self.ants.n_ants)
            nh = int(np.sum(hit))
            if nh:
                noise[hit] = np.random.random(nh) if mode is None else np.asarray(mode(hit), dtype=float)
            self.batch.update_host(noise[None])
        self.version += 1

    def activate_all_pheromones(self, new_activations):
        a = np.asarray(new_activations)
        self.batch.activate_all_pheromones(a.reshape(1, self.ants.n_ants, -1))
        self.version += 1

    def close(self):
        self.batch.close()

"""environment/ants.py of the reference (lines 8-144).  State arrays keep the reference's names and shapes; the
per-step mutators of the reference (rotate / forward / update_mandibles / emit_pheromones / give_reward / update)
run inside the CUDA step and update kernels and are not re-implemented on the host."""
import numpy as np
from typing import List

from .environment import Environment, EnvObject
from .pheromone import Pheromone


class AntsVisualization(EnvObject):
    def __init__(self, env, ants_xyt, mandibles, holding, reward_state):
        super().__init__(env)
        self.ants = ants_xyt.copy()
        self.mandibles = mandibles.copy()
        self.holding = holding.copy()
        self.reward_state = reward_state.copy()


def _device_only(name):
    def f(self, *a, **k):
        raise NotImplementedError("Ants.%s runs inside the CUDA step/update kernels of antsrl_b200 "
                                  "(use RLApi.step / Environment.update)" % name)
    f.__name__ = name
    return f


class Ants(EnvObject):
    _MIRRORS = ("ants", "prev_ants", "mandibles", "holding", "reward_state", "phero_activation", "seed")

    def __init__(self, environment: Environment, n_ants: int, max_hold, xyt=None):
        super().__init__(environment)
        self.n_ants = n_ants
        self.max_hold = max_hold
        self._ants = np.array(xyt, dtype=float).copy()
        self._ants[:, 0] = np.mod(self._ants[:, 0], self.environment.w)      # warp_xy, ants.py:69-71
        self._ants[:, 1] = np.mod(self._ants[:, 1], self.environment.h)
        self._prev_ants = self._ants.copy()
        self._phero_activation = np.zeros((n_ants, 0))
        self.pheromones: List[Pheromone] = []
        self._mandibles = np.zeros(n_ants, dtype=bool)
        self._holding = np.zeros(n_ants)
        self._reward_state = np.zeros(n_ants, dtype=np.uint8)
        self._seed = np.random.random(n_ants)                                # ants.py:41

    def visualize_copy(self, newenv):
        return AntsVisualization(newenv, self.ants, self.mandibles, self.holding, self.reward_state)

    # --- state arrays: refreshed from the device on access
    def _get(self, name):
        self._pull()
        return getattr(self, name)

    ants = property(lambda s: s._get("_ants"), lambda s, v: setattr(s, "_ants", v))
    prev_ants = property(lambda s: s._get("_prev_ants"), lambda s, v: setattr(s, "_prev_ants", v))
    mandibles = property(lambda s: s._get("_mandibles"), lambda s, v: setattr(s, "_mandibles", v))
    holding = property(lambda s: s._get("_holding"), lambda s, v: setattr(s, "_holding", v))
    reward_state = property(lambda s: s._get("_reward_state"), lambda s, v: setattr(s, "_reward_state", v))
    phero_activation = property(lambda s: s._get("_phero_activation"), lambda s, v: setattr(s, "_phero_activation", v))
    seed = property(lambda s: s._get("_seed"), lambda s, v: setattr(s, "_seed", v))

    @property
    def x(self):
        return self.ants[:, 0]

    @property
    def y(self):
        return self.ants[:, 1]

    @property
    def xy(self):
        return self.ants[:, 0:2]

    @property
    def theta(self):
        return self.ants[:, 2]

    def register_pheromone(self, pheromone: Pheromone):
        if self.environment._bridge is not None:
            raise RuntimeError("pheromones must be registered before the first observation / step / update")
        self._phero_activation = np.hstack([self._phero_activation, np.zeros((self.n_ants, 1))]).astype(bool)   # ants.py:83
        self.pheromones.append(pheromone)

    def activate_all_pheromones(self, new_activations):
        """ants.py:86-87."""
        self._phero_activation = np.array(new_activations, copy=True)
        if self.environment._bridge is not None:
            self.environment._bridge.activate_all_pheromones(self._phero_activation)

    def update_step(self):
        return 999

    warp_theta = _device_only("warp_theta")
    rotate_ants = _device_only("rotate_ants")
    warp_xy = _device_only("warp_xy")
    translate_ants = _device_only("translate_ants")
    forward_ants = _device_only("forward_ants")
    activate_pheromone = _device_only("activate_pheromone")
    emit_pheromones = _device_only("emit_pheromones")
    update_mandibles = _device_only("update_mandibles")
    give_reward = _device_only("give_reward")
    apply_func = _device_only("apply_func")

"""environment/rewards/reward.py of the reference (lines 6-45)."""
from abc import ABC
import numpy as np


class Reward(ABC):
    def __init__(self):
        self.ants = None
        self.environment = None
        self._rewards = None
        self._aliased = True

    @property
    def rewards(self):
        return self._rewards

    @rewards.setter
    def rewards(self, v):
        self._rewards = v

    def setup(self, ants):
        self.ants = ants
        self.environment = ants.environment
        self._rewards = np.zeros(self.ants.n_ants, dtype=float)
        self._aliased = True

    # pickled under the reference's attribute names (see EnvObject.__getstate__)
    _MIRRORS = ("rewards", "explored_map", "previous_dist", "ants_holding")

    def __getstate__(self):
        if self.ants is not None:
            self.ants._pull()
        d = dict(self.__dict__)
        for name in self._MIRRORS:
            if "_" + name in d:
                d[name] = d.pop("_" + name)
        return d

    def __setstate__(self, d):
        d = dict(d)
        for name in self._MIRRORS:
            if name in d:
                d["_" + name] = d.pop(name)
        d.setdefault("_aliased", False)
        self.__dict__.update(d)

    def observation(self, obs_coords, perception, agent_state):
        pass

    def step(self, done, turn_index, open_close_mandibles, on_off_pheromones):
        return self.rewards

    def visualization(self):
        return None

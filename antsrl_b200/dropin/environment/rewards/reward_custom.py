"""environment/rewards/reward_custom.py of the reference (lines 8-109).  The three shipped rewards are computed in
the epilogue of the CUDA perception kernel; these classes carry their parameters and mirror their state."""
import numpy as np

from environment.anthill import Anthill
from environment.rewards.reward import Reward


class _ExploredMixin:
    @property
    def explored_map(self):
        if self.ants is not None:
            self.ants._pull()
        return self._explored_map

    @explored_map.setter
    def explored_map(self, v):
        self._explored_map = v

    def visualization(self):
        m = self.explored_map
        return None if m is None else m.copy()


def _mirrored(name):
    """Public attribute of the reference (reward_custom.py:30,52-60) kept as _<name> and refreshed from the device."""
    def get(self):
        if self.ants is not None:
            self.ants._pull()
        return getattr(self, "_" + name)

    def set_(self, v):
        setattr(self, "_" + name, v)
    return property(get, set_)


class ExplorationReward(_ExploredMixin, Reward):
    def __init__(self):
        super(ExplorationReward, self).__init__()
        self._explored_map = None

    def setup(self, ants):
        super(ExplorationReward, self).setup(ants)
        self._explored_map = np.zeros((self.environment.w, self.environment.h), dtype=bool)


class Food_Reward(Reward):
    ants_holding = _mirrored("ants_holding")

    def __init__(self):
        super(Food_Reward, self).__init__()
        self._ants_holding = None

    def setup(self, ants):
        super(Food_Reward, self).setup(ants)
        self._ants_holding = np.zeros(ants.n_ants)


class All_Rewards(_ExploredMixin, Reward):
    ants_holding = _mirrored("ants_holding")
    previous_dist = _mirrored("previous_dist")

    def __init__(self, fct_explore=1, fct_food=1, fct_anthill=5, fct_explore_holding=0, fct_headinganthill=1):
        super(All_Rewards, self).__init__()
        self._explored_map = None
        self.fct_explore = fct_explore
        self.fct_food = fct_food
        self.fct_anthill = fct_anthill
        self.fct_explore_holding = fct_explore_holding
        self.fct_headinganthill = fct_headinganthill
        self._previous_dist = None
        self.anthill_x = 0
        self.anthill_y = 0
        self._ants_holding = None

    def compute_distance(self, x, y):
        return ((x - self.anthill_x) ** 2 + (y - self.anthill_y) ** 2) ** 0.5

    def setup(self, ants):
        super(All_Rewards, self).setup(ants)
        self._ants_holding = np.zeros(ants.n_ants)
        self._explored_map = np.zeros((self.environment.w, self.environment.h), dtype=bool)
        for obj in ants.environment.objects:
            if isinstance(obj, Anthill):
                self.anthill_x = obj.x
                self.anthill_y = obj.y
        self._previous_dist = self.compute_distance(ants._ants[:, 0], ants._ants[:, 1])    # reward_custom.py:77

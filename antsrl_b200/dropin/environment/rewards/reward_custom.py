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


class ExplorationReward(_ExploredMixin, Reward):
    def __init__(self):
        super(ExplorationReward, self).__init__()
        self._explored_map = None

    def setup(self, ants):
        super(ExplorationReward, self).setup(ants)
        self._explored_map = np.zeros((self.environment.w, self.environment.h), dtype=bool)


class Food_Reward(Reward):
    def __init__(self):
        super(Food_Reward, self).__init__()
        self._ants_holding = None

    def setup(self, ants):
        super(Food_Reward, self).setup(ants)
        self._ants_holding = np.zeros(ants.n_ants)


class All_Rewards(_ExploredMixin, Reward):
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

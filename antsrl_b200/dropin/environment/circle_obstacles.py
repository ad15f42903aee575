"""environment/circle_obstacles.py of the reference (lines 8-61)."""
import numpy as np

from .environment import Environment, EnvObject
from utils import AX  # noqa: F401  (the reference does `from utils import *`)


class CircleObstaclesVisualization(EnvObject):
    def __init__(self, env, centers, radiuses, weights):
        super().__init__(env)
        self.centers = centers.copy()
        self.radiuses = radiuses.copy()
        self.weights = weights.copy()


class CircleObstacles(EnvObject):
    _MIRRORS = ("centers",)

    def __init__(self, environment: Environment, centers, radiuses, weights):
        super().__init__(environment)
        self.w = environment.w
        self.h = environment.h
        self.n_obst = len(radiuses)
        self._centers = np.asarray(centers, dtype=float)
        self.radiuses = np.asarray(radiuses, dtype=float)
        self.weights = np.asarray(weights, dtype=float)
        self.crossed_radiuses = self.radiuses[None, :] + self.radiuses[:, None]
        self.crossed_weights = self.weights[None, :] / (self.weights[:, None] + self.weights[None, :])

    @property
    def centers(self):
        self._pull()
        return self._centers

    @centers.setter
    def centers(self, v):
        self._centers = v

    def visualize_copy(self, newenv):
        return CircleObstaclesVisualization(newenv, self.centers, self.radiuses, self.weights)

    def update_step(self):
        return 0

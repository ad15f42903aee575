"""environment/anthill.py of the reference (lines 7-46).  The disc is analytic on the device; ``area`` is provided
for callers that read it."""
import numpy as np

from .environment import Environment, EnvObject


class AnthillVisualization(EnvObject):
    def __init__(self, env, x, y, radius, food):
        super().__init__(env)
        self.x = x
        self.y = y
        self.radius = radius
        self.food = food


class Anthill(EnvObject):
    _MIRRORS = ("food",)

    def __init__(self, environment: Environment, x, y, radius):
        super().__init__(environment)
        self.w = environment.w
        self.h = environment.h
        self.x = x
        self.y = y
        self.radius = radius
        self._food = 0
        xs = np.arange(self.w, dtype=np.int64)[:, None]
        ys = np.arange(self.h, dtype=np.int64)[None, :]
        # anthill.py:28-33 (((x0-x)^2 + (y0-y)^2)^0.5 <= r) on integers == comparison of squares
        self.area = ((self.x - xs) ** 2 + (self.y - ys) ** 2 <= int(radius) ** 2) if radius >= 0 else np.zeros((self.w, self.h), bool)

    @property
    def food(self):
        self._pull()
        return self._food

    @food.setter
    def food(self, v):
        self._food = v

    def visualize_copy(self, newenv):
        return AnthillVisualization(newenv, self.x, self.y, self.radius, self.food)

    def update_step(self):
        return 1000

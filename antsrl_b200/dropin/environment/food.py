"""environment/food.py of the reference (lines 7-19)."""
import numpy as np

from .environment import Environment, EnvObject


class FoodVisualization(EnvObject):
    def __init__(self, environment: Environment, qte):
        super().__init__(environment)
        self.qte = qte.astype(np.uint8)


class Food(EnvObject):
    _MIRRORS = ("qte",)

    def __init__(self, environment: Environment, qte):
        super().__init__(environment)
        self._qte = qte.astype(float)

    @property
    def qte(self):
        self._pull()
        return self._qte

    @qte.setter
    def qte(self, v):
        self._qte = v

    def visualize_copy(self, newenv):
        return FoodVisualization(newenv, self.qte)

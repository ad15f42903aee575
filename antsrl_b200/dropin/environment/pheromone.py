"""environment/pheromone.py of the reference (lines 5-45).  DIFFUSE_FACTOR / EVAP_FACTOR are read when the
environment moves to the device; deposit and evaporation / diffusion run in the CUDA update."""
import numpy as np

from .environment import Environment, EnvObject

DIFFUSE_FACTOR = 0
EVAP_FACTOR = 0.001



def diffuse_filter(diffuse_factor, evap_factor):
    """The 3x3 stencil of pheromone.py:7-10: ring = factor, centre = 1 - 8 factor, all scaled by (1 - evaporation)."""
    f = np.full((3, 3), float(diffuse_factor))
    f[1, 1] = 1 - 8 * diffuse_factor
    return f * (1 - evap_factor)


DIFFUSE_FILTER = diffuse_filter(DIFFUSE_FACTOR, EVAP_FACTOR)    # kept for callers that read it; the device gets the factors


class PheromoneVisualization(EnvObject):
    def __init__(self, environment: Environment, color, max_val, phero):
        super().__init__(environment)
        self.color = color
        self.max_val = max_val
        self.phero = phero.astype(np.uint8)


class Pheromone(EnvObject):
    _MIRRORS = ("phero",)

    def __init__(self, environment: Environment, color=(64, 64, 64), max_val=None, phero=None):
        super().__init__(environment)
        self.color = color
        self.max_val = max_val
        self.w = environment.w
        self.h = environment.h
        self._phero = np.zeros((self.w, self.h)) if phero is None else np.array(phero, dtype=float).copy()

    @property
    def phero(self):
        self._pull()
        return self._phero

    @phero.setter
    def phero(self, v):
        self._phero = v

    def visualize_copy(self, newenv):
        return PheromoneVisualization(newenv, self.color, self.max_val, self.phero)

    def add_pheromones(self, add_xy, phero_strength):
        raise NotImplementedError("Pheromone.add_pheromones runs inside the CUDA update of antsrl_b200")

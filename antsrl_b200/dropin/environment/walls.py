"""environment/walls.py of the reference (lines 7-30): static bool map; collisions run in the CUDA update."""
from .environment import Environment, EnvObject


class Walls(EnvObject):
    def __init__(self, environment: Environment, map_in):
        super().__init__(environment)
        self.w = environment.w
        self.h = environment.h
        self.map = map_in.astype(bool)

    def visualize_copy(self, newenv):
        return self

    def update_step(self):
        return -1

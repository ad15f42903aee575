"""generator/environment_generator.py of the reference (lines 10-106): same constructor, ``setup_perception`` and
``generate(rl_api) -> Environment``; same RNG seeding and draw order, so a seed yields the same world as the
reference (tests/test_generator_golden.py).  The world itself is drawn by ``antsrl_b200.generator.generate_state``
(shared with the batched generator); this class wraps it into the reference's objects.  With ``n_rocks > 0`` the
evident intent of lines 82-84 (``self.n_rocks``) is implemented instead of the reference's NameError (quirk Q17)."""
import numpy as np

from antsrl_b200 import DEFAULT_MASK
from antsrl_b200.generator import generate_state
from antsrl_b200.snapshot import PHERO_COLORS            # environment_generator.py:13-17  # noqa: F401
from environment.environment import Environment
from environment.anthill import Anthill
from environment.ants import Ants
from environment.pheromone import Pheromone
from environment.circle_obstacles import CircleObstacles
from environment.walls import Walls
from environment.food import Food
from environment.RL_api import RLApi


class EnvironmentGenerator:
    def __init__(self, w, h, n_ants, n_pheromones, n_rocks, food_generator, walls_generator, max_steps, seed=None):
        self.w, self.h, self.n_ants = w, h, n_ants
        self.n_pheromones, self.n_rocks = n_pheromones, n_rocks
        self.food_generator, self.walls_generator = food_generator, walls_generator
        self.max_steps, self.seed = max_steps, seed
        self.perception_mask = DEFAULT_MASK.copy()           # environment_generator.py:35-41
        self.perception_shift = 4

    def setup_perception(self, new_mask, new_shift):
        self.perception_mask, self.perception_shift = new_mask, new_shift

    def generate(self, rl_api: RLApi):
        st = generate_state(self.w, self.h, self.n_ants, self.n_pheromones, self.n_rocks, self.food_generator,
                            self.walls_generator, seed=self.seed, draw_ant_seed=False)
        env = Environment(self.w, self.h, self.max_steps)
        # object order of environment.objects = construction order of the reference (lines 60-101): anthill, walls,
        # food, rocks, ants, pheromones, RL api -- Environment.update sorts by update_step but ties keep this order
        statics = [Anthill(env, *(int(v) for v in st["anthill_xyr"])),
                   Walls(env, st["walls"].astype(bool)),
                   Food(env, st["food"])]
        if self.n_rocks > 0:
            statics.append(CircleObstacles(env, centers=st["rock_centers"], radiuses=st["rock_radii"],
                                           weights=st["rock_weights"]))
        # Ants.__init__ draws Ants.seed from np.random right after the position draws, like ants.py:41
        ants = Ants(env, self.n_ants, 5, xyt=np.column_stack((st["x"], st["y"], st["theta"])))
        pheros = [Pheromone(env, color=PHERO_COLORS[p % len(PHERO_COLORS)], max_val=255)
                  for p in range(self.n_pheromones)]
        for ph in pheros:
            ants.register_pheromone(ph)
        rl_api.register_ants(ants)
        # perceived_objects as the reference's insert(0, ants) / insert(p + 1, phero) leave it (lines 94-99)
        rl_api.setup_perception(self.perception_mask.shape[0] // 2, [ants] + pheros + statics, self.perception_mask,
                                self.perception_shift)
        return env

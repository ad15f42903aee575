"""generator/environment_generator.py of the reference (lines 10-106): same constructor, ``setup_perception`` and
``generate(rl_api) -> Environment``; same RNG seeding and draw order, so a seed yields the same world as the
reference (tests/test_generator_golden.py).  With ``n_rocks > 0`` the evident intent of lines 82-84
(``self.n_rocks``) is implemented instead of the reference's NameError (quirk Q17)."""
import numpy as np

from antsrl_b200.generator import generate_state
from environment.environment import Environment
from environment.anthill import Anthill
from environment.ants import Ants
from environment.pheromone import Pheromone
from environment.circle_obstacles import CircleObstacles
from environment.walls import Walls
from environment.food import Food
from environment.RL_api import RLApi

PHERO_COLORS = [
    (255, 64, 0),
    (64, 64, 255),
    (100, 255, 100)
]


class EnvironmentGenerator:
    def __init__(self, w, h, n_ants, n_pheromones, n_rocks, food_generator, walls_generator, max_steps, seed=None):
        self.w = w
        self.h = h
        self.n_ants = n_ants
        self.n_pheromones = n_pheromones
        self.n_rocks = n_rocks
        self.food_generator = food_generator
        self.walls_generator = walls_generator
        self.perception_mask = np.array([[0, 0, 1, 1, 1, 0, 0],
                                         [0, 1, 1, 1, 1, 1, 0],
                                         [1, 1, 1, 1, 1, 1, 1],
                                         [1, 1, 1, 1, 1, 1, 1],
                                         [1, 1, 1, 1, 1, 1, 1],
                                         [0, 1, 1, 1, 1, 1, 0],
                                         [0, 0, 1, 1, 1, 0, 0]], dtype=bool)
        self.perception_shift = 4
        self.max_steps = max_steps
        self.seed = seed

    def setup_perception(self, new_mask, new_shift):
        self.perception_mask = new_mask
        self.perception_shift = new_shift

    def generate(self, rl_api: RLApi):
        st = generate_state(self.w, self.h, self.n_ants, self.n_pheromones, self.n_rocks, self.food_generator,
                            self.walls_generator, seed=self.seed, draw_ant_seed=False)
        env = Environment(self.w, self.h, self.max_steps)
        perceived_objects = []
        ax, ay, ar = [int(v) for v in st["anthill_xyr"]]
        anthill = Anthill(env, ax, ay, ar)
        perceived_objects.append(anthill)
        walls = Walls(env, st["walls"].astype(bool))
        perceived_objects.append(walls)
        food = Food(env, st["food"])
        perceived_objects.append(food)
        if self.n_rocks > 0:
            rocks = CircleObstacles(env, centers=st["rock_centers"], radiuses=st["rock_radii"], weights=st["rock_weights"])
            perceived_objects.append(rocks)
        xyt = np.array([st["x"], st["y"], st["theta"]]).T
        ants = Ants(env, self.n_ants, 5, xyt=xyt)        # draws Ants.seed from np.random like ants.py:41
        perceived_objects.insert(0, ants)
        for p in range(self.n_pheromones):
            phero = Pheromone(env, color=PHERO_COLORS[p % len(PHERO_COLORS)], max_val=255)
            ants.register_pheromone(phero)
            perceived_objects.insert(p + 1, phero)
        rl_api.register_ants(ants)
        rl_api.setup_perception(self.perception_mask.shape[0] // 2, perceived_objects, self.perception_mask,
                                self.perception_shift)
        return env

"""Snapshots of one environment of a device-resident batch in the reference's visualisation format.

The reference records an episode as ``states.append(env.save_state())`` after every ``env.update()`` and pickles the
list (main.py:136-147); ``Environment.save_state`` (environment.py:36-40) asks every object for its
``visualize_copy`` (ants.py:43-44, pheromone.py:33-34, food.py:17-18, anthill.py:35-36, circle_obstacles.py:29-30,
RL_api.py:53-54, walls.py:16-17) and ``gui/visualize.py`` replays the file.  The drop-in ``Environment`` does this
for its single environment; this module does it for environment ``e`` of a :class:`antsrl_b200.BatchedAnts` batch
without moving the other E - 1 environments off the GPU (``ants_export_env_state`` of the C ABI).

The objects are built from whichever ``environment`` package is importable -- the reference checkout or
``antsrl_b200.dropin_path()`` (same module and class names) -- so the pickle names exactly the classes the viewer
imports.
"""
import pickle
import sys

import numpy as np

PHERO_COLORS = [(255, 64, 0), (64, 64, 255), (100, 255, 100)]          # environment_generator.py:13-17

_SNAPSHOT_KEYS = ("x", "y", "theta", "mandibles", "holding", "reward_state", "phero", "food", "explored",
                  "anthill_xyr", "anthill_food", "rock_centers", "rock_radii", "rock_weights")


def _modules():
    """The `environment` package on sys.path (the reference's, or the drop-in one as a default)."""
    try:
        import environment.environment  # noqa: F401
    except ImportError:
        from . import dropin_path
        sys.path.insert(0, dropin_path())
    import environment.environment as m_env
    import environment.ants as m_ants
    import environment.pheromone as m_phero
    import environment.walls as m_walls
    import environment.food as m_food
    import environment.anthill as m_hill
    import environment.circle_obstacles as m_rocks
    import environment.RL_api as m_api
    return m_env, m_ants, m_phero, m_walls, m_food, m_hill, m_rocks, m_api


def make_walls(batch, env_index):
    """A `Walls` object (walls.py:7-17) of environment ``env_index``.  The viewer keys its background on the identity
    of this object (visualize.py:175-178), so build it once per episode and pass it to every :func:`snapshot_env`."""
    m_env, _, _, m_walls, *_ = _modules()
    st = batch.export_state(keys=("walls",), envs=(env_index, 1))
    holder = m_env.Environment(batch.W, batch.H, batch.cfg["max_time"])
    return m_walls.Walls(holder, st["walls"][0].astype(bool))


def snapshot_env(batch, env_index, walls=None, phero_colors=None, heatmap=True):
    """`Environment.save_state()` (environment.py:36-40) of environment ``env_index`` of a BatchedAnts batch: a new
    `Environment` holding the visualisation copies in the generator's object order (environment_generator.py:60-101:
    anthill, walls, food, rocks, ants, pheromones, RL api).  ``heatmap``: RLVisualization carries the exploration map
    like `All_Rewards.visualization` / `ExplorationReward.visualization` (reward_custom.py:24-25,108-109); False =
    `Reward.visualization`'s None (reward.py:40-45, Food_Reward)."""
    m_env, m_ants, m_phero, m_walls, m_food, m_hill, m_rocks, m_api = _modules()
    keys = tuple(k for k in _SNAPSHOT_KEYS if heatmap or k != "explored")
    st = batch.export_state(keys=keys, envs=(env_index, 1))
    if walls is None:
        walls = make_walls(batch, env_index)
    colors = PHERO_COLORS if phero_colors is None else list(phero_colors)
    max_val = batch.cfg["phero_max_val"]
    env = m_env.Environment(batch.W, batch.H, batch.cfg["max_time"])       # timestep stays 1, as in save_state
    # every constructor registers itself with `env` and save_state adds the copy once more (environment.py:8-9,39):
    # each visualisation object appears twice in `objects`, Walls (returned as is) once -- as in the reference's files
    ax, ay, ar = (int(v) for v in st["anthill_xyr"][0])
    env.add_object(m_hill.AnthillVisualization(env, ax, ay, ar, np.float64(st["anthill_food"][0])))
    env.add_object(walls)
    env.add_object(m_food.FoodVisualization(env, st["food"][0]))
    if batch.R > 0:
        env.add_object(m_rocks.CircleObstaclesVisualization(env, st["rock_centers"][0], st["rock_radii"][0],
                                                            st["rock_weights"][0]))
    xyt = np.stack([st["x"][0], st["y"][0], st["theta"][0]], axis=1)
    env.add_object(m_ants.AntsVisualization(env, xyt, st["mandibles"][0].astype(np.int64), st["holding"][0],
                                            st["reward_state"][0]))
    for k in range(batch.P):
        env.add_object(m_phero.PheromoneVisualization(env, colors[k % len(colors)], max_val, st["phero"][0, k]))
    env.add_object(m_api.RLVisualization(env, st["explored"][0].astype(bool) if heatmap else None))
    return env


def save_episode(states, path, append=False):
    """main.py:139-147: the file holds ONE pickled list; later episodes are appended by re-writing
    ``previous_states + states``."""
    previous = []
    if append:
        with open(path, "rb") as f:
            previous = pickle.load(f)
    with open(path, "wb") as f:
        pickle.dump(previous + list(states), f)


def load_episode(path):
    """The list of saved environments (gui/visualize.py:125); needs an `environment` package on sys.path."""
    _modules()
    with open(path, "rb") as f:
        return pickle.load(f)


class EpisodeRecorder:
    """Records environment ``env_index`` of a batch the way main.py records its environment: call :meth:`record`
    after every `update()`, :meth:`save` at the end of the episode."""

    def __init__(self, batch, env_index=0, phero_colors=None, heatmap=True):
        self.batch, self.env_index = batch, int(env_index)
        self.phero_colors, self.heatmap = phero_colors, heatmap
        self.walls = None
        self.states = []

    def new_episode(self):
        """Call after importing a new map: the next snapshot gets a new Walls object."""
        self.walls = None

    def record(self):
        if self.walls is None:
            self.walls = make_walls(self.batch, self.env_index)
        self.states.append(snapshot_env(self.batch, self.env_index, self.walls, self.phero_colors, self.heatmap))
        return self.states[-1]

    def save(self, path, append=False):
        save_episode(self.states, path, append)
        self.states = []

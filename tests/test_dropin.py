"""The drop-in layer (antsrl_b200/dropin: `environment`, `generator`, `utils` with the reference's names): API
surface on the CPU, and on the GPU the reference's own main.py loop reproduced through it."""
import inspect
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# The drop-in packages are called `environment` / `generator` / `utils` like the reference's, so they are exercised
# in a subprocess with antsrl_b200.dropin_path() first on sys.path (exactly how a user switches over).
PRELUDE = """
import sys, json
sys.path.insert(0, %r)
import antsrl_b200
sys.path.insert(0, antsrl_b200.dropin_path())
import numpy as np
from environment.RL_api import RLApi
from environment.environment import Environment, EnvObject
from environment.ants import Ants
from environment.pheromone import Pheromone
from environment.walls import Walls
from environment.food import Food
from environment.anthill import Anthill
from environment.circle_obstacles import CircleObstacles
from environment.rewards.reward import Reward
from environment.rewards.reward_custom import All_Rewards, ExplorationReward, Food_Reward
from generator.environment_generator import EnvironmentGenerator
from generator.map_generators import CirclesGenerator, PerlinGenerator
from utils import AX
""" % ROOT


def run_snippet(body):
    code = PRELUDE + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def test_api_surface_matches_reference():
    """Names, argument lists and generator output of SURVEY.md section 8-b, without touching a GPU."""
    out = run_snippet("""
        import inspect
        sig = lambda f: list(inspect.signature(f).parameters)
        assert sig(RLApi.__init__) == ['self', 'reward', 'reward_threshold', 'max_speed', 'max_rot_speed',
                                       'carry_speed_reduction', 'backward_speed_reduction']
        assert sig(RLApi.setup_perception) == ['self', 'radius', 'objects', 'mask', 'forward_delta']
        assert sig(RLApi.step) == ['self', 'rotation', 'on_off_pheromones']
        assert sig(RLApi.observation) == ['self'] and sig(RLApi.register_ants) == ['self', 'new_ants']
        assert sig(EnvironmentGenerator.__init__) == ['self', 'w', 'h', 'n_ants', 'n_pheromones', 'n_rocks',
                                                      'food_generator', 'walls_generator', 'max_steps', 'seed']
        assert sig(All_Rewards.__init__) == ['self', 'fct_explore', 'fct_food', 'fct_anthill', 'fct_explore_holding',
                                             'fct_headinganthill']
        assert sig(Environment.__init__) == ['self', 'w', 'h', 'max_time']
        assert sig(Ants.__init__) == ['self', 'environment', 'n_ants', 'max_hold', 'xyt']
        assert sig(Reward.step) == ['self', 'done', 'turn_index', 'open_close_mandibles', 'on_off_pheromones']
        api = RLApi(All_Rewards(1, 2, 10, 1, 3), 1, 1, 40 / 180 * np.pi, 0.05, 0.5)
        g = EnvironmentGenerator(200, 200, 50, 2, 0, CirclesGenerator(20, 5, 10), CirclesGenerator(10, 5, 15), 100, seed=1000)
        env = g.generate(api)
        z = np.load(%r)
        assert np.array_equal(api.ants.ants[:, 0], z['state_x']) and np.array_equal(api.ants.ants[:, 2], z['state_theta'])
        assert np.array_equal(api.ants.seed, z['state_seed'])
        assert np.array_equal(env.objects[1].map.astype(np.uint8), z['state_walls'])
        assert np.array_equal(env.objects[2].qte, z['state_food'])
        # what agents read (agents/agent.py:22-25, collect_agent.py:101-102)
        assert api.perception_coords.shape == (7, 7, 2) and len(api.perceived_objects) == 6 and api.ants.n_ants == 50
        assert [type(o).__name__ for o in api.perceived_objects] == list(z['channels'])
        assert sum(isinstance(o, Pheromone) for o in api.perceived_objects) == 2
        assert [type(o).__name__ for o in env.objects] == ['Anthill', 'Walls', 'Food', 'Ants', 'Pheromone', 'Pheromone', 'RLApi']
        assert env.timestep == 1 and env.max_time == 100 and api.ants.phero_activation.dtype == bool
        api.ants.activate_all_pheromones(np.ones((50, 2)) * 10)
        assert api.ants.phero_activation.dtype == float
        import random
        random.seed(3)
        walls = PerlinGenerator().generate(40, 24)      # restated pnoise2 (parity unpinned): shape / dtype / determinism
        random.seed(3)
        assert walls.shape == (40, 24) and walls.dtype == bool and np.array_equal(walls, PerlinGenerator().generate(40, 24))
        print('surface ok')
    """ % os.path.join(GOLDEN, "gen_200_s1000.npz"))
    assert "surface ok" in out


@pytest.mark.gpu
def test_mainloop_matches_reference_with_global_rng():
    """main.py's loop (generate -> initialize -> observation -> 80 x [step; update]) through the drop-in classes with
    the reference's own collision-noise source (global numpy RNG, one draw per colliding ant) against the recording
    of the unmodified reference (tests/golden/mainloop_s1000.npz)."""
    out = run_snippet("""
        z = np.load(%r)
        api = RLApi(All_Rewards(1, 2, 10, 1, 3), 1, 1, 40 / 180 * np.pi, 0.05, 0.5)
        g = EnvironmentGenerator(200, 200, 50, 2, 0, CirclesGenerator(20, 5, 10), CirclesGenerator(10, 5, 15), 80, seed=1000)
        env = g.generate(api)
        api.ants.activate_all_pheromones(np.ones((50, 2)) * 10)
        np.random.seed(777)
        act = np.random.RandomState(4242)
        close = lambda a, b, w: np.testing.assert_allclose(np.asarray(a, float), np.asarray(b, float), rtol=1e-5, atol=1e-7, err_msg=w)
        obs0, as0, st0 = api.observation()
        assert obs0.shape == (50, 7, 7, 6) and obs0.dtype == np.float64 and st0.shape == (50, 4)
        close(obs0, z['obs0'], 'obs0'); close(as0, z['agent_state0'], 'as0'); close(st0, z['state0'], 'state0')
        anthill = [o for o in env.objects if isinstance(o, Anthill)][0]
        for t in range(80):
            rot = act.randint(0, 3, 50) - 1
            ph = act.randint(0, 3, 50)
            obs, ast, rew, done = api.step(rot, ph)
            env.update()
            close(obs, z['obs'][t], 'obs %%d' %% t); close(ast, z['agent_state'][t], 'as %%d' %% t)
            close(rew, z['reward'][t], 'reward %%d' %% t)
            assert bool(done) == bool(z['done'][t])
            xyt = api.ants.ants
            assert np.array_equal(xyt[:, :2].astype(int), z['xyt'][t][:, :2].astype(int)), 'ant cells %%d' %% t
            close(xyt, z['xyt'][t], 'xyt %%d' %% t)
            assert np.array_equal(api.ants.holding, z['holding'][t])
            assert anthill.food == z['anthill_food'][t]
        close(np.stack([p.phero for p in api.ants.pheromones]), z['final_phero'], 'phero')
        close([o for o in env.objects if isinstance(o, Food)][0].qte, z['final_food'], 'food')
        assert np.array_equal(api.reward.explored_map.astype(np.uint8), z['final_explored'])
        assert env.timestep == 81
        # the global RNG was consumed exactly like the reference consumed it
        assert np.random.random() == float(z['next_global_draw'])
        snap = env.save_state()       # visualisation copies (environment.py:36-40)
        names = [type(o).__name__ for o in snap.objects]
        assert {'AnthillVisualization', 'Walls', 'FoodVisualization', 'AntsVisualization', 'PheromoneVisualization', 'RLVisualization'} <= set(names), names
        print('mainloop ok')
    """ % os.path.join(GOLDEN, "mainloop_s1000.npz"))
    assert "mainloop ok" in out

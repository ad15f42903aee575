"""The oracle against the LIVE reference (build container only: needs /root/reference, skipped elsewhere) on exactly
the configurations the GPU parity tests compare the CUDA path with the oracle on -- the odd configurations (radius 1
and 5, 0 / 1 / 3 pheromones, tiny maps, diffusion over several tiles, rocks with a custom channel list) and the seeded
random configurations of tests/test_gpu_parity.py, with the same per-env seeds.  Together with the committed
fixtures this pins the oracle wherever the GPU tests lean on it."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.environ.get("ANTSRL_REFERENCE", "/root/reference")),
                                reason="needs the reference checkout (build container only)")

from test_gpu_parity import ODD_CONFIGS, _random_config, _variants      # noqa: E402
from test_oracle_golden import check_oracle_against_record              # noqa: E402


def _check(scenarios):
    import make_golden
    for cfg, init, tape in scenarios:
        rec = make_golden.run_reference(cfg, init, tape)
        check_oracle_against_record(cfg, init, tape, rec)


@pytest.mark.parametrize("name,kw", ODD_CONFIGS, ids=[n for n, _ in ODD_CONFIGS])
def test_odd_configurations_against_the_reference(name, kw):
    _check(_variants(kw, 3))


@pytest.mark.parametrize("case", range(24))
def test_random_configurations_against_the_reference(case):
    _check(_variants(_random_config(np.random.RandomState(9000 + case)), 2))


def test_crowded_rocks_and_long_run_against_the_reference():
    """tests/test_gpu_parity.py::test_crowded_rocks_default_channels and ::test_long_run_crosses_generation_folds."""
    _check(_variants(dict(seed=970, w=40, h=40, n_ants=64, n_rocks=14, steps=40, n_walls=2, n_food=5), 3))
    _check(_variants(dict(seed=41, w=64, h=56, n_ants=48, n_rocks=3, steps=300, n_walls=5, n_food=8), 2))


def test_full_size_generated_env_against_the_reference():
    """BASELINE configs[3] dimensions (tests/test_gpu_parity.py::test_full_size_env_matches_oracle): a 1024x1024 map
    with 1024 ants and 64 rocks from the drop-in generator, the reference's objects built from the same state, 6 steps
    of the reference loop against the oracle."""
    import make_golden
    from antsrl_b200.generator import BatchedEnvironmentGenerator, CirclesGenerator
    gen = BatchedEnvironmentGenerator(1024, 1024, 1024, 2, 64, CirclesGenerator(320, 5, 10), CirclesGenerator(400, 5, 15),
                                      max_steps=100, seed_base=1000)
    init = gen.generate_states(1, 0)[0]
    init = dict(init, act_bool=False, activation=np.ones((1024, 2)) * 10.0)      # agent.initialize
    rs = np.random.RandomState(5)
    T = 6
    tape = {"rot": (rs.randint(0, 3, (T, 1024)) - 1).astype(np.int8), "ph": rs.randint(0, 3, (T, 1024)).astype(np.int8),
            "noise": rs.random_sample((T, 1024)), "rot_none": np.zeros(T, bool), "ph_none": np.zeros(T, bool)}
    rec = make_golden.run_reference(gen.cfg, init, tape)
    check_oracle_against_record(gen.cfg, init, tape, rec)

"""The oracle against the LIVE reference (the checkout in the build container, oracle/_ref -- the same modules
byte-compiled by oracle/build_ref.py -- on the GPU box; skipped only where neither exists) on exactly
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

sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as _rh                                    # noqa: E402

pytestmark = pytest.mark.skipif(not _rh.reference_available(),
                                reason="needs the reference (checkout or oracle/_ref built by oracle/build_ref.py)")

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


@pytest.mark.parametrize("n_rocks", [0, 5])
def test_irregular_call_order_against_the_reference(n_rocks):
    """tests/test_gpu_parity.py::test_irregular_call_order: observation / step / update in an order main.py never
    uses, the reference and the oracle driven by the same call string, outputs and full state compared after every
    call."""
    import ref_harness
    from scenarios import make_scenario
    from oracle.antsrl_oracle import OracleEnv
    from test_oracle_golden import _close
    ref = ref_harness.load_reference()
    for e in range(3):
        cfg, init, tape = make_scenario(seed=900 + e, w=56, h=48, n_ants=40, n_rocks=n_rocks, steps=16, n_walls=6, n_food=8)
        ref_harness.set_diffuse(ref, cfg["diffuse_factor"], cfg["evap_factor"])
        env, api, objs = ref_harness.build_env(ref, cfg, init)
        objs["ants"].activate_all_pheromones(np.asarray(init["activation"], dtype=float))
        o = OracleEnv(cfg, init)
        t = 0
        for k, op in enumerate("ousussuuosuosusuuussu"):
            what = "env %d call %d (%s)" % (e, k, op)
            if op == "o":
                want, got = api.observation(), o.observation()
                for a, b_ in zip(got, want):
                    _close(a, b_, what)
            elif op == "s":
                rot, ph = tape["rot"][t].astype(np.int64), tape["ph"][t].astype(np.int64)
                want, got = api.step(rot, ph), o.step(rot, ph)
                for a, b_ in zip(got[:3], want[:3]):
                    _close(a, b_, what)
                assert bool(got[3]) == bool(want[3]), what
            else:
                ref_harness.run_update(ref, env, tape["noise"][t])
                o.update(tape["noise"][t])
                t += 1
            st, mine = ref_harness.export_state(env, api, objs), o.export()
            for key in ("x", "y", "theta", "prev_x", "prev_y", "prev_theta", "holding", "activation", "phero", "food",
                        "rewards", "rw_holding_prev", "rw_prev_dist", "rock_centers", "anthill_food"):
                _close(mine[key], st[key], what + ": " + key)
            for key in ("mandibles", "reward_state", "explored"):
                assert np.array_equal(np.asarray(mine[key]).astype(np.uint8), st[key]), what + ": " + key
            assert int(mine["timestep"]) == int(st["timestep"]), what

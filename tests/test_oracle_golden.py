"""Pins the oracle (oracle/antsrl_oracle.py) to the reference: every committed fixture is a trajectory
recorded from the UNMODIFIED reference (tests/golden/make_golden.py); the oracle must reproduce it."""
import numpy as np
import pytest

from golden_io import golden_names, load_golden
from oracle.antsrl_oracle import OracleEnv

RTOL = 1e-12   # same numpy build => bit-identical; tolerance only guards a different libm


def _close(a, b, what):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, what
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-12, err_msg=what)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference(name):
    check_oracle_against_record(*load_golden(name))


def check_oracle_against_record(cfg, init, tape, rec):
    """The oracle driven by (init, tape) against a recording of the reference (make_golden.run_reference's keys)."""
    env = OracleEnv(cfg, init)
    obs0, agent_state0, state0 = env.observation()
    _close(obs0, rec["obs0"], "obs0"); _close(agent_state0, rec["agent_state0"], "agent_state0")
    _close(state0, rec["state0"], "state0"); _close(env.s["rewards"], rec["reward0"], "reward0")
    T = tape["rot"].shape[0]
    for t in range(T):
        rot = None if tape["rot_none"][t] else tape["rot"][t].astype(np.int64)
        ph = None if tape["ph_none"][t] else tape["ph"][t].astype(np.int64)
        obs, agent_state, reward, done = env.step(rot, ph)
        _close(obs, rec["t_obs"][t], "obs t=%d" % t)
        _close(agent_state, rec["t_agent_state"][t], "agent_state t=%d" % t)
        _close(reward, rec["t_reward"][t], "reward t=%d" % t)
        assert done == bool(rec["t_done"][t])
        s = env.s
        ps = rec["t_post_step"][t]
        _close(np.stack([s["x"], s["y"], s["theta"], s["holding"]]), ps[:4], "post_step t=%d" % t)
        assert np.array_equal(s["mandibles"], ps[4].astype(np.uint8)), "mandibles t=%d" % t
        assert np.array_equal(s["reward_state"], ps[5].astype(np.uint8)), "reward_state(step) t=%d" % t
        env.update(tape["noise"][t])
        pu = rec["t_post_update"][t]
        _close(np.stack([s["x"], s["y"], s["theta"], s["holding"]]), pu[:4], "post_update t=%d" % t)
        assert np.array_equal(s["reward_state"], pu[5].astype(np.uint8)), "reward_state(update) t=%d" % t
        _close(s["phero"].sum(axis=(1, 2)), rec["t_phero_sum"][t], "phero_sum t=%d" % t)
        _close(s["food"].sum(), rec["t_food_sum"][t], "food_sum t=%d" % t)
        assert int(s["explored"].sum()) == int(rec["t_explored_count"][t])
        _close(s["anthill_food"], rec["t_anthill_food"][t], "anthill_food t=%d" % t)
        if cfg["n_rocks"]:
            _close(s["rock_centers"], rec["t_rock_centers"][t], "rock_centers t=%d" % t)
    fin = env.export()
    for k in ("x", "y", "theta", "prev_x", "prev_y", "prev_theta", "holding", "activation", "seed", "phero", "food",
              "rw_holding_prev", "rw_prev_dist", "rewards", "rock_centers"):
        _close(fin[k], rec["final_" + k], "final " + k)
    for k in ("mandibles", "reward_state", "walls", "explored"):
        assert np.array_equal(np.asarray(fin[k]).astype(np.uint8), rec["final_" + k]), "final " + k
    assert int(fin["timestep"]) == int(rec["final_timestep"])


def test_oracle_reproduces_1000_step_episode():
    """BASELINE.json configs[0]'s horizon: the default-sized map (200x200, 50 ants) under random actions for 1000
    steps, recorded from the unmodified reference as per-step summaries, ant snapshots every 250 steps and the full
    final state (tests/golden/make_golden.py long)."""
    import json
    import os
    import sys
    from golden_io import GOLDEN_DIR
    from scenarios import make_scenario
    sys.path.insert(0, GOLDEN_DIR)
    z = np.load(os.path.join(GOLDEN_DIR, "long_200_s1001.npz"))
    kw = json.loads(str(z["scenario_json"]))
    for k in ("wall_r", "food_r"):
        kw[k] = tuple(kw[k])
    cfg, init, tape = make_scenario(**kw)

    def summary(obs, agent_state, reward, st):       # the recorder's long_summary, restated (it imports the reference)
        cells = st["x"].astype(np.int64) * cfg["h"] + st["y"].astype(np.int64)
        weights = np.arange(1, cells.size + 1, dtype=np.int64)
        return np.array([st["x"].sum(), st["y"].sum(), st["theta"].sum(), np.asarray(reward, dtype=float).sum(),
                         st["holding"].sum(), float(st["anthill_food"]), float(np.asarray(st["explored"]).sum()),
                         st["phero"][0].sum(), st["phero"][1].sum(), st["food"].sum(),
                         np.asarray(obs, dtype=float).sum(), np.asarray(agent_state, dtype=float).sum(),
                         float((cells * weights).sum()), float(np.asarray(st["mandibles"]).astype(np.int64).sum()),
                         float(np.asarray(st["reward_state"]).astype(np.int64).sum())])

    env = OracleEnv(cfg, init)
    env.observation()
    T = tape["rot"].shape[0]
    assert T == 1000 == z["t_summary"].shape[0]
    for t in range(T):
        obs, agent_state, reward, done = env.step(tape["rot"][t].astype(np.int64), tape["ph"][t].astype(np.int64))
        env.update(tape["noise"][t])
        s = env.s
        _close(summary(obs, agent_state, reward, s), z["t_summary"][t], "summary t=%d" % t)
        if (t + 1) % 250 == 0:
            snap = z["snap%d" % (t + 1)]
            _close(np.stack([s["x"], s["y"], s["theta"], s["holding"]]), snap[:4], "snapshot t=%d" % t)
            assert np.array_equal(s["mandibles"], snap[4].astype(np.uint8))
            assert np.array_equal(s["reward_state"], snap[5].astype(np.uint8))
    assert bool(done) == bool(z["done_last"])
    fin = env.export()
    for k in ("x", "y", "theta", "holding", "phero", "food", "anthill_food", "rewards", "rw_prev_dist", "rw_holding_prev"):
        _close(fin[k], z["final_" + k], "final " + k)
    for k in ("mandibles", "reward_state", "explored"):
        assert np.array_equal(np.asarray(fin[k]).astype(np.uint8), z["final_" + k]), "final " + k
    assert int(fin["timestep"]) == int(z["final_timestep"]) == 1001


def test_conservation_scenario():
    """tests/scenarios.py::conservation_scenario is duplicate-free: in the oracle (pinned to the reference above) food
    plane + carried + delivered stays exactly the initial amount after every step and update -- the precondition of
    tests/test_gpu_bench_conditions.py::test_food_conservation_large_batch."""
    from scenarios import conservation_scenario, food_total
    cfg, init, tape = conservation_scenario(150)
    env = OracleEnv(cfg, init)
    total0 = food_total(env.s)
    assert total0 == init["food"].sum() > 0
    env.observation()
    for t in range(150):
        env.step(tape["rot"][t].astype(np.int64), tape["ph"][t].astype(np.int64))
        assert food_total(env.s) == total0, "after step %d" % t
        env.update(tape["noise"][t])
        assert food_total(env.s) == total0, "after update %d" % t
    assert env.s["holding"].sum() > 0

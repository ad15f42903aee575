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
    cfg, init, tape, rec = load_golden(name)
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

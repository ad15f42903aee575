"""GPU parity at the conditions the benchmark and BASELINE.json's configs actually run: the 8-byte cell records with
programmatic dependent launch and Philox noise at >= 131 072 ants per launch, the reference-recorded 1000-step episode
replayed in every record format, BASELINE configs[1] (1024 default-map envs, recorded action tape, 1000 steps, 32 envs
checked against the oracle at every step), exact food conservation over a large batch, and the observation-level
pheromone decay out to the end of the decay table."""
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

from golden_io import GOLDEN_DIR
from parity_util import assert_close, compare_state, stack_init
from scenarios import conservation_scenario, food_total, make_scenario

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ bench conditions
def test_bench_conditions_match_oracle():
    """cfg4 maps (1024x1024, 1024 ants, 64 rocks, 7x7x7 obs) from the drop-in generator, 128 envs = 131 072 ants per
    launch (programmatic dependent launch switches itself on), 8-byte records, lazy field, in-kernel Philox noise --
    exactly what bench.py times -- for 30 steps; envs 0, 41, 86, 127 are compared with the oracle after every step
    (obs, agent_state, reward) and in full (exported state) after steps 10, 20, 30."""
    import torch
    from antsrl_b200 import BatchedAnts
    from antsrl_b200.generator import BatchedEnvironmentGenerator, CirclesGenerator, stack_states
    from oracle.antsrl_oracle import OracleEnv, philox_uniform
    E, N, T, seed = 128, 1024, 30, 20261018
    picks = (0, 41, 86, 127)
    gen = BatchedEnvironmentGenerator(1024, 1024, N, 2, 64, CirclesGenerator(320, 5, 10), CirclesGenerator(400, 5, 15),
                                      max_steps=T + 5, seed_base=1000)
    states = gen.generate_states(E, 0)
    cfg = gen.cfg
    oracles = {e: OracleEnv(cfg, states[e]) for e in picks}
    batch = BatchedAnts(cfg, E, evap_mode="lazy", record="compact8", rng_seed=seed, env_id_base=0)
    assert E * N >= 131072                                    # the PDL threshold of ants_create
    batch.import_state(stack_states(states, "all"))
    del states
    act = np.ones((E, N, 2)) * 10.0
    batch.activate_all_pheromones(act)
    for o in oracles.values():
        o.activate_all_pheromones(act[0])
    obs, ast, st, rew = batch.observe()
    obs_h = obs.cpu().numpy()
    for e, o in oracles.items():
        ref = o.observation()
        assert_close(obs_h[e], ref[0], "obs0 env %d" % e)
    rs = np.random.RandomState(12345)
    for t in range(T):
        rot = (rs.randint(0, 3, (E, N)) - 1).astype(np.int8)
        ph = rs.randint(0, 3, (E, N)).astype(np.int8)
        obs, ast, rew, done = batch.step(torch.from_numpy(rot).cuda(), torch.from_numpy(ph).cuda())
        obs_h, ast_h, rew_h = obs.cpu().numpy(), ast.cpu().numpy(), rew.cpu().numpy()
        for e, o in oracles.items():
            r = o.step(rot[e].astype(np.int64), ph[e].astype(np.int64))
            assert_close(obs_h[e], r[0], "obs t=%d env %d" % (t, e))
            assert_close(ast_h[e], r[1], "agent_state t=%d env %d" % (t, e))
            assert_close(rew_h[e], r[2], "reward t=%d env %d" % (t, e))
            assert bool(done) == bool(r[3])
            o.update(philox_uniform(seed, e, int(o.s["timestep"]), N))
        batch.update(None)
        if (t + 1) % 10 == 0:
            for e, o in oracles.items():
                compare_state(batch.export_state(envs=(e, 1)), [o], "bench conditions t=%d env %d" % (t, e), cfg)
    batch.close()


# ------------------------------------------------------------------------------------------------ food conservation
@pytest.mark.parametrize("mode", ["compact8", "compact", "lazy", "tiles", "dense"])
def test_food_conservation_large_batch(mode):
    """64 replicas of the duplicate-free scenario (tests/scenarios.py::conservation_scenario, verified on the oracle by
    tests/test_oracle_golden.py): food plane + carried + delivered equals the initial amount after EVERY step and update
    in sampled envs, the replicas stay bit-identical, explored only grows, pheromone stays in [0, max_val] and is zero
    inside walls."""
    import torch
    from antsrl_b200 import BatchedAnts
    T, E = 150, 64
    cfg, init, tape = conservation_scenario(T)
    N = cfg["n_ants"]
    record = mode if mode.startswith("compact") else "f64"
    b = BatchedAnts(cfg, E, evap_mode="lazy" if mode.startswith("compact") else mode, record=record)
    b.import_state(stack_init(cfg, [init] * E))
    b.observe()
    total0 = food_total(dict(food=init["food"], holding=0.0, anthill_food=0.0))
    keys = ("food", "holding", "anthill_food")
    explored_prev = 0
    for t in range(T):
        rot = torch.from_numpy(np.broadcast_to(tape["rot"][t], (E, N)).copy()).cuda()
        ph = torch.from_numpy(np.broadcast_to(tape["ph"][t], (E, N)).copy()).cuda()
        noise = torch.from_numpy(np.broadcast_to(tape["noise"][t], (E, N)).copy()).cuda()
        b.step(rot, ph)
        e = (7 * t) % E
        assert food_total(b.export_state(keys=keys, envs=(e, 1))) == total0, "after step %d, env %d" % (t, e)
        b.update(noise)
        assert food_total(b.export_state(keys=keys, envs=(e, 1))) == total0, "after update %d, env %d" % (t, e)
        if t % 50 == 49:
            n_exp = int(b.export_state(keys=("explored",), envs=(e, 1))["explored"].sum())
            assert n_exp > explored_prev
            explored_prev = n_exp
    st = b.export_state()
    for k, v in st.items():
        if isinstance(v, np.ndarray) and v.size:
            assert np.array_equal(v, np.broadcast_to(v[:1], v.shape)), "replicated envs diverged: " + k
    assert food_total(st) == E * total0
    assert st["holding"].sum() > 0, "the scenario must exercise pickups"
    assert (st["food"] >= 0).all() and (st["holding"] >= 0).all()
    assert st["phero"].min() >= 0 and st["phero"].max() <= 255.0 and st["phero"].max() > 0
    assert (st["phero"][:, :, st["walls"][0].astype(bool)] == 0).all()
    b.close()


# ------------------------------------------------------------------------------------------------ 1000-step fixture
def _long_summary(obs, agent_state, reward, st, h):
    """tests/golden/make_golden.py::long_summary on one env of an exported GPU state."""
    cells = st["x"].astype(np.int64) * h + st["y"].astype(np.int64)
    weights = np.arange(1, cells.size + 1, dtype=np.int64)
    return np.array([st["x"].sum(), st["y"].sum(), st["theta"].sum(), np.asarray(reward, dtype=float).sum(),
                     st["holding"].sum(), float(st["anthill_food"]), float(np.asarray(st["explored"]).sum()),
                     st["phero"][0].sum(), st["phero"][1].sum(), st["food"].sum(), np.asarray(obs, dtype=float).sum(),
                     np.asarray(agent_state, dtype=float).sum(), float((cells * weights).sum()),
                     float(np.asarray(st["mandibles"]).astype(np.int64).sum()),
                     float(np.asarray(st["reward_state"]).astype(np.int64).sum())])


# columns of the summary that are integer-valued state (bit-exact) vs float sums (1e-5 relative)
_LONG_EXACT = (4, 5, 6, 9, 12, 13, 14)


@pytest.mark.parametrize("mode", ["compact8", "compact", "lazy", "tiles"])
def test_reference_1000_step_episode(mode):
    """BASELINE configs[0]'s horizon: the UNMODIFIED reference's 1000-step episode on the default-sized map
    (tests/golden/long_200_s1001.npz: 2138 wall hits, 198 food units delivered) replayed on the GPU in every record
    format: the per-step summary rows (ant-cell checksum, carried / delivered / remaining food, explored count,
    mandibles, reward_state bit-exact; coordinate, reward, pheromone and observation sums to 1e-5), the ant snapshots
    every 250 steps and the full final state."""
    import torch
    from antsrl_b200 import BatchedAnts
    z = np.load(os.path.join(GOLDEN_DIR, "long_200_s1001.npz"))
    kw = json.loads(str(z["scenario_json"]))
    for k in ("wall_r", "food_r"):
        kw[k] = tuple(kw[k])
    cfg, init, tape = make_scenario(**kw)
    record = mode if mode.startswith("compact") else "f64"
    b = BatchedAnts(cfg, 1, evap_mode="lazy" if mode.startswith("compact") else mode, record=record)
    b.import_state(stack_init(cfg, [init]))
    b.observe()
    T = tape["rot"].shape[0]
    assert T == 1000 == z["t_summary"].shape[0]
    done = False
    for t in range(T):
        obs, ast, rew, done = b.step(torch.from_numpy(tape["rot"][t][None]).cuda(), torch.from_numpy(tape["ph"][t][None]).cuda())
        obs_h, ast_h, rew_h = obs.cpu().numpy()[0], ast.cpu().numpy()[0], rew.cpu().numpy()[0]
        b.update(torch.from_numpy(tape["noise"][t][None]).cuda())
        st = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in b.export_state().items()}
        row, want = _long_summary(obs_h, ast_h, rew_h, st, cfg["h"]), z["t_summary"][t]
        for c in _LONG_EXACT:
            assert row[c] == want[c], "summary column %d at t=%d: %r != %r" % (c, t, row[c], want[c])
        np.testing.assert_allclose(row, want, rtol=1e-5, atol=1e-6, err_msg="summary t=%d" % t)
        if (t + 1) % 250 == 0:
            snap = z["snap%d" % (t + 1)]
            assert np.array_equal(st["x"].astype(np.int64), snap[0].astype(np.int64)), "ant cells t=%d" % t
            assert np.array_equal(st["y"].astype(np.int64), snap[1].astype(np.int64)), "ant cells t=%d" % t
            assert_close(np.stack([st["x"], st["y"], st["theta"], st["holding"]]), snap[:4], "snapshot t=%d" % t)
            assert np.array_equal(st["mandibles"], snap[4].astype(np.uint8))
            assert np.array_equal(st["reward_state"], snap[5].astype(np.uint8))
    assert bool(done) == bool(z["done_last"])
    fin = {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in b.export_state().items()}
    for k in ("x", "y", "theta", "holding", "phero", "food", "anthill_food", "rewards", "rw_prev_dist", "rw_holding_prev"):
        assert_close(fin[k], z["final_" + k], "final " + k)
    for k in ("mandibles", "reward_state", "explored"):
        assert np.array_equal(fin[k], z["final_" + k]), "final " + k
    assert int(fin["timestep"]) == int(z["final_timestep"]) == 1001
    b.close()


# ------------------------------------------------------------------------------------------------ BASELINE configs[1]
_CFG2 = {}
PLANES_EVERY = 20          # plane summaries (remaining food, explored count, pheromone mass) every 20th step


def _cfg2_oracle_worker(k):
    """Replays env picks[k] in the oracle and compares it with what the GPU recorded at EVERY step (forked after the
    GPU run: the recordings are inherited, nothing is pickled)."""
    from oracle.antsrl_oracle import OracleEnv, philox_uniform
    c = _CFG2
    e = c["picks"][k]
    o = OracleEnv(c["cfg"], c["states"][k])
    o.activate_all_pheromones(np.ones((c["N"], 2)) * 10.0)
    o.observation()
    for t in range(c["T"]):
        r = o.step(c["rot"][t, e].astype(np.int64), c["ph"][t, e].astype(np.int64))
        what = "env %d t=%d" % (e, t)
        assert_close(c["obs"][t, k], r[0], what + ": obs")
        assert_close(c["ast"][t, k], r[1], what + ": agent_state")
        assert_close(c["rew"][t, k], r[2], what + ": reward")
        o.update(philox_uniform(c["seed"], e, int(o.s["timestep"]), c["N"]))
        s = o.s
        assert np.array_equal(c["x"][t, k].astype(np.int64), s["x"].astype(np.int64)), what + ": ant cells x"
        assert np.array_equal(c["y"][t, k].astype(np.int64), s["y"].astype(np.int64)), what + ": ant cells y"
        assert_close(np.stack([c["x"][t, k], c["y"][t, k], c["theta"][t, k]]), np.stack([s["x"], s["y"], s["theta"]]), what + ": xyt")
        assert np.array_equal(c["holding"][t, k], s["holding"]), what + ": holding"
        assert np.array_equal(c["mandibles"][t, k], s["mandibles"]), what + ": mandibles"
        assert np.array_equal(c["reward_state"][t, k], s["reward_state"]), what + ": reward_state"
        assert c["anthill_food"][t, k] == float(s["anthill_food"]), what + ": anthill_food"
        if t % PLANES_EVERY == PLANES_EVERY - 1:
            assert c["food_sum"][t, k] == s["food"].sum(), what + ": food"
            assert c["explored_count"][t, k] == int(np.asarray(s["explored"]).sum()), what + ": explored"
            np.testing.assert_allclose(c["phero_sum"][t, k], s["phero"].sum(axis=(1, 2)), rtol=1e-5, err_msg=what + ": phero")
    compare_state(c["final"][k], [o], "env %d final" % e, c["cfg"])
    return float(o.s["anthill_food"])


def test_configs1_1024_envs_1000_steps():
    """BASELINE.json configs[1]: the reference's default generated map (200x200, 50 ants, walls + food
    + anthill from the generator, seeds 1000 + e) batched to 1024 envs on one GPU, random actions replayed from a
    recorded tape, main.py's loop for 1000 steps; 32 of the envs are compared with the oracle at EVERY step
    (observations, rewards, ant cells bit-exact, positions, carried food, mandibles, reward_state, delivered food), their
    planes every 20th step (remaining food, explored count, pheromone mass) and in full at the end."""
    import torch
    from antsrl_b200 import BatchedAnts
    from antsrl_b200.generator import BatchedEnvironmentGenerator, CirclesGenerator, stack_states
    E, N, T, seed = 1024, 50, 1000, 77
    picks = list(range(0, E, 33))[:32]
    gen = BatchedEnvironmentGenerator(200, 200, N, 2, 0, CirclesGenerator(20, 5, 10), CirclesGenerator(10, 5, 15),
                                      max_steps=T, seed_base=1000)
    states = gen.generate_states(E, 0)
    cfg = gen.cfg
    b = BatchedAnts(cfg, E, evap_mode="lazy", record="compact8", rng_seed=seed, env_id_base=0)
    b.import_state(stack_states(states, "all"))
    b.activate_all_pheromones(np.ones((E, N, 2)) * 10.0)          # agent.initialize, collect_agent.py:100-102
    rs = np.random.RandomState(12345)
    rot = (rs.randint(0, 3, (T, E, N)) - 1).astype(np.int8)       # the recorded action sequence
    ph = rs.randint(0, 3, (T, E, N)).astype(np.int8)
    d_rot, d_ph = torch.from_numpy(rot).cuda(), torch.from_numpy(ph).cuda()
    K = len(picks)
    idx = torch.tensor(picks, device="cuda")
    rec = dict(obs=np.zeros((T, K, N, 7, 7, 6), np.float32), ast=np.zeros((T, K, N, 2), np.float32),
               rew=np.zeros((T, K, N)), x=np.zeros((T, K, N)), y=np.zeros((T, K, N)), theta=np.zeros((T, K, N)),
               holding=np.zeros((T, K, N)), mandibles=np.zeros((T, K, N), np.uint8), reward_state=np.zeros((T, K, N), np.uint8),
               anthill_food=np.zeros((T, K)), food_sum=np.zeros((T, K)), explored_count=np.zeros((T, K), np.int64),
               phero_sum=np.zeros((T, K, 2)))
    b.observe()                                                   # main.py:88
    done = False
    for t in range(T):
        obs, ast, rew, done = b.step(d_rot[t], d_ph[t])
        rec["obs"][t] = obs[idx].cpu().numpy()
        rec["ast"][t] = ast[idx].cpu().numpy()
        rec["rew"][t] = rew[idx].cpu().numpy()
        assert done == (t == T - 1)                               # Q15
        b.update(None)
        st = b.export_state(keys=("x", "y", "theta", "holding", "mandibles", "reward_state", "anthill_food"))
        for k in ("x", "y", "theta", "holding", "mandibles", "reward_state", "anthill_food"):
            rec[k][t] = st[k][picks]
        if t % PLANES_EVERY == PLANES_EVERY - 1:                  # the planes of the picked envs (one export per env)
            for j, e in enumerate(picks):
                pl = b.export_state(keys=("food", "explored", "phero"), envs=(e, 1))
                rec["food_sum"][t, j] = pl["food"].sum()
                rec["explored_count"][t, j] = int(pl["explored"].sum())
                rec["phero_sum"][t, j] = pl["phero"][0].sum(axis=(1, 2))
    final = [b.export_state(envs=(e, 1)) for e in picks]
    whole = b.export_state(keys=("anthill_food", "holding"))
    b.close()
    _CFG2.update(cfg=cfg, states=[states[e] for e in picks], picks=picks, N=N, T=T, seed=seed, rot=rot, ph=ph,
                 final=final, **rec)
    procs = max(1, min(os.cpu_count() or 1, K))
    with mp.get_context("fork").Pool(procs) as pool:
        delivered = pool.map(_cfg2_oracle_worker, range(K))
    _CFG2.clear()
    assert delivered == [float(f["anthill_food"][0]) for f in final]
    assert whole["anthill_food"].sum() > 0 and whole["holding"].sum() > 0      # food is found and delivered at all


# ------------------------------------------------------------------------------------------------ decay at read time
@pytest.mark.parametrize("mode", ["compact8", "compact", "lazy", "tiles"])
def test_observation_decay_out_to_table_end(mode):
    """What an OBSERVATION shows of a saturated deposit as it ages (the row kernel evaluates 2^(age log2 keep) in f32;
    the exported field goes through the exact table): standing ants (max_speed 0) deposit once, then the field only
    evaporates; observations at ages 1, 2, 300, 5000 and around the end of the decay table (the update at which the
    reference's < 0.01 cut zeroes the value) are compared with the oracle's to 1e-5 relative (absolute floor 1e-10:
    the last visible values are ~4e-5)."""
    import torch
    from antsrl_b200 import BatchedAnts
    from oracle.antsrl_oracle import OracleEnv
    scen = [make_scenario(seed=600 + e, w=40, h=36, n_ants=60, steps=2, n_walls=3, n_food=4, max_speed=0.0) for e in range(2)]
    cfg = scen[0][0]
    oracles = [OracleEnv(c, i) for c, i, _ in scen]
    record = mode if mode.startswith("compact") else "f64"
    b = BatchedAnts(cfg, 2, evap_mode="lazy" if mode.startswith("compact") else mode, record=record)
    b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
    # the table the library builds: max_val decayed with the reference's per-step rounding and cut (pheromone.py:44-45)
    v, tab_len = 255.0, 1
    while v != 0.0:
        v = v * (1 - cfg["evap_factor"])
        v = 0.0 if v < 0.01 else v
        tab_len += 1
    assert 10000 < tab_len < 10300

    def check(what):
        obs, ast, st, rew = b.observe()
        obs = obs.cpu().numpy()
        for e, o in enumerate(oracles):
            ref = o.observation()[0]
            assert_close(obs[e], ref, "%s env %d" % (what, e), rtol=1e-5, atol=1e-10)
        return obs

    check("before")
    N = cfg["n_ants"]
    ph = np.stack([s[2]["ph"][0] for s in scen]).astype(np.int8)
    rot = np.stack([s[2]["rot"][0] for s in scen]).astype(np.int8)
    for e, o in enumerate(oracles):
        o.step(rot[e].astype(np.int64), ph[e].astype(np.int64))
        o.update(scen[e][2]["noise"][0])
    b.step(torch.from_numpy(rot).cuda(), torch.from_numpy(ph).cuda())
    b.update(torch.from_numpy(np.stack([s[2]["noise"][0] for s in scen])).cuda())        # the deposit (age 0)
    zero = np.zeros((2, N, 2))
    b.activate_all_pheromones(zero)                                                     # nothing is deposited afterwards
    for o in oracles:
        o.activate_all_pheromones(zero[0])
    seen = check("age 0")
    assert (seen[..., 1:3] == 1.0).any(), "a saturated deposit must be visible"
    ages = [1, 2, 300, 5000, tab_len - 3, tab_len - 2, tab_len - 1, tab_len, tab_len + 1]
    age = 0
    noise = torch.zeros((2, N), dtype=torch.float64, device="cuda")
    for target in ages:
        while age < target:
            b.update(noise)
            for o in oracles:
                o.update(np.zeros(N))
            age += 1
        seen = check("age %d" % age)
        visible = seen[..., 1:3][seen[..., 1:3] > 0]
        if target <= tab_len - 2:
            assert visible.size > 0, "age %d: the deposit must still be visible" % age
        if target >= tab_len - 1:
            assert visible.size == 0, "age %d: the deposit must have been cut to zero" % age
    compare_state(b.export_state(), oracles, "final", cfg)
    b.close()

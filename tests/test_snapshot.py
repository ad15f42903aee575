"""Saved episodes for the viewer: `Environment.save_state()` (environment.py:36-40), the
pickled list of states main.py:136-147 writes, and the same for one environment of a device-resident batch
(antsrl_b200/snapshot.py over ants_export_env_state).  The fixture tests/golden/episode_s7.arl was written by the
UNMODIFIED reference (tests/golden/make_golden.py arl)."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import arl_util
from test_dropin import PRELUDE, ROOT, GOLDEN

ARL = os.path.join(GOLDEN, "episode_s7.arl")
CASE = json.load(open(os.path.join(GOLDEN, "episode_s7.json")))

# the fixture's episode through the drop-in classes (same generator arguments, same RNG seeding, same action draws)
EPISODE = """
import pickle
c = %r
api = RLApi(All_Rewards(1, 2, 10, 1, 3), 1, 1, 40 / 180 * np.pi, 0.05, 0.5)
g = EnvironmentGenerator(c['w'], c['h'], c['n_ants'], 2, c['n_rocks'], CirclesGenerator(*c['food']),
                         CirclesGenerator(*c['walls']), c['steps'], seed=c['seed'])
env = g.generate(api)
api.ants.activate_all_pheromones(np.ones((c['n_ants'], 2)) * 10)
np.random.seed(c['np_seed'])
act = np.random.RandomState(c['action_seed'])
""" % (CASE,)


def run_snippet(body, prelude=PRELUDE):
    code = prelude + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def live_env(state):
    walls = [o for o in state.objects if o.cls.endswith(".Walls")][0]
    return walls.environment


def test_fixture_has_the_reference_layout():
    """What the reference writes: 10 states; every visualisation copy is listed twice (its constructor registers it,
    environment.py:8-9, and save_state adds it again, :39), the live Walls object once (walls.py:16-17), and through
    it the live environment is part of the file."""
    states = arl_util.load_neutral(ARL)
    assert len(states) == CASE["steps"]
    names = [o.cls.rsplit(".", 1)[1] for o in states[0].objects]
    assert names == ["AnthillVisualization"] * 2 + ["Walls"] + ["FoodVisualization"] * 2 + \
        ["CircleObstaclesVisualization"] * 2 + ["AntsVisualization"] * 2 + ["PheromoneVisualization"] * 4 + \
        ["RLVisualization"] * 2
    assert states[0].objects[0] is states[0].objects[1]
    walls = [[o for o in s.objects if o.cls.endswith(".Walls")][0] for s in states]
    assert all(w is walls[0] for w in walls)                 # the viewer keys its background on this identity
    assert [s.timestep for s in states] == [1] * CASE["steps"]       # save_state builds a fresh Environment
    assert live_env(states[0]).timestep == CASE["steps"] + 1


def test_reference_file_loads_with_the_dropin_classes():
    out = run_snippet("""
        import pickle
        from environment.ants import AntsVisualization
        from environment.pheromone import PheromoneVisualization
        from environment.anthill import AnthillVisualization
        states = pickle.load(open(%r, 'rb'))
        assert len(states) == 10 and all(isinstance(s, Environment) for s in states)
        e = states[3]
        assert (e.w, e.h, e.max_time, e.timestep) == (48, 40, 10, 1)
        a = [o for o in e.objects if isinstance(o, AntsVisualization)][0]
        assert a.ants.shape == (12, 3) and a.reward_state.dtype == np.uint8
        p = [o for o in e.objects if isinstance(o, PheromoneVisualization)]
        assert len(p) == 4 and p[0].phero.dtype == np.uint8 and p[0].max_val == 255 and p[0].color == (255, 64, 0)
        w = [o for o in e.objects if isinstance(o, Walls)][0]
        assert w.map.dtype == bool and w.map.shape == (48, 40)
        # the live environment behind Walls arrives as drop-in objects with their host mirrors filled
        live = w.environment
        assert live.timestep == 11 and live._bridge is None
        ants = [o for o in live.objects if isinstance(o, Ants)][0]
        assert ants.ants.shape == (12, 3) and ants.mandibles.dtype == np.int64 and ants.phero_activation.shape == (12, 2)
        api = [o for o in live.objects if isinstance(o, RLApi)][0]
        assert api.reward.explored_map.dtype == bool and api.reward.previous_dist.shape == (12,)
        assert [o for o in live.objects if isinstance(o, Anthill)][0].food >= 0
        print('loaded ok')
    """ % ARL)
    assert "loaded ok" in out


def _dropin_blob(tmp_path, body):
    path = str(tmp_path / "dropin.arl")
    run_snippet(EPISODE + textwrap.dedent(body) + "\npickle.dump(states, open(%r, 'wb'))\n" % path)
    return path


def test_dropin_file_has_the_reference_structure(tmp_path):
    """Without a GPU: the state saved right after generate() has the reference's object list, attribute names and
    dtypes, for the visualisation copies and for the live objects pickled behind Walls."""
    path = _dropin_blob(tmp_path, "states = [env.save_state()]")
    got, ref = arl_util.load_neutral(path)[0], arl_util.load_neutral(ARL)[0]
    assert [o.cls for o in got.objects] == [o.cls for o in ref.objects]
    assert set(vars(got)) == set(vars(ref)) == {"w", "h", "objects", "max_time", "timestep"}
    for a, b in zip(got.objects, ref.objects):
        da, db = arl_util.describe(a), arl_util.describe(b)
        if a.cls.endswith("AntsVisualization"):
            da["mandibles"] = db["mandibles"]          # bool before the first step, int64 after (ants.py:36,112)
        if a.cls.endswith("AnthillVisualization"):
            da["food"] = db["food"]                    # int 0 before the first update, np.float64 after (anthill.py:27,46)
        assert da == db, (a.cls, da, db)
    lg, lr = live_env(got), live_env(ref)
    assert [o.cls for o in lg.objects] == [o.cls for o in lr.objects]
    for a, b in zip(lg.objects, lr.objects):
        assert set(vars(a)) == set(vars(b)), (a.cls, set(vars(a)) ^ set(vars(b)))
    rw_g = [o for o in lg.objects if o.cls.endswith("RLApi")][0].reward
    rw_r = [o for o in lr.objects if o.cls.endswith("RLApi")][0].reward
    # rewards_anthillheading is a temporary of All_Rewards.observation (reward_custom.py:101); _aliased is Q18's flag
    assert set(vars(rw_r)) - set(vars(rw_g)) <= {"rewards_anthillheading"}
    assert set(vars(rw_g)) - set(vars(rw_r)) <= {"_aliased"}


def _reference_root():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import ref_harness
    return ref_harness.REFERENCE_ROOT if ref_harness.reference_available() else None


@pytest.mark.skipif(_reference_root() is None, reason="needs the reference (checkout or oracle/_ref built by oracle/build_ref.py)")
def test_dropin_file_loads_under_the_reference_classes(tmp_path):
    """The direction the viewer needs: a file written by the drop-in layer, opened by the reference's own classes."""
    path = _dropin_blob(tmp_path, "states = [env.save_state()]")
    prelude = """
import sys, types, pickle
for name in ("noise", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, %r)
import numpy as np
""" % _reference_root()
    out = run_snippet("""
        import environment.ants as m
        assert m.__file__.startswith(%r)
        from environment.ants import Ants, AntsVisualization
        from environment.walls import Walls
        from environment.pheromone import Pheromone, PheromoneVisualization
        from environment.anthill import AnthillVisualization
        from environment.RL_api import RLVisualization, RLApi
        states = pickle.load(open(%r, "rb"))
        e = states[0]
        kinds = [type(o).__name__ for o in e.objects]
        assert kinds.count("PheromoneVisualization") == 4 and kinds.count("Walls") == 1, kinds
        w = [o for o in e.objects if isinstance(o, Walls)][0]
        assert w.map.shape == (e.w, e.h)
        a = [o for o in e.objects if isinstance(o, AntsVisualization)][0]
        assert a.ants.shape == (12, 3)
        live = w.environment
        ants = [o for o in live.objects if isinstance(o, Ants)][0]
        assert ants.ants.shape == (12, 3) and ants.x.shape == (12,) and ants.seed.shape == (12,)
        api = [o for o in live.objects if isinstance(o, RLApi)][0]
        assert api.ants is ants and api.reward.explored_map.shape == (48, 40)
        # and the reference can carry on from it: one step of its own loop on the unpickled live environment
        obs, state, rew, done = api.step(np.zeros(12, dtype=int), np.zeros(12, dtype=int))
        live.update()
        assert obs.shape == (12, 7, 7, 7) and live.timestep == 2
        print("reference loaded ok")
    """ % (_reference_root(), path), prelude=prelude)
    assert "reference loaded ok" in out


@pytest.mark.gpu
def test_dropin_episode_file_matches_the_reference(tmp_path):
    """The fixture's episode through the drop-in classes on the GPU, saved like main.py saves it: every state equal
    to the reference's (object list, names, dtypes; integer arrays and the uint8 planes identical, floats to 1e-5),
    the live environment behind Walls included."""
    path = _dropin_blob(tmp_path, """
        api.observation()
        states = []
        for t in range(c['steps']):
            rot = act.randint(0, 3, c['n_ants']) - 1
            ph = act.randint(0, 3, c['n_ants'])
            api.step(rot, ph)
            env.update()
            states.append(env.save_state())
        blob = pickle.dumps(states)           # with the live bridge attached: must not try to pickle the handle
        again = pickle.loads(blob)
        assert again[0].objects[2].environment._bridge is None and env._bridge is not None
    """)
    got, ref = arl_util.load_neutral(path), arl_util.load_neutral(ARL)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        arl_util.compare_states(g, r)
    walls = [[o for o in s.objects if o.cls.endswith(".Walls")][0] for s in got]
    assert all(w is walls[0] for w in walls)
    lg, lr = live_env(got[0]), live_env(ref[0])
    assert lg.timestep == lr.timestep
    for a, b in zip(lg.objects, lr.objects):
        assert a.cls == b.cls
        for k, v in vars(b).items():
            # original_ants_position (RL_api.py:64) is written once and never read; in the reference it is a view
            # that follows the ants, here it keeps the initial positions
            if isinstance(v, np.ndarray) and v.dtype.kind in "fbiu" and k not in ("perceptive_field", "original_ants_position"):
                w = getattr(a, k)
                assert w.shape == v.shape, (a.cls, k)
                np.testing.assert_allclose(np.asarray(w, float), np.asarray(v, float), rtol=1e-5, atol=1e-7,
                                           err_msg="%s.%s" % (a.cls, k))


@pytest.mark.gpu
@pytest.mark.parametrize("record,evap_mode", [("f64", "dense"), ("f64", "lazy"), ("compact", "lazy"), ("compact8", "lazy")])
def test_export_env_window_equals_whole_batch(record, evap_mode):
    """ants_export_env_state: the window [env0, env0 + n) of a batch is the same slice of the whole export."""
    import torch
    from antsrl_b200 import BatchedAnts
    from parity_util import import_scenarios
    from scenarios import make_scenario
    scen = [make_scenario(seed=40 + e, w=40, h=56, n_ants=17, n_rocks=2, steps=6) for e in range(5)]
    cfg = scen[0][0]
    b = BatchedAnts(cfg, len(scen), evap_mode=evap_mode, record=record)
    import_scenarios(b, scen)
    b.observe()
    for t in range(6):
        rot = torch.from_numpy(np.stack([s[2]["rot"][t] for s in scen]).astype(np.int8)).cuda()
        ph = torch.from_numpy(np.stack([s[2]["ph"][t] for s in scen]).astype(np.int8)).cuda()
        b.step(rot, ph)
        b.update(None)
    whole = b.export_state()
    for env0, n in [(0, 5), (0, 1), (2, 2), (4, 1)]:
        part = b.export_state(envs=(env0, n))
        assert set(part) == set(whole)
        for k, v in whole.items():
            if isinstance(v, np.ndarray):
                assert part[k].shape == (n,) + v.shape[1:], k
                assert np.array_equal(part[k], v[env0:env0 + n], equal_nan=True), (k, env0, n)
            else:
                assert part[k] == v, k
    part = b.export_state(keys=("x", "food"), envs=(3, 1))
    assert set(part) == {"x", "food", "timestep", "rw_alias", "act_bool"}
    for bad in [(-1, 1), (5, 1), (3, 3), (0, 0)]:
        with pytest.raises(ValueError):
            b.export_state(envs=bad)
    b.close()


@pytest.mark.gpu
def test_batch_snapshot_and_recorder(tmp_path):
    """snapshot_env / EpisodeRecorder: environment 1 of a 3-env batch recorded after every update, written as an
    .arl file, read back without the classes and compared with the state of the batch."""
    out = run_snippet("""
        import pickle, torch
        sys.path.insert(0, %r)
        from antsrl_b200 import BatchedAnts
        from antsrl_b200.snapshot import EpisodeRecorder, load_episode
        from parity_util import import_scenarios
        from scenarios import make_scenario
        scen = [make_scenario(seed=70 + e, w=48, h=40, n_ants=12, n_rocks=2, steps=5) for e in range(3)]
        b = BatchedAnts(scen[0][0], 3, evap_mode='lazy', record='compact8')
        import_scenarios(b, scen)
        rec = EpisodeRecorder(b, env_index=1)
        b.observe()
        want = []
        for t in range(5):
            rot = torch.from_numpy(np.stack([s[2]['rot'][t] for s in scen]).astype(np.int8)).cuda()
            ph = torch.from_numpy(np.stack([s[2]['ph'][t] for s in scen]).astype(np.int8)).cuda()
            b.step(rot, ph)
            b.update(None)
            rec.record()
            want.append(b.export_state())
        path = %r
        rec.save(path)
        np.savez(path + '.npz', **{'%%d_%%s' %% (t, k): np.asarray(v) for t, w in enumerate(want) for k, v in w.items()})
        states = load_episode(path)
        assert len(states) == 5 and all(isinstance(s, Environment) for s in states)
        print('recorded ok')
    """ % (os.path.join(ROOT, "tests"), str(tmp_path / "batch.arl")))
    assert "recorded ok" in out
    states = arl_util.load_neutral(str(tmp_path / "batch.arl"))
    z = np.load(str(tmp_path / "batch.arl.npz"))
    ref_names = [o.cls for o in arl_util.load_neutral(ARL)[0].objects]
    walls = [[o for o in s.objects if o.cls.endswith(".Walls")][0] for s in states]
    assert all(w is walls[0] for w in walls)
    for t, s in enumerate(states):
        assert [o.cls for o in s.objects] == ref_names
        get = lambda k: z["%d_%s" % (t, k)][1]
        assert s.timestep == 1 and int(z["%d_timestep" % t]) == t + 2
        by = {}
        for o in s.objects:
            by.setdefault(o.cls.rsplit(".", 1)[1], []).append(o)
        for name, attrs in arl_util.VIEWER_READS.items():
            for a in attrs:
                assert hasattr(by[name][0], a), (name, a)
        a = by["AntsVisualization"][0]
        assert np.array_equal(a.ants, np.stack([get("x"), get("y"), get("theta")], axis=1))
        assert np.array_equal(a.holding, get("holding")) and np.array_equal(a.reward_state, get("reward_state"))
        assert a.mandibles.dtype == np.int64 and np.array_equal(a.mandibles, get("mandibles"))
        assert np.array_equal(by["FoodVisualization"][0].qte, get("food").astype(np.uint8))
        ph = [by["PheromoneVisualization"][0], by["PheromoneVisualization"][2]]
        for k in range(2):
            assert ph[k].phero.dtype == np.uint8 and np.array_equal(ph[k].phero, get("phero")[k].astype(np.uint8))
        assert ph[0].color == (255, 64, 0) and ph[1].color == (64, 64, 255)
        assert np.array_equal(by["RLVisualization"][0].heatmap, get("explored").astype(bool))
        assert np.array_equal(by["Walls"][0].map, get("walls").astype(bool))
        assert np.array_equal(by["CircleObstaclesVisualization"][0].centers, get("rock_centers"))
        hill = by["AnthillVisualization"][0]
        assert [hill.x, hill.y, hill.radius] == list(get("anthill_xyr")) and hill.food == float(get("anthill_food"))

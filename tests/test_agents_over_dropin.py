"""The reference's own agent code over the drop-in packages (north_star: "drops in under the existing agents").

`agents/collect_agent.py` of the UNMODIFIED reference (oracle/_ref or the checkout) is imported with
`antsrl_b200.dropin_path()` FIRST on sys.path, so its `from environment.RL_api import RLApi`,
`from environment.pheromone import Pheromone` resolve to the CUDA-backed classes, and main.py's loop
(main.py:82-131: setup, initialize, observation, then get_action -> api.step -> update_replay_memory -> train ->
env.update) runs for 60 steps.  The same loop runs over the reference's own environment package on the CPU with the
same seeds; with epsilon = 1 the actions are the agent's exploration branch (collect_agent.py:172-177, drawn from the
global numpy RNG that the wall collisions also consume), so both runs must agree step by step: actions identical,
rewards and observations to 1e-5, the replay memory the agent filled, the ants' cells bit-exact."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness       # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(),
                                reason="needs the reference (checkout or oracle/_ref built by oracle/build_ref.py)")

LOOP = """
import sys, types, json, random, time
for name in ("noise", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
REF, ROOT, USE_DROPIN, OUT, STEPS = %(ref)r, %(root)r, %(dropin)r, %(out)r, %(steps)d
sys.path.insert(0, REF)
if USE_DROPIN:
    sys.path.insert(0, ROOT)
    import antsrl_b200
    sys.path.insert(0, antsrl_b200.dropin_path())
import numpy as np, torch
import environment.RL_api as m_api
assert ("dropin" in m_api.__file__) == bool(USE_DROPIN), m_api.__file__
from environment.RL_api import RLApi
from environment.rewards.reward_custom import All_Rewards
from generator.environment_generator import EnvironmentGenerator
from generator.map_generators import CirclesGenerator
# the reference's `agents` directory has no __init__.py (a namespace package): an installed distribution that happens to
# be called `agents` would shadow it, so the package is bound to the reference's directory explicitly
pkg = types.ModuleType("agents"); pkg.__path__ = [REF + "/agents"]; sys.modules["agents"] = pkg
import agents.collect_agent as m_agent
assert m_agent.__file__.startswith(REF), m_agent.__file__          # the reference's own agent code
from agents.collect_agent import CollectAgent

random.seed(11); np.random.seed(12); torch.manual_seed(13)
api = RLApi(All_Rewards(1, 2, 10, 1, 3), 1, 1, 40 / 180 * np.pi, 0.05, 0.5)       # main.py:42-50
agent = CollectAgent(epsilon=1.0, discount=0.99, rotations=3, pheromones=3)
gen = EnvironmentGenerator(200, 200, 50, 2, 0, CirclesGenerator(20, 5, 10), CirclesGenerator(10, 5, 15), STEPS, seed=1000)
env = gen.generate(api)                                                          # main.py:79
agent.setup(api, None)                                                           # main.py:83
agent.initialize(api)                                                            # main.py:86
random.seed(21); np.random.seed(22)
obs, agent_state, state = api.observation()                                      # main.py:88
rec = dict(rot=[], ph=[], reward=[], obs_sum=[], cells=[], loss=[], done=[])
t0 = time.perf_counter()
for s in range(STEPS):
    action = agent.get_action(obs, agent_state, True)                            # main.py:95
    new_state, new_agent_state, reward, done = api.step(*action[:2])             # main.py:98
    agent.update_replay_memory(obs, agent_state, action, reward, new_state, new_agent_state, done)   # main.py:102
    loss = agent.train(done, s)                                                  # main.py:105
    obs, agent_state = new_state, new_agent_state
    env.update()                                                                 # main.py:131
    rec["rot"].append(np.asarray(action[0]).tolist()); rec["ph"].append(np.asarray(action[1]).tolist())
    rec["reward"].append(np.asarray(reward, dtype=float).tolist()); rec["obs_sum"].append(float(np.asarray(obs, dtype=float).sum()))
    xy = api.ants.ants[:, :2].astype(int)
    rec["cells"].append((xy[:, 0] * 200 + xy[:, 1]).tolist()); rec["loss"].append(float(loss)); rec["done"].append(bool(done))
rec["ms_per_step"] = (time.perf_counter() - t0) * 1000 / STEPS
mem = agent.replay_memory
rec["mem_fill"] = len(mem); rec["mem_head"] = mem.head
rec["mem_states_sum"] = float(mem.states[:len(mem)].double().sum()); rec["mem_rewards_sum"] = float(mem.rewards[:len(mem)].double().sum())
rec["mem_actions_sum"] = int(mem.actions[:len(mem)].sum()); rec["obs_shape"] = list(np.asarray(obs).shape)
rec["obs_dtype"] = str(np.asarray(obs).dtype)
json.dump(rec, open(OUT, "w"))
print("loop ok")
"""


def _run_loop(tmp_path, use_dropin, steps=60):
    out = str(tmp_path / ("dropin.json" if use_dropin else "reference.json"))
    code = LOOP % dict(ref=ref_harness.REFERENCE_ROOT, root=ROOT, dropin=use_dropin, out=out, steps=steps)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0 and "loop ok" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
    return json.load(open(out))


def test_reference_agent_runs_over_the_reference_environment(tmp_path):
    """The CPU half on its own (also the proof that oracle/_ref carries the agents): the loop runs, trains and fills
    the replay memory."""
    ref = _run_loop(tmp_path, False, steps=25)
    assert ref["mem_fill"] == 25 * 50 and ref["obs_shape"] == [50, 7, 7, 6] and any(l > 0 for l in ref["loss"])


@pytest.mark.gpu
def test_reference_agent_over_the_dropin_matches_the_reference_run(tmp_path):
    ref = _run_loop(tmp_path, False)
    got = _run_loop(tmp_path, True)
    assert got["obs_shape"] == ref["obs_shape"] == [50, 7, 7, 6] and got["obs_dtype"] == ref["obs_dtype"] == "float64"
    for t in range(60):
        assert got["rot"][t] == ref["rot"][t] and got["ph"][t] == ref["ph"][t], "actions differ at step %d" % t
        assert got["cells"][t] == ref["cells"][t], "ant cells differ at step %d" % t
        np.testing.assert_allclose(got["reward"][t], ref["reward"][t], rtol=1e-5, atol=1e-7, err_msg="reward %d" % t)
        np.testing.assert_allclose(got["obs_sum"][t], ref["obs_sum"][t], rtol=1e-5, err_msg="obs %d" % t)
        assert got["done"][t] == ref["done"][t]
    assert got["done"][-1] is True
    assert got["mem_fill"] == ref["mem_fill"] == 3000 and got["mem_head"] == ref["mem_head"]
    assert got["mem_actions_sum"] == ref["mem_actions_sum"]
    np.testing.assert_allclose(got["mem_states_sum"], ref["mem_states_sum"], rtol=1e-5)
    np.testing.assert_allclose(got["mem_rewards_sum"], ref["mem_rewards_sum"], rtol=1e-5)
    # the agent trained on both (same minibatch indices: `random` is seeded; inputs equal to 1e-5)
    assert any(l > 0 for l in got["loss"]) and any(l > 0 for l in ref["loss"])
    np.testing.assert_allclose(got["loss"], ref["loss"], rtol=5e-3, atol=1e-5)
    print("drop-in %.2f ms/step, reference %.2f ms/step (agent included)" % (got["ms_per_step"], ref["ms_per_step"]))

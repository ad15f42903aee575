"""DeviceActionSelector (antsrl_b200/device_loop.py): the agents' get_action (agents/collect_agent.py:161-177) on the
device, and main.py's loop around it -- observation -> two-head policy -> epsilon-greedy selection -> step -> replay
ingest -> update -- with nothing but the per-env exploration draw crossing PCIe; the environment it drives stays equal
to the oracle driven by the same actions."""
import numpy as np
import pytest

from parity_util import compare_state, stack_init
from scenarios import make_scenario

pytestmark = pytest.mark.gpu


def test_selector_branches_match_the_reference_rule():
    import torch
    from antsrl_b200 import BatchedAnts
    from antsrl_b200.device_loop import DeviceActionSelector
    from oracle.antsrl_oracle import philox_actions
    scen = [make_scenario(seed=1500 + e, w=32, h=32, n_ants=20, steps=3) for e in range(4)]
    cfg = scen[0][0]
    b = BatchedAnts(cfg, 4, evap_mode="lazy", record="compact8", env_id_base=3)
    b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
    b.observe()
    g = torch.Generator(device="cuda").manual_seed(1)
    q_rot = torch.rand((4, 20, 3), device="cuda", generator=g)
    q_ph = torch.rand((4, 20, 3), device="cuda", generator=g)
    greedy_rot = (q_rot.argmax(2) - 1).cpu().numpy()
    greedy_ph = q_ph.argmax(2).cpu().numpy()
    sel = DeviceActionSelector(b, epsilon=0.0, seed=9)
    rot, ph, ex = sel.select(q_rot, q_ph)                      # epsilon 0: random() > 0 always follows the network
    assert not ex.any() and np.array_equal(rot.cpu().numpy(), greedy_rot) and np.array_equal(ph.cpu().numpy(), greedy_ph)
    assert rot.dtype == torch.int8 and rot.shape == (4, 20)
    sel = DeviceActionSelector(b, epsilon=1.0, seed=9)
    rot, ph, ex = sel.select(q_rot, q_ph)                      # epsilon 1: every env explores (collect_agent.py:172-177)
    assert ex.all()
    for e in range(4):
        r_ref, p_ref = philox_actions(9, 3 + e, 1, 20)
        assert np.array_equal(rot[e].cpu().numpy(), r_ref) and np.array_equal(ph[e].cpu().numpy(), p_ref)
    rot, ph, ex = sel.select(q_rot, q_ph, training=False)      # not training: the network, whatever epsilon is
    assert not ex.any() and np.array_equal(rot.cpu().numpy(), greedy_rot)
    rot, ph, ex = sel.select(q_rot, q_ph, explore=[True, False, False, True])
    assert np.array_equal(rot[1].cpu().numpy(), greedy_rot[1]) and np.array_equal(ph[2].cpu().numpy(), greedy_ph[2])
    r_ref, p_ref = philox_actions(9, 3, 1, 20)
    assert np.array_equal(rot[0].cpu().numpy(), r_ref) and np.array_equal(ph[0].cpu().numpy(), p_ref)
    rot, ph, ex = sel.select(q_rot, None, training=False)      # a policy without a pheromone head (explore agents)
    assert ph is None and np.array_equal(rot.cpu().numpy(), greedy_rot)
    b.close()


def test_on_device_loop_keeps_parity_with_the_oracle():
    """main.py:88-131 on the device for 20 steps: a random two-head MLP picks actions from the f32 observations, the
    selector mixes in exploring environments, the replay memory ingests device tensors; the oracle is driven by the
    actions the device chose (read back for the check only) and must agree with the environment state throughout."""
    import torch
    from antsrl_b200 import BatchedAnts
    from antsrl_b200.device_loop import DeviceActionSelector
    from antsrl_b200.replay import DeviceReplayMemory
    from oracle.antsrl_oracle import OracleEnv, philox_uniform
    E, N, T = 3, 40, 20
    scen = [make_scenario(seed=1600 + e, w=56, h=48, n_ants=N, n_rocks=2, steps=T, n_walls=5, n_food=10) for e in range(E)]
    cfg = scen[0][0]
    C = len(cfg["channels"])
    oracles = [OracleEnv(c, i) for c, i, _ in scen]
    b = BatchedAnts(cfg, E, evap_mode="lazy", record="compact8", rng_seed=4)
    b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
    torch.manual_seed(0)
    body = torch.nn.Sequential(torch.nn.Linear(49 * C + 2, 32), torch.nn.ReLU()).cuda()
    head_rot, head_ph = torch.nn.Linear(32, 3).cuda(), torch.nn.Linear(32, 3).cuda()
    sel = DeviceActionSelector(b, epsilon=0.4, seed=77)
    mem = DeviceReplayMemory(5000, (7, 7, C), (2,), 2)
    obs, ast, _, _ = b.observe()
    for o in oracles:
        o.observation()
    obs, ast = obs.clone(), ast.clone()
    explored_envs = 0
    for t in range(T):
        with torch.no_grad():
            hdn = body(torch.cat([obs.reshape(E * N, -1), ast.reshape(E * N, 2)], dim=1))
            rot, ph, ex = sel.select(head_rot(hdn), head_ph(hdn))
        explored_envs += int(ex.sum())
        new_obs, new_ast, rew, done = b.step(rot, ph)
        mem.extend(obs, ast, (rot, ph), rew, new_obs, new_ast, done)
        rot_h, ph_h = rot.cpu().numpy(), ph.cpu().numpy()
        assert rot_h.min() >= -1 and rot_h.max() <= 1 and ph_h.min() >= 0 and ph_h.max() <= 2
        for e, o in enumerate(oracles):
            r = o.step(rot_h[e].astype(np.int64), ph_h[e].astype(np.int64))
            np.testing.assert_allclose(new_obs[e].cpu().numpy(), r[0], rtol=1e-5, atol=1e-7, err_msg="obs t=%d" % t)
            np.testing.assert_allclose(rew[e].cpu().numpy(), r[2], rtol=1e-5, atol=1e-7)
            o.update(philox_uniform(4, e, int(o.s["timestep"]), N))
        b.update(None)
        obs, ast = new_obs.clone(), new_ast.clone()
    compare_state(b.export_state(), oracles, "after the on-device loop", cfg)
    assert len(mem) == E * N * T and mem.actions.dtype == torch.int64
    assert 0 < explored_envs < E * T
    s = mem.random_access(64)
    assert s[0].is_cuda and s[0].shape == (64, 7, 7, C) and s[6].dtype == torch.bool
    b.close()

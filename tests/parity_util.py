"""Drive the CUDA path (through the C ABI) and the oracle side by side on the same seeded scenarios."""
import numpy as np

from oracle.antsrl_oracle import OracleEnv

RTOL = 1e-5          # north_star: float state within 1e-5 relative
ATOL = 1e-7

ANT_F64 = ("x", "y", "theta", "prev_x", "prev_y", "prev_theta", "holding", "seed", "rw_holding_prev",
           "rw_prev_dist", "rewards")
INT_KEYS = ("mandibles", "reward_state", "walls", "explored")


def stack_init(cfg, inits):
    """list of per-env init dicts -> batched state dict for BatchedAnts.import_state."""
    keys = set(inits[0].keys())
    out = {}
    for k in keys:
        if k in ("act_bool", "rw_alias", "timestep"):
            out[k] = inits[0][k]
        else:
            out[k] = np.stack([np.asarray(i[k]) for i in inits])
    if "rw_prev_dist" not in out and cfg["reward_kind"] == "all":   # All_Rewards.setup, reward_custom.py:77
        ax = out["anthill_xyr"][:, 0:1].astype(float)
        ay = out["anthill_xyr"][:, 1:2].astype(float)
        out["rw_prev_dist"] = ((out["x"] - ax) ** 2 + (out["y"] - ay) ** 2) ** 0.5
    return out


def import_scenarios(batch, scenarios):
    """Upload the initial states of a list of (cfg, init, tape) scenarios into a BatchedAnts batch."""
    batch.import_state(stack_init(scenarios[0][0], [i for _, i, _ in scenarios]))


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %r vs %r" % (what, a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def compare_state(gpu_state, oracles, what, cfg):
    """integer / index state bit-exact, float state within RTOL."""
    for k in INT_KEYS:
        ref = np.stack([np.asarray(o.s[k]).astype(np.uint8) for o in oracles])
        assert np.array_equal(gpu_state[k], ref), "%s: %s differs (%d entries)" % (
            what, k, int((gpu_state[k] != ref).sum()))
    # cells of ants (truncated positions) and carried food are index / integer state
    for k in ("x", "y"):
        ref = np.stack([o.s[k] for o in oracles])
        assert np.array_equal(gpu_state[k].astype(np.int64), ref.astype(np.int64)), "%s: ant cells (%s) differ" % (what, k)
    for k in ANT_F64:
        ref = np.stack([np.asarray(o.s[k], dtype=float) for o in oracles])
        assert_close(gpu_state[k], ref, "%s: %s" % (what, k))
    assert_close(gpu_state["activation"], np.stack([o.s["activation"] for o in oracles]), what + ": activation")
    assert_close(gpu_state["phero"], np.stack([o.s["phero"] for o in oracles]), what + ": phero")
    assert_close(gpu_state["food"], np.stack([o.s["food"] for o in oracles]), what + ": food")
    assert_close(gpu_state["anthill_food"], np.array([float(o.s["anthill_food"]) for o in oracles]),
                 what + ": anthill_food")
    if cfg["n_rocks"]:
        assert_close(gpu_state["rock_centers"], np.stack([o.s["rock_centers"] for o in oracles]),
                     what + ": rock_centers")
    assert gpu_state["timestep"] == int(oracles[0].s["timestep"]), what + ": timestep"


def _np(a):
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def run_parity(scenarios, evap_mode="dense", state_every=1, use_host_api=False, noise_mode="tape", rng_seed=7,
               env_id_base=0, record="f64"):
    """scenarios: list of (cfg, init, tape) sharing cfg and tape length.  Runs main.py's loop order
    (observation; T x [step; update]) on the GPU batch and on one oracle per env, comparing everything."""
    import torch
    from antsrl_b200 import BatchedAnts
    from oracle.antsrl_oracle import philox_uniform
    cfg = scenarios[0][0]
    E = len(scenarios)
    oracles = [OracleEnv(c, i) for c, i, _ in scenarios]
    batch = BatchedAnts(cfg, E, evap_mode=evap_mode, rng_seed=rng_seed, env_id_base=env_id_base, record=record)
    batch.import_state(stack_init(cfg, [i for _, i, _ in scenarios]))
    dev = batch.device
    T = scenarios[0][2]["rot"].shape[0]

    def cmp_outputs(gpu, ref_list, what):
        obs, ast, rew = [_np(g) for g in gpu]
        assert_close(obs, np.stack([r[0] for r in ref_list]), what + ": obs")
        assert_close(ast, np.stack([r[1] for r in ref_list]), what + ": agent_state")
        assert_close(rew, np.stack([r[2] for r in ref_list]), what + ": reward")

    # main.py:88 -- the extra observation
    ref0 = []
    for o in oracles:
        p_, a_, s_ = o.observation()
        ref0.append((p_, a_, o.s["rewards"].copy(), s_))
    if use_host_api:
        obs, ast, st, rew = batch.observe_host()
    else:
        obs, ast, st, rew = batch.observe()
    cmp_outputs((obs, ast, rew), ref0, "observation0")
    assert_close(_np(st), np.stack([r[3] for r in ref0]), "observation0: state")
    compare_state(batch.export_state(), oracles, "after observation0", cfg)
    n_checked = 0
    for t in range(T):
        rot_none = bool(scenarios[0][2]["rot_none"][t])
        ph_none = bool(scenarios[0][2]["ph_none"][t])
        rot = np.stack([s[2]["rot"][t] for s in scenarios]).astype(np.int8)
        ph = np.stack([s[2]["ph"][t] for s in scenarios]).astype(np.int8)
        ref = []
        for e, o in enumerate(oracles):
            p_, a_, r_, d_ = o.step(None if rot_none else rot[e].astype(np.int64),
                                    None if ph_none else ph[e].astype(np.int64))
            ref.append((p_, a_, r_.copy(), d_))
        if use_host_api:
            obs, ast, rew, done = batch.step_host(None if rot_none else rot, None if ph_none else ph)
        else:
            obs, ast, rew, done = batch.step(None if rot_none else torch.from_numpy(rot).to(dev),
                                             None if ph_none else torch.from_numpy(ph).to(dev))
        cmp_outputs((obs, ast, rew), ref, "step %d" % t)
        assert done == ref[0][3], "done at step %d" % t
        check_now = state_every and t % state_every == 0
        if check_now:
            compare_state(batch.export_state(), oracles, "after step %d" % t, cfg)
        if noise_mode == "tape":
            noise = np.stack([s[2]["noise"][t] for s in scenarios])
            for e, o in enumerate(oracles):
                o.update(noise[e])
            if use_host_api:
                batch.update_host(noise)
            else:
                batch.update(torch.from_numpy(noise).to(dev))
        else:   # in-kernel Philox keyed by (seed, global env id, timestep, ant)
            for e, o in enumerate(oracles):
                o.update(philox_uniform(rng_seed, env_id_base + e, int(o.s["timestep"]), cfg["n_ants"]))
            batch.update(None)
        if check_now:
            compare_state(batch.export_state(), oracles, "after update %d" % t, cfg)
            n_checked += 1
    compare_state(batch.export_state(), oracles, "final", cfg)
    stats = batch.stats()
    batch.close()
    return {"envs": E, "steps": T, "state_checks": n_checked, "kernel_launches": stats["kernel_launches"]}

"""The packed observation form of the host-buffer path (include/antsrl_b200.h: AntsPackedLayout, ants_unpack_obs,
ants_step_host_packed; antsrl_b200/csrc/ants_pack.cuh): the host expander against a numpy statement of the format
(CPU), and on the GPU packed transport == dense copy, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from antsrl_b200 import _cabi
from antsrl_b200.batch import build_c_config, make_config


def _layout(cfg):
    lib = _cabi.load_library()
    c = build_c_config(cfg, 1)
    L = _cabi.AntsPackedLayout()
    _cabi.check(lib, lib.ants_packed_layout(C.byref(c), C.byref(L)))
    return lib, L


def _random_dense(cfg, n_ants, rng):
    """A dense observation obeying what RLApi.observation can produce: 0/1 channels, f32 pheromones, integer food,
    -1 in masked samples."""
    s = 2 * cfg["radius"] + 1
    obs = np.zeros((n_ants, s, s, len(cfg["channels"])), np.float32)
    for c, name in enumerate(cfg["channels"]):
        if name.startswith("phero"):
            v = rng.random_sample((n_ants, s, s)).astype(np.float32)
            v[rng.random_sample(v.shape) < 0.5] = 0
            v[rng.random_sample(v.shape) < 0.1] = 1.0
            obs[..., c] = v
        elif name == "food":
            obs[..., c] = rng.randint(0, 7, (n_ants, s, s)) * (rng.random_sample((n_ants, s, s)) < 0.3)
        else:
            obs[..., c] = rng.random_sample((n_ants, s, s)) < 0.3
    if cfg["mask"] is not None:
        obs[:, ~np.asarray(cfg["mask"], bool), :] = -1.0
    return obs


def numpy_pack(L, obs):
    """The format of ants_pack.cuh restated: per ant n_visible samples of { f32 a; f32 b; u16 food; u8 flags; u8 0 }."""
    n = obs.shape[0]
    flat = obs.reshape(n, L.n_samples, L.n_channels)
    vis = np.array(list(L.visible_index)[:L.n_visible])
    rec = np.zeros((n, L.n_visible), dtype=np.dtype([("a", "<f4"), ("b", "<f4"), ("food", "<u2"), ("flags", "u1"), ("pad", "u1")]))
    if L.value_channel[0] >= 0:
        rec["a"] = flat[:, vis, L.value_channel[0]]
    if L.value_channel[1] >= 0:
        rec["b"] = flat[:, vis, L.value_channel[1]]
    if L.food_channel >= 0:
        rec["food"] = flat[:, vis, L.food_channel].astype(np.uint16)
    for k in range(8):
        if L.flag_channel[k] >= 0:
            rec["flags"] |= (flat[:, vis, L.flag_channel[k]] != 0).astype(np.uint8) << k
    assert rec.dtype.itemsize == L.sample_bytes == 12 and L.bytes_per_ant == 12 * L.n_visible
    return rec


CONFIGS = [
    ("default_c7", dict(n_rocks=3)),
    ("default_c6", dict()),
    ("no_mask_r2", dict(radius=2, mask=None)),
    ("permuted", dict(channels=["food", "walls", "phero1", "anthill", "ants", "phero0"])),
    ("one_phero", dict(n_phero=1, channels=["phero0", "food"], radius=1, mask=None)),
    ("no_phero_c4", dict(n_phero=0, channels=["ants", "walls", "food", "anthill"])),
    ("c8", dict(n_rocks=2, channels=["rocks", "ants", "phero0", "walls", "food", "phero1", "anthill", "ants"])),
    ("last_sample_visible", dict(mask=np.array([[1, 0, 1, 0, 1, 0, 1]] * 7, bool))),
]


@pytest.mark.parametrize("name,kw", CONFIGS, ids=[n for n, _ in CONFIGS])
def test_host_expander_rebuilds_the_dense_observation(name, kw):
    cfg = make_config(64, 64, 10, **kw)
    lib, L = _layout(cfg)
    assert L.supported == 1
    assert L.n_visible == (int(np.asarray(cfg["mask"]).sum()) if cfg["mask"] is not None else (2 * cfg["radius"] + 1) ** 2)
    rng = np.random.RandomState(5)
    for n_ants, threads in ((1, 1), (5, 1), (1000, 0), (5003, 0), (4096, 3)):
        obs = _random_dense(cfg, n_ants, rng)
        rec = numpy_pack(L, obs)
        out = np.full(obs.shape, 7.0, np.float32)                       # garbage the expander must overwrite entirely
        _cabi.check(lib, lib.ants_unpack_obs(C.byref(L), rec.ctypes.data_as(C.c_void_p), n_ants,
                                             out.ctypes.data_as(C.c_void_p), threads))
        assert np.array_equal(out.view(np.uint32), obs.view(np.uint32)), "%s: %d ants" % (name, n_ants)


def test_layouts_without_a_packed_form():
    """More than two pheromone channels (or more than 8 channels) have no packed form: ants_step_host copies dense."""
    cfg = make_config(32, 32, 4, n_phero=3, channels=["phero0", "phero1", "phero2", "food"])
    _, L = _layout(cfg)
    assert L.supported == 0


@pytest.mark.gpu
@pytest.mark.parametrize("scen_kw", [dict(n_rocks=4), dict(), dict(float_food=True), dict(radius=2, mask=None, fwd_delta=0),
                                     dict(channels=["food", "walls", "phero1", "anthill", "ants", "phero0"])],
                         ids=["rocks_c7", "c6", "float_food_falls_back", "no_mask_r2", "permuted"])
def test_packed_transport_equals_dense_copy(scen_kw, monkeypatch):
    """ants_step_host through the packed transport (forced for this small batch) returns the arrays the dense copy
    returns, bit for bit, step after step; a non-integer amount of food switches the handle to dense copies.  Then
    ants_step_host_packed + ants_unpack_obs on a third handle."""
    from antsrl_b200 import BatchedAnts, AntsError
    from parity_util import stack_init
    from scenarios import make_scenario
    T = 25
    scen = [make_scenario(seed=1300 + e, w=56, h=48, n_ants=150, steps=T, n_walls=5, n_food=10, **scen_kw) for e in range(3)]
    cfg = scen[0][0]
    runs = {}
    for mode in ("dense", "packed"):
        monkeypatch.delenv("ANTS_E2E_DENSE", raising=False)
        monkeypatch.delenv("ANTS_E2E_PACKED", raising=False)
        monkeypatch.setenv("ANTS_E2E_DENSE" if mode == "dense" else "ANTS_E2E_PACKED", "1")
        b = BatchedAnts(cfg, 3, evap_mode="lazy", record="compact8")
        b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
        b.observe_host()
        out = []
        for t in range(T):
            rot = np.stack([s[2]["rot"][t] for s in scen]).astype(np.int8)
            ph = np.stack([s[2]["ph"][t] for s in scen]).astype(np.int8)
            obs, ast, rew, done = b.step_host(rot, ph)
            out.append((obs.copy(), ast.copy(), rew.copy(), done))
            b.update_host(np.stack([s[2]["noise"][t] for s in scen]))
        runs[mode] = out
        if mode == "packed":
            launched = b.kernel_ms()["pack"][1]
            assert launched > 0, "the packed transport must have run"
            if scen_kw.get("float_food"):
                assert launched < T, "a non-integer food amount must have switched the handle to dense copies"
        b.close()
    for t in range(T):
        for a, c in zip(runs["dense"][t][:3], runs["packed"][t][:3]):
            assert np.array_equal(a.view(np.uint32 if a.dtype == np.float32 else np.uint64),
                                  c.view(np.uint32 if c.dtype == np.float32 else np.uint64)), "step %d" % t
        assert runs["dense"][t][3] == runs["packed"][t][3]
    # the caller-side pair: packed buffer out, expanded by ants_unpack_obs
    monkeypatch.delenv("ANTS_E2E_PACKED", raising=False)
    b = BatchedAnts(cfg, 3, evap_mode="lazy", record="compact8")
    b.import_state(stack_init(cfg, [i for _, i, _ in scen]))
    b.observe_host()
    for t in range(T):
        rot = np.stack([s[2]["rot"][t] for s in scen]).astype(np.int8)
        ph = np.stack([s[2]["ph"][t] for s in scen]).astype(np.int8)
        try:
            packed, ast, rew, done = b.step_host_packed(rot, ph)
        except AntsError:
            assert scen_kw.get("float_food"), "only a non-integer amount of food has no packed form"
            break
        obs = b.unpack_obs(packed)
        assert np.array_equal(obs.view(np.uint32), runs["dense"][t][0].view(np.uint32)), "packed API step %d" % t
        assert np.array_equal(rew, runs["dense"][t][2])
        b.update_host(np.stack([s[2]["noise"][t] for s in scen]))
    b.close()

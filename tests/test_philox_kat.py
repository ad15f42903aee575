"""Known-answer test of the counter-based generator behind the in-kernel collision noise and the on-device action
sampler: Philox4x32-10 with the published Random123 vectors (kat_vectors of the Random123 distribution, Salmon et al.
SC'11).  The CUDA path is compared with oracle.philox_uniform / philox_actions on the GPU
(tests/test_gpu_parity.py::test_philox_noise_matches_oracle, ::test_device_action_sampler_matches_oracle_and_sharding),
so this pins the device generator to the standard as well."""
import numpy as np

from oracle.antsrl_oracle import philox4x32_10, philox_actions, philox_uniform

KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2, (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(*(np.array([c], dtype=np.uint64) for c in ctr), key[0], key[1])
        assert tuple(int(v[0]) for v in got) == want


def test_uniforms_and_actions_are_functions_of_the_counter():
    u = philox_uniform(0x1234567890ABCDEF, 17, 3, 1000)
    assert u.dtype == np.float64 and (u >= 0).all() and (u < 1).all() and 0.45 < u.mean() < 0.55
    # keyed by (seed, env, step, ant): an env's draws do not depend on what else is in the batch or on n_ants
    assert np.array_equal(philox_uniform(0x1234567890ABCDEF, 17, 3, 10), u[:10])
    assert not np.array_equal(philox_uniform(0x1234567890ABCDEF, 18, 3, 10), u[:10])
    assert not np.array_equal(philox_uniform(0x1234567890ABCDEF, 17, 4, 10), u[:10])
    # counter (ant 0, step 0, env 0, stream 0) under key 0 is the first KAT vector: u = ((r0>>5)*2^26 + (r1>>6)) / 2^53
    r0, r1 = 0x6627e8d5, 0xe169c58d
    assert philox_uniform(0, 0, 0, 1)[0] == ((r0 >> 5) * 67108864.0 + (r1 >> 6)) / 9007199254740992.0
    rot, ph = philox_actions(77, 10, 5, 30000)
    assert set(np.unique(rot)) == {-1, 0, 1} and set(np.unique(ph)) == {0, 1, 2}
    for v in (-1, 0, 1):
        assert abs((rot == v).mean() - 1 / 3) < 0.01
    rot5, ph4 = philox_actions(77, 10, 5, 30000, 5, 4)
    assert set(np.unique(rot5)) == {-2, -1, 0, 1, 2} and set(np.unique(ph4)) == {0, 1, 2, 3}

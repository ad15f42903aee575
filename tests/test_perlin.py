"""Perlin walls.  The third-party `noise` package the reference calls is absent and
the reference holds no vectors for it: parity is UNPINNED; these are known-answer properties of the published
algorithm (improved Perlin gradient noise, `noise` 1.2.x `_perlin.c`) that the restatement must satisfy."""
import random

import numpy as np

from antsrl_b200.perlin import PerlinGenerator, noise2, perlin_noise_generator, pnoise2


def test_zero_on_the_integer_lattice():
    # every gradient contribution vanishes at a lattice point; with lacunarity 2 all octaves hit lattice points too
    xs, ys = np.meshgrid(np.arange(-40, 40, dtype=np.float32), np.arange(-37, 41, dtype=np.float32), indexing="ij")
    assert np.all(noise2(xs, ys) == 0)
    assert np.all(pnoise2(xs, ys, octaves=3) == 0)


def test_bounds_and_smoothness():
    rs = np.random.RandomState(1)
    x = (rs.random_sample(200000) * 2000 - 1000).astype(np.float32)
    y = (rs.random_sample(200000) * 2000 - 1000).astype(np.float32)
    n = noise2(x, y)
    assert n.dtype == np.float32 and np.abs(n).max() <= 1.0 and np.abs(n).max() > 0.6
    assert abs(float(n.mean())) < 0.01                       # unbiased
    # C2-continuous: a step of 1e-3 moves the value by at most ~ (max gradient 2 * sqrt 2 ... ) * 1e-3
    d = np.abs(noise2(x + np.float32(1e-3), y) - n)
    assert d.max() < 1e-2
    # across a cell boundary too
    xb = np.float32(17.0)
    eps = np.float32(1e-4)
    yy = np.linspace(3.1, 3.9, 50).astype(np.float32)
    assert np.abs(noise2(xb - eps, yy) - noise2(xb + eps, yy)).max() < 1e-3


def test_one_octave_known_answers():
    # inside cell (0, 0) with base 0: corners hash to PERM[PERM[PERM[i] + j]] & 15 =
    #   (0,0): PERM[PERM[151 + 0]] = PERM[PERM[151]] ...  evaluated here from the tables by the textbook formula
    from antsrl_b200.perlin import _PERM, _GX, _GY

    def ref(x, y):
        i, j = int(np.floor(x)) & 255, int(np.floor(y)) & 255
        fx, fy = np.float32(x - np.floor(x)), np.float32(y - np.floor(y))
        fade = lambda t: t * t * t * (t * (t * np.float32(6) - np.float32(15)) + np.float32(10))
        g = lambda h, a, b: a * _GX[h & 15] + b * _GY[h & 15]
        h00 = _PERM[_PERM[_PERM[i] + j]]
        h10 = _PERM[_PERM[_PERM[(i + 1) & 255] + j]]
        h01 = _PERM[_PERM[_PERM[i] + ((j + 1) & 255)]]
        h11 = _PERM[_PERM[_PERM[(i + 1) & 255] + ((j + 1) & 255)]]
        one = np.float32(1)
        a = g(h00, fx, fy) + fade(fx) * (g(h10, fx - one, fy) - g(h00, fx, fy))
        b = g(h01, fx, fy - one) + fade(fx) * (g(h11, fx - one, fy - one) - g(h01, fx, fy - one))
        return a + fade(fy) * (b - a)

    for x, y in [(0.5, 0.5), (3.25, 7.75), (100.1, 255.9), (255.5, 0.5), (12.0625, 200.5)]:
        assert noise2(np.float32(x), np.float32(y)) == np.float32(ref(np.float32(x), np.float32(y)))


def test_octave_sum_and_period():
    x = np.linspace(-5, 5, 101).astype(np.float32)[:, None]
    y = np.linspace(2, 9, 71).astype(np.float32)[None, :]
    two = (noise2(x, y) + noise2(x * np.float32(2), y * np.float32(2), 2048.0, 2048.0) * np.float32(0.5)) / np.float32(1.5)
    assert np.array_equal(pnoise2(x, y, octaves=2, persistence=0.5, lacunarity=2.0), two.astype(np.float32))
    # the permutation repeats every 256 cells
    assert np.allclose(noise2(x, y), noise2(x + np.float32(256), y + np.float32(512)), atol=1e-4)


def test_generator_interface_and_rng_draws():
    random.seed(11)
    g = PerlinGenerator()
    walls = g.generate(200, 200)                           # main.py:75 default walls of the reference map
    after = random.random()
    assert walls.shape == (200, 200) and walls.dtype == bool
    assert 0.15 < walls.mean() < 0.6                       # blobs above the density threshold, neither empty nor full
    random.seed(11)
    ox, oy = random.randint(-10000, 10000), random.randint(-10000, 10000)   # exactly the reference's two draws
    assert random.random() == after
    ref = perlin_noise_generator(200, 200, ox, oy) > 0.05
    assert np.array_equal(walls, ref)
    # blobs, not salt and pepper: most wall cells have a wall neighbour
    nb = walls[1:-1, 1:-1] & (walls[:-2, 1:-1] | walls[2:, 1:-1] | walls[1:-1, :-2] | walls[1:-1, 2:])
    assert nb.sum() > 0.95 * walls[1:-1, 1:-1].sum()


def test_perlin_walls_feed_the_generator():
    from antsrl_b200.generator import CirclesGenerator, generate_state
    st = generate_state(96, 80, 20, 2, 0, CirclesGenerator(6, 3, 6), PerlinGenerator(scale=12.0), seed=5)
    assert st["walls"].shape == (96, 80) and st["walls"].any()
    ax, ay, ar = st["anthill_xyr"]
    assert not st["walls"][ax, ay]                          # cleared inside the anthill (environment_generator.py:66-68)
    assert not (st["food"].astype(bool) & st["walls"].astype(bool)).any()

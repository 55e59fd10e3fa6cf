"""CPU: the numpy-only host helpers -- synthetic inputs (generate.py) and the statistics the parallel-mode
parity checks use (analysis.py)."""
import numpy as np

from simplewinderosion_b200 import analysis, generate
from tests.oracle_lib import PortWorld


def test_gaussian_hill_equals_map_init():
    """generate.gaussian_hill restates cellpool.h:378-385,416-428; numpy's exp may differ from libm's in
    the last bit of a double, which the float rounding almost always hides: allow 1 ulp."""
    n = 64
    w = PortWorld(n)
    w.init_gaussian()
    a, b = generate.gaussian_hill(n), np.array(w.plane(0))
    assert a.shape == b.shape and a.max() == 0.5 and 0.0 < a.min() < 1e-4   # min/max fold from 0.0: the floor stays
    assert np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max() <= 1


def test_synthetic_inputs_have_the_documented_shape():
    f = generate.flat_sand(32, 48, seed=1)
    assert f.shape == (32, 48) and f.dtype == np.float32 and abs(float(f.mean()) - 0.25) < 1e-3 and float(np.ptp(f)) <= 2.1e-3
    r = generate.random_slopes(16, seed=3)
    assert r.shape == (16, 16) and 0.0 <= r.min() and r.max() < 1.0
    p = generate.pyramid(64, apex=0.5)
    assert p.max() <= 0.5 and p[0, 0] == 0.0 and p[32, 32] == p.max() and (p[32, 28:36] > 0).all()
    assert np.array_equal(p, p[::-1][:, ::-1][np.ix_(np.r_[63, 0:63], np.r_[63, 0:63])]) or p[31, 32] == p[33, 32]
    n = generate.normalise_half(np.array([[-1.0, 0.0], [1.0, 3.0]], np.float32))
    assert n.min() == 0.0 and n.max() == 0.5 and n[0, 1] == np.float32(0.125)


def test_spectrum_finds_a_planted_dune_wavelength():
    n = 256
    x = np.arange(n, dtype=np.float64)[:, None]
    rs = np.random.RandomState(0)
    h = (0.1 * np.sin(2 * np.pi * x / 32.0) + 0.01 * rs.standard_normal((n, n))).astype(np.float32)
    assert analysis.dominant_wavelength(h) == 32.0
    # a strong along-wind trend (sand draining towards the downwind edge) must not hide the dunes
    trend = (h + 8.0 * (x / n) ** 2 - 3.0 * x / n).astype(np.float32)
    assert analysis.dominant_wavelength(trend) == 32.0
    assert analysis.dominant_wavelength(trend, kmin=1, detrend=False) > 64.0       # what round 1 reported: n/2, n/3
    p = analysis.wind_spectrum(h)
    assert int(np.argmax(p[1:])) + 1 == n // 32
    shifted = np.roll(h, 5, axis=0)
    assert analysis.spectrum_similarity(h, shifted, 1, n // 4) > 0.999     # phase does not matter
    assert analysis.spectrum_similarity(h, rs.standard_normal((n, n)), 1, n // 4) < 0.9


def test_histogram_distance_is_a_metric_like_score():
    rs = np.random.RandomState(1)
    a = rs.standard_normal((128, 128))
    assert analysis.histogram_distance(a, a) == 0.0
    assert analysis.histogram_distance(a, a + 10.0) > 0.99
    b = rs.standard_normal((128, 128))
    assert analysis.histogram_distance(a, b) < 0.1
    p, edges = analysis.height_histogram(a, bins=16)
    assert abs(p.sum() - 1.0) < 1e-12 and len(edges) == 17

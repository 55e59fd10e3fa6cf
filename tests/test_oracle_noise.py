"""CPU: the oracle's restatement of the map::init noise stack (cellpool.h:366-368,388-411,416-428 +
FastNoiseLite.h OpenSimplex2 FBm) against golden values produced by the UNMODIFIED FastNoiseLite
class (tests/golden/noise.npz, tests/golden/make_noise_golden.py) and, where oracle/_ref is
present, against that class directly. Bit-exact: the noise is fp32 + wrapping int arithmetic."""
import os

import numpy as np
import pytest

from tests.oracle_lib import PortWorld, RefWorld, ref_available

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noise.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_pointwise_getnoise_matches_golden():
    w = PortWorld(8)
    got = np.array([w.get_noise(*map(float, p)) for p in G["points"]], dtype=np.float32)
    assert np.array_equal(bits(got), bits(G["values"]))
    assert np.abs(G["values"]).max() <= 1.0 and np.abs(G["values"]).max() > 0.3    # bounded, not degenerate


@pytest.mark.parametrize("seed", [1, 4711])
def test_stack_and_normalise_match_golden(seed):
    w = PortWorld(64)
    w.noise_stack(seed, 8)
    assert np.array_equal(bits(w.plane(0)), bits(G[f"stack_raw_seed{seed}"]))
    w.normalize()
    assert np.array_equal(bits(w.plane(0)), bits(G[f"stack_norm_seed{seed}"]))
    h = w.plane(0)
    assert h.min() == 0.0 and h.max() == 0.5


def test_stack_adds_to_the_hill():
    w = PortWorld(64)
    w.init_gaussian()
    assert np.array_equal(bits(w.plane(0)), bits(G["hill"]))
    w.noise_stack(1, 8)
    w.normalize()
    assert np.array_equal(bits(w.plane(0)), bits(G["hill_plus_stack_norm_seed1"]))


def test_layers_accumulate_one_by_one():
    a, b = PortWorld(32), PortWorld(32)
    a.noise_stack(3, 3)
    # layer o alone has frequency 2^o and scale 0.6^(o+1): the first layer equals a 1-layer stack
    b.noise_stack(3, 1)
    assert not np.array_equal(bits(a.plane(0)), bits(b.plane(0)))
    want = np.float32(0.6) * np.array([[b.get_noise(1.0, x / 32, z / 32, 3.0) for z in range(32)] for x in range(32)],
                                      dtype=np.float32)
    assert np.array_equal(bits(b.plane(0)), bits(want.astype(np.float32)))


@pytest.mark.skipif(not ref_available(32), reason="oracle/_ref not built")
def test_port_equals_unmodified_fastnoiselite():
    r = RefWorld(32)
    if not hasattr(r.L, "ref_get_noise"):
        pytest.skip("oracle/_ref predates the noise harness")
    w = PortWorld(32)
    rs = np.random.RandomState(5)
    for _ in range(20000):
        f = float(2 ** rs.randint(0, 10))
        x, y, z = (float(np.float32(v)) for v in rs.uniform(-300, 300, 3))
        assert np.float32(r.get_noise(f, x, y, z)).view(np.uint32) == np.float32(w.get_noise(f, x, y, z)).view(np.uint32)
    for seed in (0, 9999, 10001, 123456789):
        zero = np.zeros((32, 32), np.float32)
        r.set_plane(0, zero); w.set_plane(0, zero)
        r.noise_stack(seed, 8); w.noise_stack(seed, 8)
        assert np.array_equal(bits(r.plane(0)), bits(w.plane(0)))
        r.normalize(); w.normalize()
        assert np.array_equal(bits(r.plane(0)), bits(w.plane(0)))


def test_banded_multithreaded_map_equals_whole_map():
    """tests.oracle_lib.noise_map (what bench.py --impl reference feeds the CPU arm): row bands on
    several threads, then the normalisation -- the same bits as the one-piece stack."""
    from tests.oracle_lib import noise_map
    got = noise_map(64, seed=1, layers=8, workers=3)
    assert np.array_equal(bits(got), bits(G["stack_norm_seed1"]))
    rag = noise_map(37, 50, seed=4711, layers=5, workers=2)
    w = PortWorld(37, 50)
    w.noise_stack(4711, 5)
    w.normalize()
    assert np.array_equal(bits(rag), bits(w.plane(0)))

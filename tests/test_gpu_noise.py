"""GPU: quad::map::init's noise layers generated on the device (swe_generate_noise,
swe_height_range, swe_normalize_height) against the oracle restatement and the golden values of
the UNMODIFIED FastNoiseLite class (tests/golden/noise.npz). Bar: BIT-EXACT (fp32 + wrapping int
arithmetic; the kernels are compiled without FMA contraction)."""
import os

import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import stripes
from tests.oracle_lib import PortWorld

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noise.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("seed", [1, 4711])
def test_stack_matches_unmodified_fastnoiselite_golden(seed):
    w = swe.World(64)
    w.generate_noise(seed, 8)
    assert np.array_equal(bits(w.download()), bits(G[f"stack_raw_seed{seed}"]))
    lo, hi = w.height_range()
    raw = G[f"stack_raw_seed{seed}"]
    assert lo == min(0.0, float(raw.min())) and hi == max(0.0, float(raw.max()))
    w.normalize()
    assert np.array_equal(bits(w.download()), bits(G[f"stack_norm_seed{seed}"]))


def test_stack_adds_to_uploaded_hill():
    w = swe.World(64)
    w.upload(G["hill"])
    w.generate_noise(1, 8)
    w.normalize()
    assert np.array_equal(bits(w.download()), bits(G["hill_plus_stack_norm_seed1"]))


@pytest.mark.parametrize("nx,nz,tile,layers,seed", [(96, 160, 0, 8, 3), (257, 130, 100, 5, 12345),
                                                   (33, 7, 64, 16, 9999), (512, 512, 0, 8, 20001)])
def test_stack_matches_oracle_on_ragged_maps(nx, nz, tile, layers, seed):
    ora = PortWorld(nx, nz)
    ora.noise_stack(seed, layers, tile or None)
    w = swe.World(nx, nz)
    w.generate_noise(seed, layers, tile)
    assert np.array_equal(bits(w.download()), bits(ora.plane(0)))
    ora.normalize()
    w.normalize()
    assert np.array_equal(bits(w.download()), bits(ora.plane(0)))


def test_argument_errors():
    w = swe.World(16)
    with pytest.raises(swe.SweError):
        w.generate_noise(1, 17)
    with pytest.raises(swe.SweError):
        w.generate_noise(1, 8, -3)
    with pytest.raises(swe.SweError):
        w.normalize()                 # flat map: the reference divides by zero (cellpool.h:427)
    w.generate_noise(1, 0)            # zero layers: nothing to do
    assert not w.download().any()


def test_full_size_map_sampled_against_oracle():
    """BASELINE cfg3 size: 4096^2, eight layers; 20,000 sampled cells against the oracle's pointwise
    noise, plus the normalisation invariants (min 0, max 0.5)."""
    n, seed = 4096, 1
    w = swe.World(n)
    w.generate_noise(seed, 8)
    h = w.download()
    ora = PortWorld(8)
    rs = np.random.RandomState(11)
    xs, zs = rs.randint(0, n, 20000), rs.randint(0, n, 20000)
    scale, freq = np.float32(0.6), np.float32(1.0)
    want = np.zeros(len(xs), np.float32)
    for _ in range(8):
        layer = np.array([ora.get_noise(float(freq), float(np.float32(x) / np.float32(n)),
                                        float(np.float32(z) / np.float32(n)), float(seed % 10000))
                          for x, z in zip(xs, zs)], dtype=np.float32)
        want = (want + scale * layer).astype(np.float32)
        freq = np.float32(freq * 2)
        scale = np.float32(np.float64(scale) * 0.6)
    assert np.array_equal(bits(h[xs, zs]), bits(want))
    w.normalize()
    h = w.download()
    assert h.min() == 0.0 and h.max() == 0.5
    assert 0.01 < float(h.std()) < 0.2


@pytest.mark.parametrize("nx,nz,world", [(64, 48, 2), (130, 64, 5)])
def test_stripes_generate_their_rows_and_ghosts(nx, nz, world):
    single = swe.World(nx, nz)
    single.generate_noise(5, 8)
    ws = [stripes.StripedWorld(nx, nz, r, world, 1024) for r in range(world)]
    stripes.StripedWorld.connect_in_process(ws)
    for w in ws:
        w.generate_noise(5, 8)
    got = stripes.gather_global(ws)
    assert np.array_equal(bits(got), bits(single.download()))
    lo = min(w.height_range()[0] for w in ws)
    hi = max(w.height_range()[1] for w in ws)
    assert (lo, hi) == single.height_range()
    for w in ws:
        w.normalize(lo, hi)
    single.normalize()
    assert np.array_equal(bits(stripes.gather_global(ws)), bits(single.download()))
    # ghost rows were generated and normalised too: a whole-map sweep needs no halo sync first
    g = stripes.StripeGroup(ws)
    g.cascade_sweep(1)
    single.cascade_sweep(1)
    assert np.array_equal(bits(g.gather()), bits(single.download()))


def test_cfg1_variant_noise_map_then_deterministic_erode():
    """BASELINE configs[0] in SURVEY 8(d)'s 256^2 noise-map variant: heightmap = the eight noise layers
    (SEED 1) normalised, generated on the device, then World::erode(8*256) frames in deterministic
    mode -- every plane bitwise equal to the oracle fed with ITS OWN restatement of the generator."""
    n, frames = 256, 2
    ora = PortWorld(n)
    ora.noise_stack(1, 8)
    ora.normalize()
    ora.srand(1)
    w = swe.World(n)
    w.generate_noise(1, 8)
    w.normalize()
    w.srand(1)
    assert np.array_equal(bits(w.download()), bits(ora.plane(0)))
    steps = 0
    for _ in range(frames):
        w.erode(8 * n, swe.MODE_DETERMINISTIC)
        steps += ora.erode(8 * n).steps
    assert w.stats().steps == steps
    for k in range(9):
        assert np.array_equal(bits(w.download(k)), bits(ora.plane(k))), k

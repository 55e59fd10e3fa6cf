"""GPU: the TILE-RESIDENT schedule of the parallel mode (SWE_OPT_SCHEDULE = 2, csrc/swe_tile.cu) against the
sequential reference (oracle port = the unmodified reference bit for bit, tests/test_oracle_ref.py) at the
benchmarked particle densities, with the tolerances of tests/test_gpu_density.py / test_gpu_parallel.py:
    steps per particle +-3 %, exported sand +-5 %, height-histogram TV < 0.12, log-spectrum corr > 0.99,
    erosion-field corr > 0.95 (structured terrain), along-wind profile corr > 0.9, mass balance < 2e-6.
Plus what is specific to tiles: maps that are not multiples of the tile edge, non-square maps, tiles whose
particle segment exceeds the shared-memory capacity (the excess is parked, never lost), the layered
extension, options."""
import os

import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import analysis, generate
from tests.oracle_lib import PortWorld, noise_map
from tests.test_gpu_density import compare

pytestmark = pytest.mark.gpu
TILES = {swe.OPT_SCHEDULE: swe.SCHEDULE_TILES}


@pytest.mark.parametrize("shape", [2, 1, 0], ids=["240x60", "120x120", "60x60"])
def test_tiles_headline_density_quarter_particle_per_cell(shape):
    """1024^2 / 262,144 particles = 0.25 per cell: the density of bench.py's headline (4096^2 / 4,194,304), with the
    tile geometry the headline runs (240 x 60 cells: elongated along the wind, 24 rounds per super-round), the square
    one of the same size and the small one."""
    n, cycles = 1024, 1 << 18
    hm = noise_map(n, seed=1, layers=8, workers=8)
    r = compare(n, cycles, hm, opts={**TILES, swe.OPT_TILE_SHAPE: shape})
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r
    assert abs(r["massflow"]) < 0.1, r
    assert r["drift"] < 2e-6, r


@pytest.mark.parametrize("shape", [0, 1, 2, 3], ids=["60x60", "120x120", "240x60", "60x240"])
def test_tiles_quarter_particle_per_cell_512(shape):
    """Every compiled tile geometry (a 512^2 map would pick 60 x 60 by itself; 4096^2 with the wind along x picks 240 x 60;
    60 x 240 is what a wind along z would pick -- here it is elongated ACROSS the wind, the unfavourable case)."""
    n, cycles = 512, 1 << 16
    hm = noise_map(n, seed=1, layers=8, workers=4)
    r = compare(n, cycles, hm, opts={**TILES, swe.OPT_TILE_SHAPE: shape})
    assert abs(r["steps"]) < 0.03 and abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r
    assert r["drift"] < 2e-6, r


def test_tiles_cfg2_density_one_particle_per_cell_flat_sand():
    n, cycles = 256, 1 << 16
    hm = generate.flat_sand(n, seed=42)
    r = compare(n, cycles, hm, opts=TILES)
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["profile"] > 0.9, r
    assert r["drift"] < 2e-6, r


def test_tiles_one_particle_per_cell_structured_terrain_two_calls():
    n, cycles = 256, 1 << 16
    hm = noise_map(n, seed=1, layers=8, workers=4)
    r = compare(n, cycles, hm, calls=2, opts=TILES)
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r


def test_tiles_sparse_reference_frame():
    """The reference's own frame: 8 * tilesize particles (0.016 per cell), injected at once."""
    n = 512
    hm = noise_map(n, seed=1, layers=8, workers=4)
    r = compare(n, 8 * n, hm, calls=4, opts=TILES)
    assert abs(r["steps"]) < 0.03 and abs(r["export"]) < 0.08, r      # (4 x 4096 particles: +-3 % sampling noise on the export)
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["drift"] < 2e-6, r


@pytest.mark.parametrize("shape", [0, 1, 2, 3], ids=["60x60", "120x120", "240x60", "60x240"])
@pytest.mark.parametrize("nx,nz", [(100, 100), (61, 59), (200, 90), (90, 333), (20, 500)])
def test_tiles_odd_map_shapes_counters_and_mass_balance(nx, nz, shape):
    """Maps that are not multiples of the tile edge (60), smaller than a tile, and non-square: every particle is
    accounted for, the mass balance closes, the result is statistically the colour-queue scheduler's."""
    rs = np.random.RandomState(nx * 7 + nz)
    hm = (0.3 * generate.fbm(max(nx, nz), seed=3)[:nx, :nz] + 0.01 * rs.random_sample((nx, nz))).astype(np.float32)
    cycles = nx * nz // 2
    out = {}
    for name, sched in (("tiles", swe.SCHEDULE_TILES), ("queues", swe.SCHEDULE_COLOUR_QUEUES)):
        w = swe.World(nx, nz); w.set_option(swe.OPT_SCHEDULE, sched); w.set_option(swe.OPT_TILE_SHAPE, shape)
        w.upload(hm); w.srand(5)
        for _ in range(2):
            w.erode(cycles, swe.MODE_PARALLEL)
        st = w.stats()
        assert st.particles == 2 * cycles == st.exits_oob + st.exits_age, (name, st.particles, st.exits_oob, st.exits_age)
        assert st.queue_dropped == 0
        hg = w.download()
        assert np.isfinite(hg).all()
        drift = w.sum_plane(swe.HEIGHT) + st.sediment_lost - float(hm.astype(np.float64).sum())
        assert abs(drift) < 2e-6 * float(np.abs(hm.astype(np.float64)).sum()) + 1e-5, (name, drift)
        track = w.download(swe.MASSFLOW_TRACK)
        assert np.isfinite(track).all() and float(track.min()) >= 0.0
        out[name] = (st.steps, st.sediment_lost, hg)
        w.close()
    assert abs(out["tiles"][0] / out["queues"][0] - 1.0) < 0.05, (out["tiles"][0], out["queues"][0])
    assert abs(out["tiles"][1] / out["queues"][1] - 1.0) < 0.15, (out["tiles"][1], out["queues"][1])
    assert np.corrcoef((out["tiles"][2] - hm).ravel(), (out["queues"][2] - hm).ravel())[0, 1] > 0.9


def test_tiles_segment_larger_than_shared_memory_capacity_is_parked_not_lost():
    """One particle per cell injected AT ONCE (stagger 0): 3,600 particles per tile against a shared-memory
    capacity of a few hundred records -- the excess waits in HBM, nothing is dropped, counters and mass close."""
    n, cycles = 240, 240 * 240
    hm = noise_map(n, seed=1, layers=8, workers=4)
    w = swe.World(n); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES); w.set_option(swe.OPT_STAGGER_ROUNDS, 0)
    w.upload(hm); w.srand(1)
    w.erode(cycles, swe.MODE_PARALLEL)
    st = w.stats()
    assert st.particles == cycles == st.exits_oob + st.exits_age and st.queue_dropped == 0
    drift = w.sum_plane(swe.HEIGHT) + st.sediment_lost - float(hm.astype(np.float64).sum())
    assert abs(drift) < 2e-6 * float(np.abs(hm.astype(np.float64)).sum())
    assert st.steps > 100 * cycles
    w.close()


@pytest.mark.parametrize("abrasion", [0.0, 0.3])
def test_tiles_layered_erosion_invariants_pyramid(abrasion):
    """The layered extension on the tile schedule (rock plane in shared memory beside the heights): the
    invariants of tests/test_gpu_layered.py, and the pick-up never digs below the bedrock (owned cells)."""
    n = 512
    sand = generate.fbm(n, seed=2) * np.float32(0.3)
    rock = generate.pyramid(n, apex=0.35)
    height = np.maximum(sand, rock).astype(np.float32)
    w = swe.World(n); w.set(abrasion=abrasion); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES)
    w.upload(height); w.upload_rock(rock); w.srand(3)
    h0, r0 = float(height.astype(np.float64).sum()), float(rock.astype(np.float64).sum())
    for _ in range(3):
        w.erode(n * n // 4, swe.MODE_PARALLEL)
    st = w.stats()
    assert st.particles == 3 * (n * n // 4) == st.exits_oob + st.exits_age
    h1, r1 = w.sum_plane(), float(w.download_rock().astype(np.float64).sum())
    assert abs(h1 + st.sediment_lost - h0) < 2e-5 * h0
    hg, rg = w.download(), w.download_rock()
    assert np.isfinite(hg).all() and np.isfinite(rg).all()
    assert (rg <= rock + 1e-7).all()
    assert float((rg - hg).max()) < 1e-6
    if abrasion == 0.0:
        assert np.array_equal(rg.view(np.uint32), rock.view(np.uint32))
    else:
        ground = r0 - r1
        q = swe.World(n); q.set(abrasion=abrasion); q.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_COLOUR_QUEUES)
        q.upload(height); q.upload_rock(rock); q.srand(3)
        for _ in range(3):
            q.erode(n * n // 4, swe.MODE_PARALLEL)
        ground_q = r0 - float(q.download_rock().astype(np.float64).sum())
        assert ground > 1e-4 and abs(ground / ground_q - 1.0) < 0.2, (ground, ground_q)
    w.close()


def test_tiles_options_and_fallbacks():
    w = swe.World(96)
    w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES)
    assert w.get_option(swe.OPT_SCHEDULE) == swe.SCHEDULE_TILES
    assert w.get_option(swe.OPT_TILE_ROUNDS) == 0 and w.get_option(swe.OPT_TILE_SHAPE) == -1     # auto
    w.set_option(swe.OPT_TILE_ROUNDS, 8)
    assert w.get_option(swe.OPT_TILE_ROUNDS) == 8
    for bad in (-1, 65):
        with pytest.raises(swe.SweError):
            w.set_option(swe.OPT_TILE_ROUNDS, bad)
    for bad in (-2, 4):
        with pytest.raises(swe.SweError):
            w.set_option(swe.OPT_TILE_SHAPE, bad)
    with pytest.raises(swe.SweError):
        w.set_option(swe.OPT_SCHEDULE, 3)
    w.upload(generate.flat_sand(96, seed=1)); w.srand(2)
    w.erode(96 * 96, swe.MODE_PARALLEL)
    st = w.stats()
    assert st.particles == 96 * 96 == st.exits_oob + st.exits_age and st.last_micro_rounds > 0
    # a torus has no tile schedule (tiles do not wrap): the call runs on the colour queues instead
    w.set_option(swe.OPT_PERIODIC, 1)
    w.erode(96 * 96, swe.MODE_PARALLEL)
    st2 = w.stats()
    assert st2.particles == 2 * 96 * 96 and st2.exits_oob == st.exits_oob       # (only age exits on a torus)
    assert st2.exits_age - st.exits_age == 96 * 96
    w.close()


def test_tiles_dune_wavelength_matches_sequential_reference():
    """cfg2 in small on the tile schedule: six calls from flat sand, dominant along-wind wavelength in the same or an
    adjacent spectral bin as the sequential oracle's."""
    n, cycles, calls = 256, 1 << 16, 6
    hm = generate.flat_sand(n, seed=42)
    ora = PortWorld(n); ora.set_plane(0, hm); ora.srand(1)
    w = swe.World(n); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES); w.upload(hm); w.srand(1)
    for _ in range(calls):
        ora.erode(cycles)
        w.erode(cycles, swe.MODE_PARALLEL)
    ho, hg = np.array(ora.plane(0)), w.download()
    ko, kg = n / analysis.dominant_wavelength(ho), n / analysis.dominant_wavelength(hg)
    assert abs(ko - kg) <= 1.0, (ko, kg)
    assert analysis.spectrum_similarity(ho, hg, 1, n // 4) > 0.99
    assert analysis.histogram_distance(ho, hg) < 0.12
    assert abs(float(ho.std()) / float(hg.std()) - 1.0) < 0.1

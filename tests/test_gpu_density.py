"""GPU: statistical parity of the parallel mode with the sequential reference AT THE BENCHMARKED PARTICLE
DENSITIES (round-1 verdict, item 1): the headline flies 0.25 particles per cell per World::erode call
(BASELINE cfg3), cfg2 flies 1.0 per cell. The sequential side is the oracle port (bit-identical to the
unmodified reference, tests/test_oracle_ref.py) flying the same number of particles one after another
(world.hpp:47-63, cascade after every step wind.hpp:114).

Same tolerances as tests/test_gpu_parallel.py:
    steps per particle +-3 %, exported sand +-5 %, height-histogram TV < 0.12, log-spectrum corr > 0.99,
    erosion-field corr > 0.95 (structured terrain), along-wind profile corr > 0.9.
Measured on B200 (colour-queue scheduler, automatic stagger), tools/colour_probe.py:
    1024^2 noise, 262,144 particles : steps +0.15 %, export +3.3 %, TV 0.009, spectrum 0.99987, field 0.969, profile 0.992
     512^2 noise,  65,536 particles : steps -0.34 %, export -0.3 %, TV 0.009, spectrum 0.99989, field 0.986, profile 0.996
The lock-step schedule of round 1 (SWE_OPT_SCHEDULE = 1) fails them at this density (+24 % exported sand,
field 0.80) -- asserted below so that the reason for the scheduler stays on record."""
import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import analysis, generate
from tests.oracle_lib import PortWorld, noise_map

pytestmark = pytest.mark.gpu


def compare(n, cycles, hm, calls=1, opts=None):
    ora = PortWorld(n)
    ora.set_plane(0, hm)
    ora.srand(1)
    so = sum(ora.erode(cycles).steps for _ in range(calls))
    w = swe.World(n)
    for k, v in {swe.OPT_SCHEDULE: swe.SCHEDULE_COLOUR_QUEUES, **(opts or {})}.items():   # (this file pins the colour queues; the
        w.set_option(k, v)                                                                 # default schedule is tests/test_gpu_tiles.py's)
    w.upload(hm)
    w.srand(1)
    for _ in range(calls):
        w.erode(cycles, swe.MODE_PARALLEL)
    st = w.stats()
    ho, hg = np.array(ora.plane(0)), w.download()
    lost_o = float(hm.astype(np.float64).sum() - ho.astype(np.float64).sum())
    drift = w.sum_plane(swe.HEIGHT) + st.sediment_lost - float(hm.astype(np.float64).sum())
    out = {
        "steps": st.steps / so - 1.0,
        "export": st.sediment_lost / lost_o - 1.0,
        "tv": analysis.histogram_distance(ho, hg),
        "spec": analysis.spectrum_similarity(ho, hg, 1, n // 4),
        "field": float(np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1]),
        "profile": float(np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1]),
        "massflow": float(w.download(swe.MASSFLOW).mean()) / float(np.array(ora.plane(1)).mean()) - 1.0,
        "drift": abs(drift) / float(np.abs(hm.astype(np.float64)).sum()),
        "stats": st,
    }
    assert st.particles == calls * cycles and st.exits_oob + st.exits_age == st.particles
    assert st.queue_dropped == 0
    assert np.isfinite(hg).all()
    w.close()
    return out


def test_headline_density_quarter_particle_per_cell():
    """1024^2 / 262,144 particles = 0.25 per cell: the density of bench.py's headline (4096^2 / 4,194,304)."""
    n, cycles = 1024, 1 << 18
    hm = noise_map(n, seed=1, layers=8, workers=8)
    r = compare(n, cycles, hm)
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r
    assert abs(r["massflow"]) < 0.1, r
    assert r["drift"] < 2e-6, r


def test_cfg2_density_one_particle_per_cell_flat_sand():
    """256^2 flat sand / 65,536 particles = 1.0 per cell (BASELINE cfg2's density; on flat sand the cell-level
    erosion pattern is turbulence noise, so the field correlation is not asserted -- two sequential runs with
    different seeds correlate 0.38)."""
    n, cycles = 256, 1 << 16
    hm = generate.flat_sand(n, seed=42)
    r = compare(n, cycles, hm)
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["profile"] > 0.9, r
    assert r["drift"] < 2e-6, r


def test_one_particle_per_cell_structured_terrain_two_calls():
    """1.0 particle per cell on the noise-layer dune field, two successive calls (errors compound)."""
    n, cycles = 256, 1 << 16
    hm = noise_map(n, seed=1, layers=8, workers=4)
    r = compare(n, cycles, hm, calls=2)
    assert abs(r["steps"]) < 0.03, r
    assert abs(r["export"]) < 0.05, r
    assert r["tv"] < 0.12 and r["spec"] > 0.99, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r


def test_lockstep_schedule_fails_these_tolerances_at_headline_density():
    """Round 1's schedule, kept as SWE_OPT_SCHEDULE = 1: faster per step, but out of tolerance here."""
    n, cycles = 512, 1 << 16
    hm = noise_map(n, seed=1, layers=8, workers=4)
    lock = compare(n, cycles, hm, opts={swe.OPT_SCHEDULE: swe.SCHEDULE_LOCKSTEP})
    col = compare(n, cycles, hm)
    assert abs(col["export"]) < 0.05 and col["field"] > 0.95 and col["profile"] > 0.9, col
    assert lock["export"] > 0.10 and lock["field"] < 0.95, lock


@pytest.mark.parametrize("persistent", [0, 1])
def test_launch_per_micro_round_form_matches_persistent_form(persistent):
    """SWE_OPT_PERSISTENT = 0 (CUDA graph of one kernel per micro-round + host completion check) is the same
    schedule as the cooperative kernel; both must sit inside the tolerances."""
    n, cycles = 256, 1 << 14
    hm = noise_map(n, seed=1, layers=8, workers=4)
    r = compare(n, cycles, hm, opts={swe.OPT_PERSISTENT: persistent})
    assert abs(r["steps"]) < 0.03 and abs(r["export"]) < 0.05, r
    assert r["field"] > 0.95 and r["profile"] > 0.9, r
    assert r["stats"].last_micro_rounds > 0


def test_stagger_option_and_sparse_calls():
    """The reference's own frame (8 * tilesize particles: 0.016 per cell) is injected at once (auto stagger 0)
    and needs ~9 micro-rounds per step; a forced stagger spreads the injection."""
    n = 512
    hm = noise_map(n, seed=1, layers=8, workers=4)
    w = swe.World(n); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_COLOUR_QUEUES); w.upload(hm); w.srand(1)
    w.erode(8 * n, swe.MODE_PARALLEL)
    a = w.stats().last_micro_rounds
    w.set_option(swe.OPT_STAGGER_ROUNDS, 600)
    w.erode(8 * n, swe.MODE_PARALLEL)
    b = w.stats().last_micro_rounds
    assert 0 < a < 9 * 514 + 200 and b > a + 9 * 400, (a, b)
    with pytest.raises(swe.SweError):
        w.set_option(swe.OPT_STAGGER_ROUNDS, -2)
    with pytest.raises(swe.SweError):
        w.set_option(swe.OPT_SCHEDULE, 3)
    st = w.stats()
    assert st.exits_oob + st.exits_age == st.particles == 16 * n
    w.close()


def test_cfg2_dune_wavelength_matches_sequential_reference():
    """Dune formation from flat sand at cfg2's density (1.0 particle per cell per call), six calls: the dominant
    along-wind wavelength (detrended spectrum, k >= 4: analysis.dominant_wavelength) of the parallel run must sit in
    the same or an adjacent spectral bin as the sequential oracle's, and the spectra must correlate."""
    n, cycles, calls = 256, 1 << 16, 6
    hm = generate.flat_sand(n, seed=42)
    ora = PortWorld(n); ora.set_plane(0, hm); ora.srand(1)
    w = swe.World(n); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_COLOUR_QUEUES); w.upload(hm); w.srand(1)
    for _ in range(calls):
        ora.erode(cycles)
        w.erode(cycles, swe.MODE_PARALLEL)
    ho, hg = np.array(ora.plane(0)), w.download()
    ko, kg = n / analysis.dominant_wavelength(ho), n / analysis.dominant_wavelength(hg)
    assert abs(ko - kg) <= 1.0, (ko, kg)
    assert analysis.spectrum_similarity(ho, hg, 1, n // 4) > 0.99
    assert analysis.histogram_distance(ho, hg) < 0.12
    assert abs(float(ho.std()) / float(hg.std()) - 1.0) < 0.1

"""GPU: layered-terrain extension (bedrock under the sand).

NOT in the reference snapshot (its cell has one height, source/cellpool.h:226-240), so parity
is UNPINNED against the reference: the checker for the cascade is the oracle's restatement of
OUR specification (oracle/swe_oracle.c, `rock`), and the erosion is checked through
invariants. With the rock far below the surface the layered engine must reduce to the pinned
single-layer results bit for bit."""
import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
from tests.oracle_lib import PortWorld

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["tile-tma", "percolour", "tile-plain"])
@pytest.mark.parametrize("nx,nz", [(64, 64), (130, 260), (300, 1000)])
def test_layered_cascade_sweep_matches_specification(nx, nz, variant):
    rs = np.random.RandomState(nx + nz)
    hm = rs.random_sample((nx, nz)).astype(np.float32)
    rock = (hm - rs.random_sample((nx, nz)).astype(np.float32) * 0.02).astype(np.float32)   # thin sand
    rock[: nx // 3] = -10.0                                   # deep sand: unconstrained there
    rock[nx // 2: nx // 2 + 10, : nz // 2] = hm[nx // 2: nx // 2 + 10, : nz // 2] + 0.1     # bare rock above surface
    ora = PortWorld(nx, nz); ora.set_plane(0, hm); ora.set_rock(rock)
    w = swe.World(nx, nz); w.set_option(swe.OPT_CASCADE_VARIANT, variant)
    w.upload(hm); w.upload_rock(rock)
    s0 = w.sum_plane()
    for _ in range(3):
        w.cascade_sweep(1); ora.cascade_sweep(1)
        assert np.array_equal(bits(w.download()), bits(ora.plane(0)))
    assert abs(w.sum_plane() - s0) < 1e-6 * abs(s0)           # sand only moves
    assert np.array_equal(bits(w.download_rock()), bits(rock))  # cascade never touches rock
    # cells whose sand ran out (clamped at the bedrock) kept a slope steeper than repose
    assert w.max_excess() > 0.0
    # single cell, layered
    w.cascade((nx // 2 + 3, 5)); ora.cascade(nx // 2 + 3, 5)
    assert np.array_equal(bits(w.download()), bits(ora.plane(0)))


def test_deep_rock_reduces_to_single_layer_bitwise():
    n = 256
    hm = np.random.RandomState(4).random_sample((n, n)).astype(np.float32)
    a = swe.World(n); a.upload(hm)
    b = swe.World(n); b.upload(hm); b.upload_rock(np.full((n, n), -1e30, np.float32))
    a.cascade_sweep(4); b.cascade_sweep(4)
    assert np.array_equal(bits(a.download()), bits(b.download()))
    b.clear_rock()
    a.cascade_sweep(1); b.cascade_sweep(1)
    assert np.array_equal(bits(a.download()), bits(b.download()))


@pytest.mark.parametrize("abrasion", [0.0, 0.3])
def test_layered_erosion_invariants_pyramid(abrasion):
    """Dune field over a rock pyramid (the 4096^2 configuration in small): sand mass + exported
    sediment is conserved up to the rock ground to sand, rock only shrinks and only when
    abrasion is on, the surface never ends up meaningfully below the bedrock."""
    n = 512
    sand = generate.fbm(n, seed=2) * np.float32(0.3)
    rock = generate.pyramid(n, apex=0.35)
    height = np.maximum(sand, rock).astype(np.float32)
    w = swe.World(n); w.set(abrasion=abrasion)
    w.upload(height); w.upload_rock(rock); w.srand(3)
    h0, r0 = float(height.astype(np.float64).sum()), float(rock.astype(np.float64).sum())
    for _ in range(3):
        w.erode(n * n // 4, swe.MODE_PARALLEL)
    st = w.stats()
    h1, r1 = w.sum_plane(), float(w.download_rock().astype(np.float64).sum())
    assert abs(h1 + st.sediment_lost - h0) < 2e-5 * h0        # total surface mass balance (rock included)
    hg, rg = w.download(), w.download_rock()
    assert np.isfinite(hg).all() and np.isfinite(rg).all()
    assert (rg <= rock + 1e-7).all()                           # rock never grows
    if abrasion == 0.0:
        assert np.array_equal(bits(rg), bits(rock))            # untouched, bit for bit
    else:
        assert r0 - r1 > 1e-4                                  # some rock was ground
        assert r0 - r1 < 0.5 * st.sediment_lost + 1.0
    assert float((rg - hg).max()) < 2e-3                       # concurrent takes may overshoot by one grain
    # the bare apex cannot be lowered when abrasion is off (sand may still settle on it)
    apex = (slice(n // 2 - 8, n // 2 + 8),) * 2
    if abrasion == 0.0:
        assert float((height[apex] - hg[apex]).max()) < 1e-5


def test_layered_mode_guards():
    w = swe.World(64)
    with pytest.raises(swe.SweError):
        w.download_rock()
    w.upload_rock(np.zeros((64, 64), np.float32))
    with pytest.raises(swe.SweError):
        w.erode(8, swe.MODE_DETERMINISTIC)
    with pytest.raises(swe.SweError):
        w.set(abrasion=1.5)
    w.clear_rock()
    w.erode(8, swe.MODE_DETERMINISTIC)


@pytest.mark.parametrize("ring", [True, False], ids=["colour-queues", "lock-step"])
def test_layered_erosion_on_stripes(ring):
    """The layered extension on multi-GPU stripes (logical stripes on one GPU): every rank holds the bedrock of its
    stored rows (owned + ghost), the invariants of the single-GPU test hold across the stripe boundaries, and the
    rock ground away matches a single engine's statistically."""
    from simplewinderosion_b200 import stripes
    n, world, cycles = 256, 2, 256 * 256 // 4
    sand = generate.fbm(n, seed=2) * np.float32(0.3)
    rock = generate.pyramid(n, apex=0.35)             # the pyramid straddles the stripe boundary at row n/2
    height = np.maximum(sand, rock).astype(np.float32)
    ws = [stripes.StripedWorld(n, n, r, world, cycles) for r in range(world)]
    stripes.StripedWorld.connect_in_process(ws, ring=ring)
    g = stripes.StripeGroup(ws)
    single = swe.World(n)
    single.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_COLOUR_QUEUES if ring else swe.SCHEDULE_LOCKSTEP)
    for w in ws + [single]:
        w.set(abrasion=0.3)
        w.srand(3)
    for w in ws:
        w.upload_global(height)
        w.upload_rock_global(rock)
    single.upload(height); single.upload_rock(rock)
    g.sync_halos()
    h0, r0 = float(height.astype(np.float64).sum()), float(rock.astype(np.float64).sum())
    if not ring:
        # the lock-step stripe round recomputes its ghost rows from bit-identical inputs on both sides; abrasion changes
        # the bedrock under a neighbour's ghost copy, so layered stripes are a colour-queue feature
        with pytest.raises(swe.SweError):
            ws[0].erode(cycles, swe.MODE_PARALLEL)
        return
    for _ in range(2):
        g.erode(cycles)
        single.erode(cycles, swe.MODE_PARALLEL)
    sts = g.stats()
    assert sum(s.particles for s in sts) == 2 * cycles == sum(s.exits_oob + s.exits_age for s in sts)
    assert sum(s.stripe_overflow + s.queue_dropped for s in sts) == 0
    hg = g.gather()
    rg = np.concatenate([w.download_rock() for w in ws], axis=0)
    assert hg.shape == rg.shape == (n, n) and np.isfinite(hg).all() and np.isfinite(rg).all()
    lost = sum(s.sediment_lost for s in sts)
    assert abs(float(hg.astype(np.float64).sum()) + lost - h0) < 2e-5 * h0
    assert (rg <= rock + 1e-7).all()
    assert float((rg - hg).max()) < (1e-6 if ring else 2e-3)     # colour queues own their cell: never below the bedrock
    ground, ground_single = r0 - float(rg.astype(np.float64).sum()), r0 - float(single.download_rock().astype(np.float64).sum())
    assert ground > 1e-4 and abs(ground / ground_single - 1.0) < 0.2, (ground, ground_single)


def test_colour_queue_pickup_never_digs_below_the_bedrock():
    """Dense layered erode (one particle per cell per call): with exclusive cell ownership the pick-up is capped by the
    sand exactly, so height >= rock holds everywhere (the lock-step schedule needs a compare-and-swap loop for that)."""
    n = 192
    rock = generate.pyramid(n, apex=0.3)
    height = (rock + np.float32(0.002)).astype(np.float32)      # a thin veneer of sand everywhere
    w = swe.World(n); w.set(abrasion=0.0)
    w.upload(height); w.upload_rock(rock); w.srand(1)
    for _ in range(2):
        w.erode(n * n, swe.MODE_PARALLEL)
    hg, rg = w.download(), w.download_rock()
    assert np.array_equal(bits(rg), bits(rock))
    assert float((hg - rg).min()) >= -1e-7
    v = swe.World(n); v.set(abrasion=0.0); v.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_LOCKSTEP)
    v.upload(height); v.upload_rock(rock); v.srand(1)
    v.erode(n * n, swe.MODE_PARALLEL)
    assert float((v.download() - rock).min()) >= -1e-6          # CAS pick-up (always on in layered lock-step mode)

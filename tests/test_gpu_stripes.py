"""GPU: multi-GPU row stripes exercised on ONE GPU as N logical stripes in one process (the
halo / hand-off logic is device-count independent; real 2-8 GPU runs go through
tests/mgpu_worker.py under torchrun, see test_real_multi_gpu).

* whole-map cascade sweeps on stripes are BITWISE equal to the single-engine sweep (and the
  oracle): each rank recomputes its rows from a 3/5-row halo covering the nine-colour
  dependency cone;
* parallel erode on stripes: every particle is accounted for exactly once across hand-offs,
  mass balance closes, and the result matches the single-engine run statistically (float
  atomics make neither run bitwise reproducible)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate, stripes
from tests.oracle_lib import PortWorld

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_stripes(nx, nz, world, max_particles, hm, ring=True):
    ws = [stripes.StripedWorld(nx, nz, r, world, max_particles) for r in range(world)]
    stripes.StripedWorld.connect_in_process(ws, ring=ring)
    for w in ws:
        w.upload_global(hm)
    g = stripes.StripeGroup(ws)
    g.sync_halos()
    return g


@pytest.mark.parametrize("nx,nz,world", [(64, 48, 2), (100, 260, 3), (256, 256, 4), (130, 64, 8)])
def test_striped_cascade_sweep_bitwise(nx, nz, world):
    rs = np.random.RandomState(nx + world)
    hm = rs.random_sample((nx, nz)).astype(np.float32)
    ora = PortWorld(nx, nz)
    ora.set_plane(0, hm)
    g = make_stripes(nx, nz, world, 1024, hm)
    for _ in range(3):
        g.cascade_sweep(1)
        ora.cascade_sweep(1)
        got = g.gather()
        assert np.array_equal(bits(got), bits(ora.plane(0))), np.argwhere(bits(got) != bits(ora.plane(0)))[:4]
    assert np.float32(g.max_excess()).view(np.uint32) == np.float32(ora.max_excess()).view(np.uint32)
    assert all(s.stripe_overflow == 0 for s in g.stats())


@pytest.mark.parametrize("n,world,cycles,ring", [(128, 2, 4096, True), (256, 4, 16384, True), (128, 2, 4096, False),
                                                 (256, 4, 16384, False)])
def test_striped_erode_accounts_for_every_particle(n, world, cycles, ring):
    """ring = True: colour-queue schedule on stripes (one persistent kernel per stripe, barrier across the stripes,
    boundary blocks read/written in the neighbour's memory); ring = False: round 1's lock-step schedule with halo
    exchange per round. Compared with a single engine running the same schedule."""
    hm = generate.gaussian_hill(n)
    g = make_stripes(n, n, world, cycles, hm, ring=ring)
    ws = g.worlds
    for w in ws:
        w.srand(7)
    sched = swe.SCHEDULE_COLOUR_QUEUES if ring else swe.SCHEDULE_LOCKSTEP
    single = swe.World(n)
    single.set_option(swe.OPT_SCHEDULE, sched)
    single.upload(hm)
    single.srand(7)
    other = swe.World(n)          # same engine, another seed: how much two runs differ anyway
    other.set_option(swe.OPT_SCHEDULE, sched)
    other.upload(hm)
    other.srand(8)
    s0 = float(hm.astype(np.float64).sum())
    for _ in range(2):
        g.erode(cycles)
        single.erode(cycles, swe.MODE_PARALLEL)
        other.erode(cycles, swe.MODE_PARALLEL)
    sts = [w.stats() for w in ws]
    assert sum(s.stripe_overflow for s in sts) == 0
    assert sum(s.particles for s in sts) == 2 * cycles
    assert sum(s.exits_oob + s.exits_age for s in sts) == 2 * cycles
    steps = sum(s.steps for s in sts)
    ss = single.stats()
    if ring:
        assert all(s.last_micro_rounds == sts[0].last_micro_rounds > 0 for s in sts)   # one global micro-round sequence
        assert sum(s.queue_dropped for s in sts) == 0
    assert abs(steps / ss.steps - 1.0) < 0.01
    lost = sum(s.sediment_lost for s in sts)
    total = sum(w.sum_plane() for w in ws)
    assert abs(total + lost - s0) < 2e-5 * abs(s0)        # fp32 atomics: ~1e7 rounded adds per call
    got, ref = g.gather(), single.download()
    assert np.isfinite(got).all()
    baseline = np.corrcoef((other.download() - hm).ravel(), (ref - hm).ravel())[0, 1]
    assert np.corrcoef((got - hm).ravel(), (ref - hm).ravel())[0, 1] > baseline - 0.03, baseline
    assert abs(lost / ss.sediment_lost - 1.0) < 0.06
    # flow fields exist on every stripe and look like the single-engine ones
    f = g.gather(swe.MOMENTUMX)
    assert abs(float(f.mean()) / float(single.download(swe.MOMENTUMX).mean()) - 1.0) < 0.05


@pytest.mark.parametrize("nx,nz,world,shape", [(512, 512, 2, -1), (768, 320, 3, -1), (1024, 256, 4, 0), (1024, 384, 4, 1), (512, 200, 2, 1),
                                               (1024, 256, 2, 2), (768, 512, 3, 3)])
def test_striped_erode_tile_schedule(nx, nz, world, shape):
    """The tile-resident schedule on stripes (stripes at least two tiles tall): global tiling, a tile belongs to the rank
    that owns its centre row and reads / writes a neighbour's rows in the neighbour's memory, survivors are scattered
    into the owner's segment tables, three cross-stripe barriers per super-round. Against a single engine on the same
    schedule: every particle accounted for, mass balance, same statistics."""
    from tests.oracle_lib import noise_map
    cycles = nx * nz // 4
    hm = (noise_map(max(nx, nz), seed=1, layers=8, workers=4)[:nx, :nz]).astype(np.float32).copy()
    g = make_stripes(nx, nz, world, cycles, hm)
    ws = g.worlds
    single, other = swe.World(nx, nz), swe.World(nx, nz)
    for w, seed in [(w, 7) for w in ws] + [(single, 7), (other, 8)]:
        w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES)
        w.set_option(swe.OPT_TILE_SHAPE, shape)
        w.srand(seed)
    single.upload(hm); other.upload(hm)
    if shape in (1, 2) and nx // world < (248 if shape == 1 else 488):
        with pytest.raises(swe.SweError):        # stripes of 256 rows are fine for 120-row tiles, shorter ones are refused loudly
            g.erode(cycles)
        return
    s0 = float(hm.astype(np.float64).sum())
    for _ in range(2):
        g.erode(cycles)
        single.erode(cycles, swe.MODE_PARALLEL)
        other.erode(cycles, swe.MODE_PARALLEL)
    sts = [w.stats() for w in ws]
    ss = single.stats()
    assert all(s.last_super_rounds == sts[0].last_super_rounds > 0 for s in sts)          # the tile path ran, in lock step
    assert sum(s.particles for s in sts) == 2 * cycles == sum(s.exits_oob + s.exits_age for s in sts)
    assert sum(s.stripe_overflow + s.queue_dropped for s in sts) == 0
    assert abs(sum(s.steps for s in sts) / ss.steps - 1.0) < 0.01
    lost = sum(s.sediment_lost for s in sts)
    total = sum(w.sum_plane() for w in ws)
    assert abs(total + lost - s0) < 2e-6 * abs(s0)
    assert abs(lost / ss.sediment_lost - 1.0) < 0.03
    got, ref = g.gather(), single.download()
    assert np.isfinite(got).all()
    baseline = np.corrcoef((other.download() - hm).ravel(), (ref - hm).ravel())[0, 1]
    assert np.corrcoef((got - hm).ravel(), (ref - hm).ravel())[0, 1] > baseline - 0.02, baseline
    # the rows either side of every stripe boundary look like the single engine's (no seam in the erosion field)
    for w in ws[:-1]:
        hi = w.bounds[w.rank][1]
        seam = slice(hi - 3, hi + 3)
        assert np.corrcoef((got[seam] - hm[seam]).ravel(), (ref[seam] - hm[seam]).ravel())[0, 1] > baseline - 0.1
    # flow fields: the track vectors of a neighbour's rows were added in the neighbour's memory
    f, fs = g.gather(swe.MOMENTUMX), single.download(swe.MOMENTUMX)
    assert abs(float(f.mean()) / float(fs.mean()) - 1.0) < 0.03
    for w in ws[:-1]:
        hi = w.bounds[w.rank][1]
        assert abs(float(f[hi - 2:hi + 2].mean()) / float(fs[hi - 2:hi + 2].mean()) - 1.0) < 0.15
    # whole-map sweeps still work on the refreshed ghost rows
    g.cascade_sweep(1)
    assert np.isfinite(g.gather()).all()


def test_stripe_engine_guards():
    ws = make_stripes(64, 64, 2, 256, generate.flat_sand(64)).worlds
    w = ws[0]
    with pytest.raises(swe.SweError):
        w.erode(16, swe.MODE_DETERMINISTIC)
    with pytest.raises(swe.SweError):
        w.erode(100000, swe.MODE_PARALLEL)        # > max_particles
    with pytest.raises(swe.SweError):
        w.cascade((1, 1))
    with pytest.raises(swe.SweError):
        w.map.normal((40, 3))                     # row owned by the other stripe
    assert w.nx == 32 and w.global_nx == 64 and w.download().shape == (32, 64)
    lone = stripes.StripedWorld(64, 64, 0, 2, 256)
    with pytest.raises(swe.SweError):
        lone.erode(16, swe.MODE_PARALLEL)         # not connected
    with pytest.raises(swe.SweError):
        lone.connect_links(None, None)            # rank 0 of 2 needs a neighbour above


def test_real_multi_gpu():
    """One process per GPU under torchrun (CUDA IPC + NVLink P2P), when the box has >= 2 GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29733",
                        os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_OK" in r.stdout

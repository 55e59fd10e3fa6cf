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

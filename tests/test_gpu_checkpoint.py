"""GPU: checkpoint / resume. A deterministic run stopped, saved, loaded into a
fresh engine and continued must be BITWISE identical to the uninterrupted run and to the oracle
(the libc-rand replica travels in the run state); a parallel run resumes with the same call
counter, so its particle spawn sequence continues (statistical check: float atomics)."""
import os

import numpy as np
import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import checkpoint, generate
from tests.oracle_lib import PortWorld

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_deterministic_resume_is_bitwise(tmp_path):
    n = 96
    ora = PortWorld(n)
    ora.init_gaussian()
    ora.srand(9)
    w = swe.World(n)
    w.upload(ora.plane(0))
    w.set(maxAge=300, lrate=0.2)
    ora.set(maxAge=300, lrate=0.2)
    w.srand(9)
    w.erode(200, swe.MODE_DETERMINISTIC); ora.erode(200)
    path = str(tmp_path / "run.swe")
    header = checkpoint.save(w, path, user={"frame": 1})
    assert header["params"]["max_age"] == 300 and header["user"] == {"frame": 1}
    w.close()

    w2 = checkpoint.load(path)
    assert w2.params.max_age == 300 and abs(w2.params.lrate - 0.2) < 1e-7
    for k in range(9):
        assert np.array_equal(bits(w2.download(k)), bits(ora.plane(k)))
    w2.erode(200, swe.MODE_DETERMINISTIC); ora.erode(200)      # continues the SAME rand() stream
    for k in range(9):
        assert np.array_equal(bits(w2.download(k)), bits(ora.plane(k))), k
    # planes are readable without a GPU
    _, planes = checkpoint.read_planes(path)
    assert planes["height"].shape == (n, n)


def test_state_blob_errors():
    import ctypes as C
    w = swe.World(16)
    buf = (C.c_ubyte * 256)()
    with pytest.raises(swe.SweError):
        swe._capi.check(swe.lib.swe_export_state(w._h, buf, 8))
    with pytest.raises(swe.SweError):
        swe._capi.check(swe.lib.swe_import_state(w._h, buf, 256))      # all zeros: wrong magic


def test_parallel_resume_continues_the_call_counter(tmp_path):
    n = 256
    hm = generate.gaussian_hill(n)

    def fresh():
        w = swe.World(n)
        w.upload(hm)
        w.srand(5)
        return w

    a = fresh()
    a.erode(20000); a.erode(20000)
    b = fresh()
    b.erode(20000)
    path = str(tmp_path / "p.swe")
    checkpoint.save(b, path)
    c = checkpoint.load(path)
    c.erode(20000)
    # same spawn sequence (call index 1 after the resume) => same particle statistics up to atomics order
    sa, sc = a.stats(), c.stats()
    assert sc.erode_calls == 2 == sa.erode_calls                # the counters continue across the resume
    assert sc.particles == sa.particles == 40000
    assert abs(sa.steps - sc.steps) / sc.steps < 0.03
    # ... and so does the mass balance: sum(height) + sediment carried away is still the initial mass
    s0 = float(hm.astype(np.float64).sum())
    assert abs(c.sum_plane() + sc.sediment_lost - s0) < 2e-5 * abs(s0)
    assert c.get_option(swe.OPT_SCHEDULE) == swe.SCHEDULE_TILES          # (the default)
    ha, hc = a.download(), c.download()
    assert np.corrcoef((ha - hm).ravel(), (hc - hm).ravel())[0, 1] > 0.9
    d = fresh()
    d.erode(20000); d.srand(5); d.erode(20000)          # a run that REPEATS call 0 instead differs more
    hd = d.download()
    assert np.abs(ha - hc).mean() < np.abs(ha - hd).mean()


def test_layered_checkpoint_keeps_rock(tmp_path):
    n = 64
    w = swe.World(n)
    w.upload(generate.gaussian_hill(n) + np.float32(0.1))
    rock = generate.pyramid(n, apex=0.05)
    w.upload_rock(rock)
    path = str(tmp_path / "l.swe")
    h = checkpoint.save(w, path)
    assert h["has_rock"] and h["planes"][-1] == "rock"
    w2 = checkpoint.load(path)
    assert np.array_equal(bits(w2.download_rock()), bits(rock))


def test_checkpoint_is_written_atomically_and_keeps_options(tmp_path):
    n = 64
    w = swe.World(n)
    w.upload(generate.gaussian_hill(n))
    w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_LOCKSTEP)
    w.set_option(swe.OPT_CALL_BATCHES, 3)
    w.set_option(swe.OPT_STAGGER_ROUNDS, 77)
    path = str(tmp_path / "o.swe")
    h = checkpoint.save(w, path)
    assert h["options"] == {"cascade_interval": 1, "call_batches": 3, "schedule": 1, "stagger_rounds": 77, "tile_rounds": 0,
                            "tile_shape": -1}
    assert os.listdir(str(tmp_path)) == ["o.swe"]               # no temporary file left behind
    w2 = checkpoint.load(path)
    assert (w2.get_option(swe.OPT_SCHEDULE), w2.get_option(swe.OPT_CALL_BATCHES), w2.get_option(swe.OPT_STAGGER_ROUNDS)) == (1, 3, 77)


def test_stripe_checkpoint_resume(tmp_path):
    """One file per stripe; after loading into connected stripe engines the ghost rows are refreshed and the run
    continues with exact accounting."""
    from simplewinderosion_b200 import stripes
    n, world, cycles = 128, 2, 4096
    hm = generate.gaussian_hill(n)

    def make():
        ws = [stripes.StripedWorld(n, n, r, world, cycles) for r in range(world)]
        stripes.StripedWorld.connect_in_process(ws)
        return stripes.StripeGroup(ws)

    g = make()
    for w in g.worlds:
        w.upload_global(hm); w.srand(3)
    g.sync_halos()
    g.erode(cycles)
    paths = [str(tmp_path / f"s{r}.swe") for r in range(world)]
    for w, p in zip(g.worlds, paths):
        checkpoint.save(w, p)
    g2 = make()
    for w, p in zip(g2.worlds, paths):
        checkpoint.load(p, into=w)
    g2.sync_halos()
    assert np.array_equal(bits(g2.gather()), bits(g.gather()))
    g2.erode(cycles)
    sts = g2.stats()
    assert sum(s.particles for s in sts) == 2 * cycles == sum(s.exits_oob + s.exits_age for s in sts)
    s0 = float(hm.astype(np.float64).sum())
    total = sum(w.sum_plane() for w in g2.worlds) + sum(s.sediment_lost for s in sts)
    assert abs(total - s0) < 2e-5 * abs(s0)

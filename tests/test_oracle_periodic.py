"""CPU: the periodic-boundary variant. The reference carries it as two comment blocks beside
the live closed boundary (cellpool.h:96-113 slice::oob/get, :447-465 map::oob/get); oracle/Makefile compiles the
reference with exactly those blocks swapped in (libswe_ref_p<N>.so). The port's `periodic` flag is pinned to it bit
for bit: erode runs (particles never leave the map: every exit is an age exit), single cascades across the seam,
normals on the seam, whole-map sweeps."""
import numpy as np
import pytest

from tests.oracle_lib import PortWorld, RefWorld, ref_available

need_ref = pytest.mark.skipif(not ref_available(96, periodic=True), reason="oracle/_ref periodic build missing (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@need_ref
@pytest.mark.parametrize("n,calls,cycles", [(48, 3, 384), (96, 2, 768)])
def test_periodic_erode_matches_compiled_reference(n, calls, cycles):
    r = RefWorld(n, seed=1, periodic=True)
    w = PortWorld(n, periodic=True)
    w.init_gaussian()
    w.srand(1)
    assert np.array_equal(bits(w.plane(0)), bits(r.plane(0)))
    tot = dict(steps=0, exits_oob=0, exits_age=0)
    for _ in range(calls):
        a, b = r.erode_counted(cycles), w.erode(cycles)
        assert (a["steps"], a["exits_oob"], a["exits_age"]) == (b.steps, b.exits_oob, b.exits_age)
        tot["steps"] += b.steps; tot["exits_oob"] += b.exits_oob; tot["exits_age"] += b.exits_age
    assert tot["exits_oob"] == 0 and tot["exits_age"] == calls * cycles      # nothing leaves a torus
    assert tot["steps"] == calls * cycles * 513                              # every particle lives maxAge + 1 steps
    for k in range(9):
        assert np.array_equal(bits(w.plane(k)), bits(r.plane(k))), k
    assert w.hash_height() == r.hash_height()


@need_ref
def test_periodic_cascade_normal_and_sweep_on_the_seam():
    n = 48
    hm = np.random.RandomState(4).random_sample((n, n)).astype(np.float32)
    r = RefWorld(n, periodic=True); r.set_plane(0, hm)
    w = PortWorld(n, periodic=True); w.set_plane(0, hm)
    for x, z in [(0, 0), (0, n - 1), (n - 1, 0), (n - 1, n - 1), (0, 17), (23, n - 1)]:
        assert np.array_equal(bits(w.normal(x, z)), bits(r.normal(x, z))), (x, z)
        r.cascade(x, z); w.cascade(x, z)
        assert np.array_equal(bits(w.plane(0)), bits(r.plane(0))), (x, z)
    for _ in range(2):
        r.cascade_sweep(1); w.cascade_sweep(1)
        assert np.array_equal(bits(w.plane(0)), bits(r.plane(0)))


def test_periodic_port_conserves_mass_exactly_enough():
    n = 48
    w = PortWorld(n, periodic=True)
    w.init_gaussian()
    w.srand(2)
    s0 = float(w.plane(0).astype(np.float64).sum())
    st = w.erode(512)
    assert st.exits_oob == 0
    assert abs(float(w.plane(0).astype(np.float64).sum()) + st.sediment_lost - s0) < 1e-3

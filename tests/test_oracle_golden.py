"""CPU: the plain-C oracle restatement (oracle/swe_oracle.c) against the committed golden
fixtures that tests/golden/make_golden.py generated from the UNMODIFIED reference
(oracle/_ref). Bit-exact everywhere: the reference's results are deterministic fp32."""
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import PortWorld, port

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP = json.load(open(os.path.join(G, "fingerprints.json")))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def mt19937_uniform_map(n, seed=42):
    rs = np.random.RandomState(seed)
    u = rs.randint(0, 2 ** 32, size=n * n, dtype=np.uint32)
    f = (u.astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)
    f[f >= 1.0] = np.nextafter(np.float32(1), np.float32(0))
    return f.reshape(n, n)


def test_rand_stream_matches_glibc():
    w = PortWorld(8)
    for seed, want in FP["rand"].items():
        w.srand(int(seed))
        assert [w.rand() for _ in range(len(want))] == want


@pytest.mark.parametrize("run", [r for r in FP["erode"] if r["n"] <= 512], ids=lambda r: f"n{r['n']}")
def test_erode_fingerprints(run):
    w = PortWorld(run["n"])
    w.init_gaussian()
    w.srand(run["seed"])
    assert abs(float(w.plane(0).astype(np.float64).sum()) - run["sum_initial"]) < 1e-9
    tot = dict(steps=0, exits_oob=0, exits_age=0, sediment_lost=0.0)
    for _ in range(run["calls"]):
        st = w.erode(run["cycles"])
        tot["steps"] += st.steps; tot["exits_oob"] += st.exits_oob
        tot["exits_age"] += st.exits_age; tot["sediment_lost"] += st.sediment_lost
    assert tot["steps"] == run["steps"]
    assert tot["exits_oob"] == run["exits_oob"] and tot["exits_age"] == run["exits_age"]
    assert f"{w.hash_height():016x}" == run["hash"]
    assert abs(tot["sediment_lost"] - run["sediment_lost"]) < 1e-9
    # mass balance: what left the map is what the dead particles carried (fp32 round-off)
    final = float(w.plane(0).astype(np.float64).sum())
    assert abs(final + tot["sediment_lost"] - run["sum_initial"]) < 1e-3


def test_survey_fingerprint_512():
    """srand(1), 512^2 Gaussian, 5 x erode(4096) -> 6,006,434 steps,
    hash 5791fd81902c0af6 (an independent probe of the reference)."""
    run = [r for r in FP["erode"] if r["n"] == 512][0]
    assert run["steps"] == 6006434 and run["hash"] == "5791fd81902c0af6"


def test_all_planes_small_run():
    g = np.load(os.path.join(G, "erode64.npz"))
    w = PortWorld(64)
    w.init_gaussian()
    assert np.array_equal(bits(w.plane(0)), bits(g["initial"]))
    w.srand(int(g["seed"]))
    for _ in range(int(g["calls"])):
        w.erode(int(g["cycles"]))
    for k in range(9):
        assert np.array_equal(bits(w.plane(k)), bits(g["planes"][k])), f"plane {k}"


def test_trajectories():
    g = np.load(os.path.join(G, "traj512.npz"))
    w = PortWorld(512)
    w.init_gaussian()
    for i in range(6):
        w.srand(100 + i)
        x, z = g[f"t{i}_start"]
        rec, n = w.trace(float(x), float(z), 600)
        assert n == int(g[f"t{i}_steps"]), i
        assert np.array_equal(bits(rec), bits(g[f"t{i}_rec"])), i
    assert np.array_equal(bits(w.plane(0)), bits(g["height_after"]))
    # SURVEY.md 8c first-particle values
    r0 = g["t0_rec"]
    assert int(g["t0_steps"]) >= 2


def test_normals_incl_borders():
    g = np.load(os.path.join(G, "normals32.npz"))
    w = PortWorld(32)
    w.set_plane(0, g["height"])
    for x in range(32):
        for z in range(32):
            assert np.array_equal(bits(w.normal(x, z)), bits(g["normals"][x, z])), (x, z)


def test_cascade_single_cells():
    g = np.load(os.path.join(G, "cascade_cells32.npz"))
    for before, (x, z), after in zip(g["before"], g["centre"], g["after"]):
        w = PortWorld(32)
        w.set_plane(0, before)
        w.cascade(int(x), int(z))
        assert np.array_equal(bits(w.plane(0)), bits(after)), (x, z)


@pytest.mark.parametrize("case", FP["cascade_sweep"], ids=lambda c: f"n{c['n']}")
def test_cascade_sweep_fingerprints(case):
    w = PortWorld(case["n"])
    w.set_plane(0, mt19937_uniform_map(case["n"], case["seed"]))
    w.cascade_sweep(case["sweeps"])
    assert f"{w.hash_height():016x}" == case["hash"]
    if case["n"] == 512:
        assert case["hash"] == "f10ceca1a3f34b9c"  # SURVEY.md 8c probe fingerprint


def test_sweep_map_fixture():
    g = np.load(os.path.join(G, "sweep64.npz"))
    w = PortWorld(64)
    w.set_plane(0, g["before"])
    w.cascade_sweep(int(g["sweeps"]))
    assert np.array_equal(bits(w.plane(0)), bits(g["after"]))


def test_expf_restatement_vs_libm_sample():
    """swo_expf == this host's libm expf on a dense sample (the exhaustive 2^32 sweep was
    run once when the restatement was written; see oracle/swe_oracle.c)."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    L = port()
    rs = np.random.RandomState(3)
    xs = np.concatenate([
        -rs.random_sample(20000).astype(np.float32) * 4,          # the physical range: x <= 0
        (rs.random_sample(5000).astype(np.float32) - 0.5) * 200,
        np.array([0.0, -0.0, 88.0, 89.0, -103.0, -104.0, -150.0, np.inf, -np.inf,
                  np.float32(32.56463623046875), np.float32(-63.0994605)], dtype=np.float32)])
    for x in xs:
        a, b = libm.expf(float(x)), L.swo_expf(float(x))
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), float(x)


def test_cascade_properties():
    """Sweeps conserve mass (fp32 round-off) and never increase the worst excess much;
    a relaxed map is a fixed point."""
    w = PortWorld(48, 40)   # non-square, not a multiple of 3
    rs = np.random.RandomState(9)
    w.set_plane(0, rs.random_sample((48, 40)).astype(np.float32))
    s0 = float(w.plane(0).astype(np.float64).sum())
    e0 = w.max_excess()
    for _ in range(400):
        w.cascade_sweep(1)
    assert abs(float(w.plane(0).astype(np.float64).sum()) - s0) < 1e-2
    assert w.max_excess() < 1e-4 < e0
    # all pairs within the repose gap: further sweeps move (almost) nothing
    before = w.plane(0).copy()
    w.cascade_sweep(1)
    assert np.abs(w.plane(0) - before).max() < 1e-4

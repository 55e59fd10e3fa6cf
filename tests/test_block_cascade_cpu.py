"""CPU: the register 3x3 cascade of the colour-queue flight kernel (csrc/swe_block.cuh, compiled here as
plain C++) against the oracle's World::cascade (world.hpp:79-153) -- bit for bit, including ties, signed
zeros, map borders/corners and the layered extension."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import PortWorld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OFF = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libblock_shim.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "helpers", "block_shim.cpp"), "-o", so])
    L = C.CDLL(so)
    L.block_cascade.restype = C.c_uint
    L.block_cascade.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint, C.c_float, C.c_float,
                                C.c_float, C.c_int, C.c_float, C.POINTER(C.c_float)]
    return L


def run_block(L, hm, x, z, maxdiff, settling, rock=None):
    n0, n1 = hm.shape
    h8 = (C.c_float * 8)()
    rk8 = (C.c_float * 8)()
    valid = 0
    for k, (dx, dz) in enumerate(OFF):
        if 0 <= x + dx < n0 and 0 <= z + dz < n1:
            valid |= 1 << k
            h8[k] = hm[x + dx, z + dz]
            if rock is not None:
                rk8[k] = rock[x + dx, z + dz]
    c = C.c_float(hm[x, z])
    ortho = np.float32(1.0) * np.float32(maxdiff)
    diag = np.float32(np.sqrt(np.float32(2.0))) * np.float32(maxdiff)
    ch = L.block_cascade(C.byref(c), h8, valid, float(ortho), float(diag), settling, int(rock is not None),
                         float(rock[x, z]) if rock is not None else 0.0, rk8)
    out = hm.copy()
    if ch & 256:
        out[x, z] = c.value
    for k, (dx, dz) in enumerate(OFF):
        if (ch >> k) & 1:
            out[x + dx, z + dz] = h8[k]
    return out


@pytest.mark.parametrize("kind", ["random", "ties", "steep", "zeros"])
def test_block_cascade_matches_oracle(shim, kind):
    rng = np.random.default_rng(5)
    n = 7
    for trial in range(300):
        if kind == "random":
            hm = rng.uniform(0, 0.05, (n, n)).astype(np.float32)
        elif kind == "ties":
            hm = (rng.integers(0, 4, (n, n)) * 0.01).astype(np.float32)
        elif kind == "steep":
            hm = rng.uniform(-1, 1, (n, n)).astype(np.float32)
        else:
            hm = (rng.integers(-1, 2, (n, n)) * 0.02).astype(np.float32)
            hm[hm == 0] = rng.choice(np.array([0.0, -0.0], dtype=np.float32), int((hm == 0).sum()))
        ora = PortWorld(n)
        ora.set_plane(0, hm)
        x, z = int(rng.integers(0, n)), int(rng.integers(0, n))
        if trial % 5 == 0:
            x, z = [(0, 0), (0, n - 1), (n - 1, 0), (n - 1, n - 1), (0, 3), (3, 0), (n - 1, 3), (3, n - 1)][trial // 5 % 8]
        ora.cascade(x, z)
        got = run_block(shim, hm, x, z, 0.005, 1.0)
        assert np.array_equal(got.view(np.uint32), np.array(ora.plane(0)).view(np.uint32)), (kind, trial, x, z)


def test_block_cascade_params_and_layered(shim):
    rng = np.random.default_rng(9)
    n = 6
    for trial in range(200):
        hm = rng.uniform(0, 0.1, (n, n)).astype(np.float32)
        rock = (hm - rng.uniform(0, 0.02, (n, n))).astype(np.float32)
        ora = PortWorld(n)
        ora.set(maxdiff=0.002, settling=0.7)
        ora.set_plane(0, hm)
        ora.set_rock(rock)
        x, z = int(rng.integers(0, n)), int(rng.integers(0, n))
        ora.cascade(x, z)
        got = run_block(shim, hm, x, z, 0.002, 0.7, rock)
        assert np.array_equal(got.view(np.uint32), np.array(ora.plane(0)).view(np.uint32)), (trial, x, z)

"""CPU: the SCHEDULE of the tile-resident scheduler (csrc/swe_tile.cu), emulated on the oracle's physics
(tools/par_emul.c mode D: three shifted tilings, interior-only service, parking, staggered injection at super-round
granularity), against the oracle flying the same particles one after another. This is the design check that was run
before the kernel existed; the kernel itself is checked on the GPU (tests/test_gpu_tiles.py). Small
case: 256^2 noise-layer map, 16,384 particles (0.25 per cell), a few seconds."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from simplewinderosion_b200 import analysis
from tests.oracle_lib import OWorld, OStats, ORand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libpar_emul.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(ROOT, "tools", "par_emul.c"),
                           "-o", so, "-lm"])
    L = C.CDLL(so)
    W = C.POINTER(OWorld)
    L.swo_create.restype = W; L.swo_create.argtypes = [C.c_int, C.c_int]
    L.swo_destroy.argtypes = [W]
    L.swo_srand.argtypes = [C.POINTER(ORand), C.c_uint]
    L.swo_erode.argtypes = [W, C.c_int, C.POINTER(OStats)]
    L.swo_noise_stack.argtypes = [W, C.c_int, C.c_int, C.c_int]
    L.swo_normalize.argtypes = [W]
    L.emul_erode_tiles.restype = C.c_longlong
    L.emul_erode_tiles.argtypes = [W, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    return L


def plane(w, k, n):
    return np.ctypeslib.as_array(w.contents.plane[k], shape=(n, n))


@pytest.mark.parametrize("tile,rounds", [(60, 16), (120, 24), (120030, 16)], ids=["60x60", "120x120", "120x30-along-the-wind"])
def test_emulated_tile_schedule_matches_the_sequential_reference(emul, tile, rounds):
    n, cycles, stagger = 256, 256 * 256 // 4, 1028
    seq = emul.swo_create(n, n)
    emul.swo_noise_stack(seq, 1, 8, n); emul.swo_normalize(seq)
    hm = plane(seq, 0, n).copy()
    emul.swo_srand(C.byref(seq.contents.rng), 1)
    st = OStats()
    emul.swo_erode(seq, cycles, C.byref(st))
    ho = plane(seq, 0, n).copy()
    par = emul.swo_create(n, n)
    plane(par, 0, n)[...] = hm
    emul.swo_srand(C.byref(par.contents.rng), 3)
    lost, nrounds = C.c_double(), C.c_longlong()
    steps = emul.emul_erode_tiles(par, cycles, stagger, tile, rounds, 0, C.byref(lost), C.byref(nrounds))
    hg = plane(par, 0, n).copy()
    # every particle died (the emulator's loop ends only then) and the sand is accounted for
    assert abs(float(hg.astype(np.float64).sum()) + lost.value - float(hm.astype(np.float64).sum())) < 2e-5 * float(hm.sum())
    assert abs(steps / st.steps - 1.0) < 0.03
    assert abs(lost.value / st.sediment_lost - 1.0) < 0.05
    assert analysis.histogram_distance(ho, hg) < 0.12 and analysis.spectrum_similarity(ho, hg, 1, n // 4) > 0.99
    assert np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1] > 0.95
    assert np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1] > 0.9
    assert stagger < nrounds.value < 4 * stagger           # the call ends soon after the injection does
    emul.swo_destroy(seq); emul.swo_destroy(par)

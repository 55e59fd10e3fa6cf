"""TOOL (CPU, test infrastructure): how the statistical parity of parallel-mode scheduling variants with
the sequential reference depends on the schedule, using tools/par_emul.c (oracle physics).
    python tools/par_emul_study.py n density map variants...
a variant is B,S,V,C,K = batches, stagger rounds, visibility, cascade mode, steps per round (lock-step family), or
    qS,order         colour queues (csrc/swe_queue.cu): stagger S in MICRO-rounds, colour order 0 / 1
    pB,skip,T,S      colour-phased rounds, optionally tile-local with stale halos (T > 0)
    tS,T,K,flags     tile-resident super-rounds (csrc/swe_tile.cu): stagger S in rounds, tile edge T, K rounds per
                     super-round, flags & 1 = per-round instead of per-super-round injection
    cS,T,A,flags     the same with block claims instead of colour phases (not built): S and A in attempts"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplewinderosion_b200 import analysis  # noqa: E402

so = "/tmp/libpar_emul.so"
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(ROOT, "tools", "par_emul.c"),
                       "-o", so, "-lm"])
from tests.oracle_lib import OWorld, OStats, ORand  # noqa: E402
L = C.CDLL(so)
W = C.POINTER(OWorld)
L.swo_create.restype = W; L.swo_create.argtypes = [C.c_int, C.c_int]
L.swo_destroy.argtypes = [W]
L.swo_srand.argtypes = [C.POINTER(ORand), C.c_uint]
L.swo_erode.argtypes = [W, C.c_int, C.POINTER(OStats)]
L.swo_noise_stack.argtypes = [W, C.c_int, C.c_int, C.c_int]
L.swo_normalize.argtypes = [W]
L.emul_erode.restype = C.c_longlong
L.emul_erode.argtypes = [W, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
L.emul_erode_queues.restype = C.c_longlong
L.emul_erode_queues.argtypes = [W, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
L.emul_erode_tiles.restype = C.c_longlong
L.emul_erode_tiles.argtypes = [W, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
L.emul_erode_claims.restype = C.c_longlong
L.emul_erode_claims.argtypes = [W, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
L.emul_erode_phased.restype = C.c_longlong
L.emul_erode_phased.argtypes = [W, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]


def plane(w, k, n):
    return np.ctypeslib.as_array(w.contents.plane[k], shape=(n, n))


def make(n, kind, hm=None):
    w = L.swo_create(n, n)
    if hm is not None:
        plane(w, 0, n)[...] = hm
    elif kind == "noise":
        L.swo_noise_stack(w, 1, 8, n); L.swo_normalize(w)
    else:
        rng = np.random.default_rng(42)
        plane(w, 0, n)[...] = (0.25 + 1e-3 * rng.uniform(-1, 1, (n, n))).astype(np.float32)
    return w


def metrics(hm, ho, hg, lost_o, lost_g, so_, sg, n):
    return (f"steps {sg / so_ - 1:+.3%} export {lost_g / lost_o - 1:+.3%} TV {analysis.histogram_distance(ho, hg):.4f} "
            f"spec {analysis.spectrum_similarity(ho, hg, 1, n // 4):.5f} field {np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1]:.4f} "
            f"profile {np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1]:.4f}")


def main():
    n, dens, kind, calls = int(sys.argv[1]), float(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    cycles = int(n * n * dens)
    w = make(n, kind)
    hm = plane(w, 0, n).copy()
    L.swo_srand(C.byref(w.contents.rng), 1)
    st = OStats()
    so_, lost_o = 0, 0.0
    t0 = time.time()
    for _ in range(calls):
        L.swo_erode(w, cycles, C.byref(st)); so_ += st.steps; lost_o += st.sediment_lost
    ho = plane(w, 0, n).copy()
    print(f"{n}^2 {kind} density {dens} x {calls} calls: sequential {so_} steps, export {lost_o:.3f}, {time.time() - t0:.1f} s", flush=True)
    # noise floor: a second sequential run with another seed
    w2 = make(n, kind, hm); L.swo_srand(C.byref(w2.contents.rng), 7)
    s2, l2 = 0, 0.0
    for _ in range(calls):
        L.swo_erode(w2, cycles, C.byref(st)); s2 += st.steps; l2 += st.sediment_lost
    print("  sequential, other seed:  ", metrics(hm, ho, plane(w2, 0, n), lost_o, l2, so_, s2, n), flush=True)
    for v in sys.argv[5:]:
        if v.startswith("q"):   # colour queues: qS,order
            B, S = (int(x) for x in v[1:].split(","))
            Cm, V, K = -2, 0, 1
        elif v.startswith("c"):   # tile-resident super-rounds with block claims: cS,T,A,flags (S in attempts)
            B, S, V, K = (int(x) for x in v[1:].split(","))
            Cm = -4
        elif v.startswith("t"):   # tile-resident super-rounds: tS,T,K,flags
            B, S, V, K = (int(x) for x in v[1:].split(","))   # stagger, T, K, flags
            Cm = -3
        elif v.startswith("p"):   # phased: pB,claim_skip,T
            B, S, V, K = (int(x) for x in v[1:].split(","))   # batches, claim_skip, T, stagger
            Cm = -1
        else:
            B, S, V, Cm, K = (int(x) for x in v.split(","))
        wv = make(n, kind, hm); L.swo_srand(C.byref(wv.contents.rng), int(os.environ.get("EMUL_SEED", "3")))
        sg, lg, rounds = 0, 0.0, 0
        lost, rr = C.c_double(), C.c_longlong()
        t0 = time.time()
        for _ in range(calls):
            if Cm == -2:
                sg += L.emul_erode_queues(wv, cycles, B, S, C.byref(lost), C.byref(rr))
            elif Cm == -4:
                sg += L.emul_erode_claims(wv, cycles, B, S, V, K, C.byref(lost), C.byref(rr))
            elif Cm == -3:
                sg += L.emul_erode_tiles(wv, cycles, B, S, V, K, C.byref(lost), C.byref(rr))
            elif Cm == -1:
                sg += L.emul_erode_phased(wv, cycles, B, S, V, K, C.byref(lost), C.byref(rr))
            else:
                sg += L.emul_erode(wv, cycles, B, S, V, Cm, K, C.byref(lost), C.byref(rr))
            lg += lost.value; rounds += rr.value
        print(f"  {v:14s} rounds/call {rounds // calls}: ",
              metrics(hm, ho, plane(wv, 0, n), lost_o, lg, so_, sg, n), f"({time.time() - t0:.1f} s)", flush=True)


main()

"""Generates the committed golden fixtures FROM THE REFERENCE ITSELF (oracle/_ref: the
unmodified reference headers compiled headless by oracle/Makefile). Run it in the build
container, where /root/reference is mounted:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own; these files are
what pins the oracle port (tests/test_oracle_golden.py) and, through it, the CUDA path.
Each RefWorld size lives in its own subprocess-free library; one map per size per process.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.oracle_lib import RefWorld  # noqa: E402


def mt19937_uniform_map(n: int, seed: int = 42) -> np.ndarray:
    """std::mt19937 g(seed); std::uniform_real_distribution<float>(0,1), x outer / z inner
    (libstdc++: one 32-bit draw per float, float(u)/2^32, values >= 1 -> nextafter(1,0))."""
    rs = np.random.RandomState(seed)
    u = rs.randint(0, 2 ** 32, size=n * n, dtype=np.uint32)
    f = (u.astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)
    f[f >= 1.0] = np.nextafter(np.float32(1), np.float32(0))
    return f.reshape(n, n)


def main() -> None:
    fp: dict = {"generator": "tests/golden/make_golden.py from oracle/_ref (unmodified reference)"}

    # ---- libc rand() stream (world.hpp:52,54; wind.hpp:83) ---------------------------------
    r64 = RefWorld(64)
    streams = {}
    for seed in (0, 1, 42, 2026):
        r64.srand(seed)
        streams[str(seed)] = [r64.rand() for _ in range(40)]
    fp["rand"] = streams

    # ---- erode fingerprints: Gaussian-hill init, srand(1) ----------------------------------
    runs = []
    for n, calls, cycles in ((64, 3, 512), (128, 2, 1024), (256, 2, 2048), (512, 5, 4096), (1024, 1, 2048)):
        r = RefWorld(n, seed=1)
        h0 = r.sum_height()
        tot = {"steps": 0, "exits_oob": 0, "exits_age": 0, "sediment_lost": 0.0}
        for _ in range(calls):
            st = r.erode_counted(cycles)
            for k in tot:
                tot[k] += st[k]
        runs.append({"n": n, "calls": calls, "cycles": cycles, "seed": 1, "sum_initial": h0,
                     "sum_final": r.sum_height(), "hash": f"{r.hash_height():016x}", **tot})
        print(runs[-1])
    fp["erode"] = runs

    # ---- full 9-plane dump of a small run (64^2, 2 x erode(256)) ---------------------------
    r = RefWorld(64, seed=7)
    init = r.plane(0)
    r.erode(256)
    r.erode(256)
    planes = np.stack([r.plane(k) for k in range(9)])
    np.savez_compressed(os.path.join(HERE, "erode64.npz"), initial=init, planes=planes,
                        seed=7, calls=2, cycles=256)

    # ---- trajectories -----------------------------------------------------------------------
    traj = {}
    r = RefWorld(512, seed=1)
    for i, (x, z) in enumerate(((454.915009, 359.777008), (10.25, 500.5), (128.0, 256.0),
                                (0.0, 0.0), (511.4, 511.4), (300.499, 17.5))):
        r.srand(100 + i)
        rec, nsteps = r.trace(x, z, 600)
        traj[f"t{i}_start"] = np.array([x, z], dtype=np.float32)
        traj[f"t{i}_rec"] = rec
        traj[f"t{i}_steps"] = nsteps
    traj["height_after"] = r.plane(0)
    np.savez_compressed(os.path.join(HERE, "traj512.npz"), **traj)

    # ---- normals incl. borders/corners (cellpool.h:200-223) ----------------------------------
    r = RefWorld(32, seed=1)
    rs = np.random.RandomState(5)
    hm = rs.random_sample((32, 32)).astype(np.float32)
    r.set_plane(0, hm)
    nrm = np.zeros((32, 32, 3), dtype=np.float32)
    for x in range(32):
        for z in range(32):
            nrm[x, z] = r.normal(x, z)
    np.savez_compressed(os.path.join(HERE, "normals32.npz"), height=hm, normals=nrm)

    # ---- single-cell cascade cases (world.hpp:79-153): interior, edges, corners, ties --------
    cases = []
    rs = np.random.RandomState(11)
    r = RefWorld(32, seed=1)
    for ci in range(24):
        hm = (rs.random_sample((32, 32)) * (0.05 if ci % 3 == 0 else 1.0)).astype(np.float32)
        if ci % 4 == 1:   # exact ties around the centre
            hm = np.round(hm * 8).astype(np.float32) / 8
        x, z = [(5, 5), (0, 0), (0, 31), (31, 0), (31, 31), (0, 9), (9, 0), (31, 12), (12, 31)][ci % 9]
        r.set_plane(0, hm)
        r.cascade(x, z)
        cases.append((hm, np.array([x, z]), r.plane(0)))
    np.savez_compressed(os.path.join(HERE, "cascade_cells32.npz"),
                        before=np.stack([c[0] for c in cases]),
                        centre=np.stack([c[1] for c in cases]),
                        after=np.stack([c[2] for c in cases]))

    # ---- whole-map cascade sweeps (9-colour order) --------------------------------------------
    sweeps = []
    for n, ns in ((64, 3), (128, 2), (512, 5)):
        r = RefWorld(n, seed=1)
        hm = mt19937_uniform_map(n, 42)
        r.set_plane(0, hm)
        r.cascade_sweep(ns)
        sweeps.append({"n": n, "sweeps": ns, "seed": 42, "hash": f"{r.hash_height():016x}",
                       "sum": r.sum_height()})
        if n == 64:
            np.savez_compressed(os.path.join(HERE, "sweep64.npz"), before=hm, after=r.plane(0), sweeps=ns)
        print(sweeps[-1])
    fp["cascade_sweep"] = sweeps

    with open(os.path.join(HERE, "fingerprints.json"), "w") as f:
        json.dump(fp, f, indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()

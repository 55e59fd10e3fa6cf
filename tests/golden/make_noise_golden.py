"""Generates tests/golden/noise.npz from the UNMODIFIED reference FastNoiseLite class
(oracle/_ref, built by `make -C oracle ref` where /root/reference is mounted):
  * pointwise GetNoise values (OpenSimplex2, FBm of 3 octaves, seed 1337) at seeded inputs
  * the layer stack of cellpool.h:392-409 on a zero map and on the Gaussian hill, with the
    normalisation of cellpool.h:416-428
Run from the repo root:  python tests/golden/make_noise_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.oracle_lib import RefWorld  # noqa: E402


def main():
    n = 64
    r = RefWorld(n)
    rs = np.random.RandomState(2026)
    pts = np.empty((4096, 4), dtype=np.float32)
    pts[:, 0] = 2.0 ** rs.randint(0, 8, size=len(pts))                 # layer frequency
    pts[:, 1:3] = rs.random_sample((len(pts), 2))                      # p = cell / tile in [0,1)
    pts[:, 3] = rs.randint(0, 10000, size=len(pts))                    # SEED % 10000
    far = rs.rand(len(pts)) < 0.25                                     # generic coordinates too
    pts[far, 1:] = rs.uniform(-40, 40, size=(int(far.sum()), 3))
    vals = np.array([r.get_noise(*map(float, p)) for p in pts], dtype=np.float32)

    out = {"points": pts, "values": vals}
    for seed in (1, 4711):
        r.set_plane(0, np.zeros((n, n), np.float32))
        r.noise_stack(seed, 8)
        out[f"stack_raw_seed{seed}"] = r.plane(0)
        r.normalize()
        out[f"stack_norm_seed{seed}"] = r.plane(0)
    r2 = RefWorld(n)                                                   # re-initialises the Gaussian hill
    hill = r2.plane(0)
    r2.noise_stack(1, 8)
    r2.normalize()
    out["hill"] = hill
    out["hill_plus_stack_norm_seed1"] = r2.plane(0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "noise.npz"), **out)
    print("wrote noise.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

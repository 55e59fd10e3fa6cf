"""Times the on-device map::init noise stack (swe_generate_noise + normalise) on cuda:0."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

for n in (1024, 4096, 16384):
    w = swe.World(n)
    w.generate_noise(1, 8); w.sync()
    w.upload(np.zeros((n, n), np.float32))
    t0 = time.perf_counter()
    w.generate_noise(1, 8); w.sync()
    t1 = time.perf_counter()
    w.normalize(); w.sync()
    t2 = time.perf_counter()
    print(f"n={n}: noise stack {1e3 * (t1 - t0):.2f} ms ({n * n * 48 / (t1 - t0) / 1e9:.1f} G lattice evals/s), "
          f"range+normalise {1e3 * (t2 - t1):.2f} ms, hash {w.hash_height():016x}")
    w.close()

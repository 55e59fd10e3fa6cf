"""Small invocations of the newer kernels for compute-sanitizer (memcheck / racecheck):
python tools/sanitize_target.py [sweep|erode|all]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 600
w = swe.World(n, 1000)
if what in ("all", "gen"):
    w.generate_noise(3, 4)
    w.normalize()
    v = w.vertices()
if what in ("all", "sweep"):
    w.upload(np.random.RandomState(1).random_sample((n, 1000)).astype(np.float32))
    w.cascade_sweep(3)                                   # dense path
    flat = (0.25 + 0.004 * np.random.RandomState(2).random_sample((n, 1000))).astype(np.float32)
    flat[100:110, 300:330] += np.float32(0.05)           # a few hot pairs: pair-bitmap path with candidates
    w.upload(flat)
    w.cascade_sweep(3)
if what in ("all", "erode"):
    w.upload((0.25 + 0.004 * np.random.RandomState(2).random_sample((n, 1000))).astype(np.float32))
    w.erode(20000, swe.MODE_PARALLEL)
print("ok", w.hash_height())

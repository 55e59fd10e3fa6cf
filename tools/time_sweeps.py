"""Times whole-map cascade sweeps on the cfg5 workload (8192^2 U(0,1)) as the relaxation proceeds."""
import sys

import numpy as np

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = swe.World(n)
w.upload(np.random.RandomState(42).random_sample((n, n)).astype(np.float32))
w.cascade_sweep(1)
tot = 0.0
for k in range(1, 240):
    ex = w.cascade_sweep(1, want_excess=True)
    ms = w.stats().last_sweep_ms
    tot += ms
    if k < 4 or k % 20 == 0 or k in (5, 10):
        print(f"sweep {k:3d}: {ms:.3f} ms  ({8.0 * n * n / ms / 1e6:.0f} GB/s)  max excess {ex:.3g}", flush=True)
    if ex < 1e-4:
        print(f"converged after {k} sweeps, mean {tot / k:.3f} ms, last {ms:.3f} ms")
        break
print("hash", f"{w.hash_height():016x}")
flat = (0.25 + 0.001 * np.random.RandomState(1).random_sample((n, n))).astype(np.float32)
w.upload(flat)
for k in range(4):
    w.cascade_sweep(1)
    print(f"flat map sweep {k}: {w.stats().last_sweep_ms:.3f} ms")

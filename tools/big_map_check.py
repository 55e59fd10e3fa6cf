"""One-off scale check: a map with more than 2^31 cells (64-bit indexing everywhere), generated on the
device, one parallel erode and two whole-map sweeps; checks mass accounting without downloading planes.
python tools/big_map_check.py [edge]   (48000^2 = 2.3e9 cells needs ~100 GB of HBM)"""
import sys
import time

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
t0 = time.time()
w = swe.World(n)
w.generate_noise(1, 8)
lo, hi = w.height_range()
w.normalize(lo, hi)
s0 = w.sum_plane()
lo2, hi2 = w.height_range()
print(f"n={n} cells={n * n:.3e} range before {lo:.4f}..{hi:.4f} after {lo2}..{hi2} sum {s0:.6e} ({time.time() - t0:.1f} s)", flush=True)
assert lo2 == 0.0 and hi2 == 0.5
w.srand(1)
w.erode(1 << 22, swe.MODE_PARALLEL)
st = w.stats()
s1 = w.sum_plane()
drift = abs(s1 + st.sediment_lost - s0)
print(f"erode: steps={st.steps} particles={st.particles} exits={st.exits_oob + st.exits_age} "
      f"mass drift {drift:.3e} (rel {drift / s0:.2e}) {st.last_erode_ms:.1f} ms", flush=True)
assert st.exits_oob + st.exits_age == st.particles and drift < 1e-5 * s0
e0 = w.max_excess()
e1 = w.cascade_sweep(2, want_excess=True)
s2 = w.sum_plane()
print(f"sweeps: max excess {e0:.4g} -> {e1:.4g}, {w.stats().last_sweep_ms / 2:.2f} ms/sweep, sum drift {abs(s2 - s1) / s1:.2e}", flush=True)
assert e1 <= e0 and abs(s2 - s1) < 1e-6 * s1
# last rows / last cells are addressable and sane
import numpy as np
assert np.isfinite(w.map.height((n - 1, n - 1))) and np.isfinite(w.map.normal((n - 1, n - 2))).all()
print("big map ok")

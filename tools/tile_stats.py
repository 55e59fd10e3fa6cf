"""Debug build only (-DSWE_TILE_STATS): how many centres reach the relax queue / the generic fallback."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = swe.World(n); w.upload(generate.random_slopes(n))
f = swe.lib.swe_debug_tile_stats; f.argtypes = [C.c_void_p, C.c_int]
out = (C.c_ulonglong * 4)()
f(out, 1)
for i in range(0, 200, 8):
    ex = w.cascade_sweep(8, want_excess=True)
    ms = w.stats().last_sweep_ms / 8
    f(out, 1)
    print(f"sweeps {i+8:4d}: {ms:.3f} ms/sweep  queued {out[0]/out[2]:.4f} of scanned, generic {out[1]/max(1,out[0]):.4f} of queued, max excess {ex:.2e}", flush=True)
    if ex < 1e-4: break

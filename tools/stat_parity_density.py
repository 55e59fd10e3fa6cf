"""Statistical parity at the HEADLINE particle density: one World::erode call with 0.25 particles per
cell in flight at once (parallel mode, GPU) against the same number of particles flown one after
another by the sequential oracle (the reference semantics), on map::init's noise-layer dune field.
The headline size itself (4096^2, 4.19 M particles) would take the oracle ~11 minutes per call, so
this runs 1024^2 / 262,144 particles (~40 s of oracle time per call); the per-cell concurrency -- what
the parallel mode changes -- is the same.  python tools/stat_parity_density.py [calls]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe  # noqa: E402
from simplewinderosion_b200 import analysis  # noqa: E402
from tests.oracle_lib import PortWorld  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, cycles = 1024, 1 << 18
w = swe.World(n)
w.set_option(swe.OPT_CALL_BATCHES, batches)
w.generate_noise(1, 8); w.normalize(); w.srand(1)
hm = w.download()
ora = PortWorld(n)
ora.set_plane(0, hm)
ora.srand(1)
print(f"map {n}^2 noise layers (seed 1), {cycles} particles per call = {cycles / n / n:.2f} per cell, "
      f"{batches} batch(es) per call", flush=True)
so = 0
for c in range(1, calls + 1):
    t0 = time.time()
    so += ora.erode(cycles).steps
    t_ora = time.time() - t0
    w.erode(cycles, swe.MODE_PARALLEL)
    st = w.stats()
    ho, hg = np.array(ora.plane(0)), w.download()
    lost_o = float(hm.astype(np.float64).sum() - ho.astype(np.float64).sum())
    print(f"call {c}: steps/particle oracle {so / (c * cycles):.2f} gpu {st.steps / (c * cycles):.2f} | exported mass oracle {lost_o:.3f} "
          f"gpu {st.sediment_lost:.3f} | hist TV {analysis.histogram_distance(ho, hg):.4f} | log-spectrum corr "
          f"{analysis.spectrum_similarity(ho, hg, 1, n // 4):.5f} | erosion-field corr {np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1]:.4f} | "
          f"along-wind profile corr {np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1]:.4f} | "
          f"massflow mean ratio {float(w.download(swe.MASSFLOW).mean()) / float(np.array(ora.plane(1)).mean()):.4f} | "
          f"oracle {t_ora:.1f} s, gpu {st.last_erode_ms:.1f} ms", flush=True)

"""GPU probe of the colour-queue scheduler: mass balance, counters, statistical parity with the sequential
oracle at a given density, and timing.  python tools/colour_probe.py n density [stagger] [schedule] [map]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe  # noqa: E402
from simplewinderosion_b200 import analysis, generate  # noqa: E402
from tests.oracle_lib import PortWorld  # noqa: E402

n, dens = int(sys.argv[1]), float(sys.argv[2])
stagger = int(sys.argv[3]) if len(sys.argv) > 3 else -1
sched = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kind = sys.argv[5] if len(sys.argv) > 5 else "noise"
calls = int(sys.argv[6]) if len(sys.argv) > 6 else 1
oracle = os.environ.get("NO_ORACLE") is None
cycles = int(n * n * dens)
w = swe.World(n)
w.set_option(swe.OPT_SCHEDULE, sched)
w.set_option(swe.OPT_STAGGER_ROUNDS, stagger)
w.set_option(swe.OPT_PERSISTENT, int(os.environ.get("PERSISTENT", "1")))
if "TILE_SHAPE" in os.environ:
    w.set_option(swe.OPT_TILE_SHAPE, int(os.environ["TILE_SHAPE"]))
if "TILE_ROUNDS" in os.environ:
    w.set_option(swe.OPT_TILE_ROUNDS, int(os.environ["TILE_ROUNDS"]))
if kind == "noise":
    w.generate_noise(1, 8); w.normalize()
else:
    w.upload(generate.flat_sand(n, seed=42))
w.srand(1)
hm = w.download()
s0 = w.sum_plane(swe.HEIGHT)
print(f"map {n}^2 {kind}, {cycles} particles/call ({dens}/cell), schedule {sched}, stagger {stagger}", flush=True)
if oracle:
    ora = PortWorld(n); ora.set_plane(0, hm); ora.srand(1)
so = 0
for c in range(1, calls + 1):
    t0 = time.time()
    w.erode(cycles, swe.MODE_PARALLEL)
    st = w.stats()
    t_gpu = time.time() - t0
    hg = w.download()
    drift = w.sum_plane(swe.HEIGHT) + st.sediment_lost - s0
    line = (f"call {c}: particles {st.particles} steps {st.steps} ({st.steps / (c * cycles):.2f}/particle) oob {st.exits_oob} age {st.exits_age} "
            f"lost {st.sediment_lost:.3f} drift {drift:.3e} micro-rounds {st.last_micro_rounds} dropped {st.queue_dropped} "
            f"erode {st.last_erode_ms:.2f} ms (wall {t_gpu * 1e3:.1f}) finite {bool(np.isfinite(hg).all())}")
    if oracle:
        t0 = time.time(); so += ora.erode(cycles).steps; t_o = time.time() - t0
        ho = np.array(ora.plane(0))
        lost_o = float(hm.astype(np.float64).sum() - ho.astype(np.float64).sum())
        line += (f" | vs oracle: steps {st.steps / so - 1:+.3%} export {st.sediment_lost / lost_o - 1:+.3%} TV {analysis.histogram_distance(ho, hg):.4f} "
                 f"spec {analysis.spectrum_similarity(ho, hg, 1, n // 4):.5f} field {np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1]:.4f} "
                 f"profile {np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1]:.4f} "
                 f"massflow {float(w.download(swe.MASSFLOW).mean()) / float(np.array(ora.plane(1)).mean()):.4f} (oracle {t_o:.1f} s)")
    print(line, flush=True)

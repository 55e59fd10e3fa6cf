"""GPU probe: the tile-resident scheduler on logical stripes (one process, one GPU) against a single engine.
    python tools/stripe_tile_probe.py nx nz world density [calls]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe  # noqa: E402
from simplewinderosion_b200 import stripes  # noqa: E402
from tests.oracle_lib import noise_map  # noqa: E402

nx, nz, world, dens = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 2
cycles = int(nx * nz * dens)
single = swe.World(nx, nz)
single.generate_noise(1, 8); single.normalize(); single.srand(7)
hm = single.download()
ws = [stripes.StripedWorld(nx, nz, r, world, cycles) for r in range(world)]
stripes.StripedWorld.connect_in_process(ws)
for w in ws:
    w.upload_global(hm); w.srand(7)
g = stripes.StripeGroup(ws)
g.sync_halos()
s0 = float(hm.astype(np.float64).sum())
for c in range(1, calls + 1):
    g.erode(cycles)
    single.erode(cycles, swe.MODE_PARALLEL)
    sts = g.stats()
    ss = single.stats()
    lost = sum(s.sediment_lost for s in sts)
    total = sum(w.sum_plane() for w in ws)
    got, ref = g.gather(), single.download()
    print(f"call {c}: stripes particles {sum(s.particles for s in sts)} exits {sum(s.exits_oob + s.exits_age for s in sts)} steps {sum(s.steps for s in sts)} "
          f"super-rounds {[s.last_super_rounds for s in sts]} dropped {sum(s.queue_dropped + s.stripe_overflow for s in sts)} lost {lost:.3f} "
          f"drift {(total + lost - s0) / s0:.2e} erode ms {[round(s.last_erode_ms, 2) for s in sts]} | single steps {ss.steps} lost {ss.sediment_lost:.3f} "
          f"super-rounds {ss.last_super_rounds} erode {ss.last_erode_ms:.2f} ms | field corr {np.corrcoef((got - hm).ravel(), (ref - hm).ravel())[0, 1]:.4f} "
          f"finite {bool(np.isfinite(got).all())} per-stripe particles {[s.particles for s in sts]}", flush=True)

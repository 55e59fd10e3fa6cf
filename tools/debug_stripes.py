import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate, stripes

def make(n, world, cycles, hm):
    ws = [stripes.StripedWorld(n, n, r, world, cycles) for r in range(world)]
    stripes.StripedWorld.connect_in_process(ws)
    for w in ws: w.upload_global(hm)
    g = stripes.StripeGroup(ws); g.sync_halos(); return g

for (n, world, cycles, with_single) in [(128, 2, 4096, False), (128, 2, 4096, True), (256, 4, 16384, True)]:
    hm = generate.gaussian_hill(n)
    g = make(n, world, cycles, hm)
    for w in g.worlds: w.srand(7)
    single = swe.World(n); single.upload(hm); single.srand(7)
    s0 = float(hm.astype(np.float64).sum())
    for it in range(4):
        t = time.time(); g.erode(cycles); t1 = time.time() - t
        if with_single: single.erode(cycles, swe.MODE_PARALLEL)
        t = time.time()
        try:
            sts = g.stats(); err = ""
        except Exception as e:
            err = str(e); sts = []
        t2 = time.time() - t
        if sts:
            lost = sum(s.sediment_lost for s in sts); total = sum(w.sum_plane() for w in g.worlds)
            print(n, world, with_single, it, f"enqueue {t1:.3f}s wait {t2:.3f}s particles {sum(s.particles for s in sts)} exits {sum(s.exits_oob+s.exits_age for s in sts)} "
                  f"overflow {sum(s.stripe_overflow for s in sts)} drift {total+lost-s0:.3e} erode_ms {[round(s.last_erode_ms,1) for s in sts]}", flush=True)
        else:
            print(n, world, with_single, it, f"enqueue {t1:.3f}s wait {t2:.3f}s ERROR {err}", flush=True)
            break

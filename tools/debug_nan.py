import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
for n, parts, hm in ((128, 4096, generate.gaussian_hill(128)), (1024, 1 << 20, generate.flat_sand(1024))):
    w = swe.World(n); w.upload(hm); w.srand(1)
    s0 = w.sum_plane()
    for i in range(4):
        w.erode(parts, swe.MODE_PARALLEL)
        h = w.download()
        st = w.stats()
        print(n, i, "nan", int(np.isnan(h).sum()), "min", float(h.min()), "max", float(h.max()), "lost", st.sediment_lost,
              "steps", st.steps, "mass drift", w.sum_plane() + st.sediment_lost - s0, "ms", st.last_erode_ms,
              "max excess", w.max_excess(), flush=True)

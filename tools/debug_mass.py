import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
for n, cycles in ((128, 4096), (512, 65536)):
    hm = generate.gaussian_hill(n)
    for rep in range(3):
        w = swe.World(n); w.upload(hm); w.srand(11 + rep)
        s0 = w.sum_plane()
        out = []
        for _ in range(3):
            w.erode(cycles, swe.MODE_PARALLEL); st = w.stats()
            out.append(w.sum_plane() + st.sediment_lost - s0)
        print(os.environ.get("SWE_OLD_STAMPED", "0"), n, rep, ["%.2e" % d for d in out], flush=True)

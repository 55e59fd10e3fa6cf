"""GPU probe: parallel mode on a torus (SWE_OPT_PERIODIC) against the periodic oracle port, for several staggers.
    python tools/periodic_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import analysis, generate
from tests.oracle_lib import PortWorld
n, cycles = 384, 36864
for name, hm in (("hill+fbm", generate.gaussian_hill(n) + generate.fbm(n, seed=3) * np.float32(0.2)), ("hill", generate.gaussian_hill(n))):
    ora = PortWorld(n, periodic=True); ora.set_plane(0, hm); ora.srand(1); so = ora.erode(cycles)
    ora2 = PortWorld(n, periodic=True); ora2.set_plane(0, hm); ora2.srand(5); so2 = ora2.erode(cycles)
    ho = np.array(ora.plane(0))
    print(name, "oracle other seed: lost", so2.sediment_lost / so.sediment_lost - 1, "field", np.corrcoef((ho - hm).ravel(), (np.array(ora2.plane(0)) - hm).ravel())[0, 1])
    for stag in (-1, 0, 2056):
        w = swe.World(n); w.set_option(swe.OPT_PERIODIC, 1); w.set_option(swe.OPT_STAGGER_ROUNDS, stag); w.upload(hm); w.srand(1)
        w.erode(cycles, swe.MODE_PARALLEL); st = w.stats(); hg = w.download()
        print(name, "stagger", stag, "lost", st.sediment_lost / so.sediment_lost - 1, "massflow", float(w.download(swe.MASSFLOW).mean()) / float(np.array(ora.plane(1)).mean()) - 1,
              "TV", analysis.histogram_distance(ho, hg), "field", np.corrcoef((ho - hm).ravel(), (hg - hm).ravel())[0, 1],
              "profile", np.corrcoef((ho - hm).mean(axis=1), (hg - hm).mean(axis=1))[0, 1], "micro", st.last_micro_rounds, flush=True)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
n = 512
sand = generate.fbm(n, seed=2) * np.float32(0.3)
rock = generate.pyramid(n, apex=0.35)
height = np.maximum(sand, rock).astype(np.float32)
w = swe.World(n); w.set(abrasion=0.0)
w.upload(height); w.upload_rock(rock); w.srand(3)
for it in range(3):
    w.erode(n * n // 4, swe.MODE_PARALLEL)
    hg, rg = w.download(), w.download_rock()
    d = rg - hg
    i = np.unravel_index(np.argmax(d), d.shape)
    print(it, "max rock-height", d.max(), "at", i, "rock", rg[i], "h now", hg[i], "h0", height[i], "sand0", sand[i], "cells below", int((d > 1e-3).sum()))
    x, z = i
    print("   neighbourhood h now:\n", hg[x-2:x+3, z-2:z+3], "\n   rock:\n", rg[x-2:x+3, z-2:z+3], "\n   h0:\n", height[x-2:x+3, z-2:z+3])

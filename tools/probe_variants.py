import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
w = swe.World(4096); w.upload(generate.fbm(4096, seed=1)); w.srand(1); w.set_option(swe.OPT_KERNEL_TIMING, 1)
for i in range(3):
    w.erode(1 << 22, swe.MODE_PARALLEL); st = w.stats()
print(os.environ.get("SWE_B200_LIB", "default").split("/")[-1], f"erode {st.last_erode_ms:.1f} ms round {st.last_round_ms:.1f} cascade {st.last_cascade_ms:.1f}")
''' % ROOT
for lib in sorted(os.listdir(os.path.join(ROOT, "build", "variants"))):
    if lib.endswith(".so"):
        env = dict(os.environ, SWE_B200_LIB=os.path.join(ROOT, "build", "variants", lib))
        subprocess.run([sys.executable, "-c", code], env=env)

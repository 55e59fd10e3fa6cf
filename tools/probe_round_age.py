import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate
hm = generate.fbm(4096, seed=1)
for max_age in (16, 64, 128, 256, 512):
    w = swe.World(4096); w.upload(hm); w.srand(1); w.set(maxAge=max_age); w.set_option(swe.OPT_KERNEL_TIMING, 1)
    for i in range(2):
        w.erode(1 << 22, swe.MODE_PARALLEL); st = w.stats()
    r = max_age + 2
    print(f"maxAge {max_age}: per round: round kernel {1e3*st.last_round_ms/r:.1f} us, cascade {1e3*st.last_cascade_ms/r:.1f} us", flush=True)
    w.close()

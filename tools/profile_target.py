"""Small, fixed invocations for ncu captures: python tools/profile_target.py <what>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate

what = sys.argv[1] if len(sys.argv) > 1 else "fly4096"
if what == "fly4096":
    w = swe.World(4096); w.generate_noise(1, 8); w.normalize(); w.srand(1)   # bench.py's dune field
    for _ in range(2): w.erode(1 << 22, swe.MODE_PARALLEL)
    w.sync(); print(w.stats().last_fly_ms)
elif what == "fly4096tile1":
    w = swe.World(4096); w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES)
    if "TILE_ROUNDS" in os.environ: w.set_option(swe.OPT_TILE_ROUNDS, int(os.environ["TILE_ROUNDS"]))
    w.generate_noise(1, 8); w.normalize(); w.srand(1)
    w.erode(1 << 22, swe.MODE_PARALLEL)
    w.sync(); print(w.stats().last_fly_ms)
elif what == "fly1024flat":
    w = swe.World(1024); w.upload(generate.flat_sand(1024)); w.srand(1)
    for _ in range(2): w.erode(1 << 20, swe.MODE_PARALLEL)
    w.sync(); print(w.stats().last_fly_ms)
elif what == "cascade8192":
    w = swe.World(8192); w.upload(generate.random_slopes(8192))
    w.cascade_sweep(3); w.sync(); print(w.stats().last_sweep_ms)
elif what == "cascade8192relaxed":
    w = swe.World(8192); w.upload(generate.flat_sand(8192))
    w.cascade_sweep(3); w.sync(); print(w.stats().last_sweep_ms)
elif what == "cascade4096conv":
    w = swe.World(4096); w.upload(generate.random_slopes(4096))
    w.cascade_sweep(164); w.sync(); print(w.stats().last_sweep_ms / 164)
elif what == "cascade8192long":
    w = swe.World(8192); w.upload(generate.random_slopes(8192))
    w.cascade_sweep(44); w.sync(); print(w.stats().last_sweep_ms / 44)

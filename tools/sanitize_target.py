"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
python tools/sanitize_target.py [gen|sweep|erode|stripes|all]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402
from simplewinderosion_b200 import stripes  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 600
w = swe.World(n, 1000)
w.set(max_age=48)                                        # short lives: a few hundred micro-rounds under the tool
if what in ("all", "gen"):
    w.generate_noise(3, 4)
    w.normalize()
    v = w.vertices()
if what in ("all", "sweep"):
    w.upload(np.random.RandomState(1).random_sample((n, 1000)).astype(np.float32))
    w.cascade_sweep(3)                                   # dense-scan, then dense-direct path
    flat = (0.25 + 0.004 * np.random.RandomState(2).random_sample((n, 1000))).astype(np.float32)
    flat[100:110, 300:330] += np.float32(0.05)           # a few hot pairs: pair-bitmap path with candidates
    w.upload(flat)
    w.cascade_sweep(3)
if what in ("all", "erode"):
    w.upload((0.25 + 0.004 * np.random.RandomState(2).random_sample((n, 1000))).astype(np.float32))
    w.erode(20000, swe.MODE_PARALLEL)                    # colour queues, persistent cooperative kernel
    w.set_option(swe.OPT_STAGGER_ROUNDS, 20)
    w.erode(20000, swe.MODE_PARALLEL)                    # ... with staggered injection
    w.set_option(swe.OPT_PERSISTENT, 0)
    w.erode(3000, swe.MODE_PARALLEL)                     # ... one launch per micro-round
    w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_LOCKSTEP)
    w.erode(20000, swe.MODE_PARALLEL)                    # lock-step rounds + stamped tile cascade
    w.erode(64, swe.MODE_DETERMINISTIC)
    st = w.stats()
    assert st.particles == st.exits_oob + st.exits_age, "particle accounting"
if what in ("all", "stripes"):
    hm = (0.25 + 0.004 * np.random.RandomState(3).random_sample((192, 256))).astype(np.float32)
    ws = [stripes.StripedWorld(192, 256, r, 2, 4096) for r in range(2)]
    for s in ws:
        s.set(max_age=48)
    stripes.StripedWorld.connect_in_process(ws)
    g = stripes.StripeGroup(ws)
    for s in ws:
        s.upload_global(hm)
    g.sync_halos()
    g.erode(4096)                                        # colour queues across two logical stripes (cross-stripe barrier)
    g.cascade_sweep(2)
    sts = g.stats()
    assert sum(s.particles for s in sts) == 4096 == sum(s.exits_oob + s.exits_age for s in sts)
print("ok", w.hash_height())

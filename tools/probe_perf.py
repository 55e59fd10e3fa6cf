"""Exploratory GPU probe (not the bench): quick throughput numbers for each kernel."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate

def par(n, particles, calls=3, hm=None):
    w = swe.World(n)
    w.upload(generate.fbm(n, seed=1) if hm is None else hm)
    w.srand(1)
    for i in range(calls):
        s0 = w.stats().steps
        w.erode(particles, swe.MODE_PARALLEL)
        st = w.stats()
        print(f"par n={n} particles={particles} call={i}: erode {st.last_erode_ms:.2f} ms fly {st.last_fly_ms:.2f} ms "
              f"steps={st.steps - s0} -> {(st.steps - s0) / st.last_fly_ms / 1e6:.2f} Gsteps/s(fly) "
              f"oob={st.exits_oob} age={st.exits_age}", flush=True)
    w.close()

def det(n, cycles):
    w = swe.World(n)
    w.upload(generate.gaussian_hill(n)); w.srand(1)
    w.erode(cycles, swe.MODE_DETERMINISTIC)
    st = w.stats()
    print(f"det n={n} cycles={cycles}: {st.last_erode_ms:.1f} ms steps={st.steps} -> {st.steps / st.last_erode_ms / 1e3:.3f} Msteps/s", flush=True)
    w.close()

def casc(n, sweeps=10):
    w = swe.World(n)
    w.upload(generate.random_slopes(n))
    for i in range(sweeps):
        ex = w.cascade_sweep(1, want_excess=True)
        st = w.stats()
        gbs = n * n * 8 / st.last_sweep_ms / 1e6
        print(f"cascade n={n} sweep {i}: {st.last_sweep_ms:.3f} ms -> {gbs:.1f} GB/s algorithmic, max excess {ex:.5f}", flush=True)
    w.close()

if __name__ == "__main__":
    det(512, 4096)
    par(512, 4096)
    par(1024, 1 << 20, hm=generate.flat_sand(1024))
    par(4096, 1 << 22)
    casc(2048, 4)
    casc(8192, 6)

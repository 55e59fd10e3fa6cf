"""Exploration: parallel mode (GPU) vs the sequential oracle on the same configuration."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate, analysis
from tests.oracle_lib import PortWorld

def run(n, frames, cycles, hm, label, interval=1):
    ora = PortWorld(n); ora.set_plane(0, hm); ora.srand(1)
    t = time.time(); so = 0
    for _ in range(frames): so += ora.erode(cycles).steps
    to = time.time() - t
    w = swe.World(n); w.upload(hm); w.srand(1); w.set_option(swe.OPT_CASCADE_INTERVAL, interval)
    t = time.time()
    for _ in range(frames): w.erode(cycles, swe.MODE_PARALLEL)
    hg = w.download(); tg = time.time() - t
    st = w.stats(); ho = np.array(ora.plane(0))
    d_o, d_g = ho - hm, hg - hm
    print(f"[{label}] n={n} frames={frames} cycles={cycles} K={interval}: steps/particle oracle {so/(frames*cycles):.1f} gpu {st.steps/(frames*cycles):.1f} | "
          f"oracle {to:.1f}s gpu {tg:.1f}s")
    print(f"   height: oracle mean {ho.mean():.5f} std {ho.std():.5f} min {ho.min():.4f} max {ho.max():.4f} | gpu mean {hg.mean():.5f} std {hg.std():.5f} min {hg.min():.4f} max {hg.max():.4f}")
    print(f"   hist TV distance {analysis.histogram_distance(ho, hg):.4f}  spectrum corr {analysis.spectrum_similarity(ho, hg, 1, n//4):.4f}  "
          f"wavelength oracle {analysis.dominant_wavelength(ho):.1f} gpu {analysis.dominant_wavelength(hg):.1f}  delta corr {np.corrcoef(d_o.ravel(), d_g.ravel())[0,1]:.4f}")
    print(f"   max excess oracle {ora.max_excess():.4f} gpu {w.max_excess():.4f}; mass lost oracle {hm.astype(np.float64).sum()-ho.astype(np.float64).sum():.3f} gpu {st.sediment_lost:.3f}", flush=True)

if __name__ == "__main__":
    run(256, 50, 2048, generate.gaussian_hill(256), "gauss")
    run(256, 200, 2048, generate.flat_sand(256), "flat")
    run(256, 200, 2048, generate.flat_sand(256), "flat", interval=4)
    run(256, 20, 65536, generate.flat_sand(256), "flat-dense")
    run(512, 20, 4096, generate.fbm(512, seed=3), "fbm")

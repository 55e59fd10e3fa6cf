"""One-off evidence run for parity (2): long dune-formation runs, parallel GPU mode vs the sequential
oracle (bit-identical to the reference), same map, same particle counts. Prints height histogram
distance, along-wind spectrum correlation and the dominant dune wavelength of both."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplewinderosion_b200 as swe
from simplewinderosion_b200 import generate, analysis
from tests.oracle_lib import PortWorld

def run(n, frames, cycles, hm, label, check=(0.25, 0.5, 1.0)):
    ora = PortWorld(n); ora.set_plane(0, hm); ora.srand(1)
    w = swe.World(n); w.upload(hm); w.srand(1)
    marks = [int(frames * c) for c in check]
    so = 0; t0 = time.time()
    for f in range(1, frames + 1):
        so += ora.erode(cycles).steps
        w.erode(cycles, swe.MODE_PARALLEL)
        if f in marks:
            ho, hg = np.array(ora.plane(0)), w.download()
            st = w.stats()
            # detrend along the wind before looking for the dune peak
            def detr(h):
                x = np.arange(h.shape[0])[:, None]
                m = h.mean(axis=1, keepdims=True)
                p = np.polyfit(x.ravel(), m.ravel(), 3)
                return h - np.polyval(p, x)
            print(f"[{label}] n={n} frame {f}/{frames} ({cycles} particles/frame, {time.time()-t0:.0f}s): steps oracle {so} gpu {st.steps} ({st.steps/so-1:+.2%}) | "
                  f"mean h {ho.mean():.4f}/{hg.mean():.4f} std {ho.std():.4f}/{hg.std():.4f} | hist TV {analysis.histogram_distance(ho, hg):.3f} | "
                  f"spectrum corr (k=1..n/4) {analysis.spectrum_similarity(ho, hg, 1, n//4):.4f} | "
                  f"dune wavelength (detrended, k>=4) oracle {analysis.dominant_wavelength(detr(ho), 4):.1f} gpu {analysis.dominant_wavelength(detr(hg), 4):.1f} cells | "
                  f"roughness along wind {np.abs(np.diff(ho, axis=0)).mean():.5f}/{np.abs(np.diff(hg, axis=0)).mean():.5f}", flush=True)

if __name__ == "__main__":
    run(256, 1200, 2048, generate.flat_sand(256), "flat sand 256")
    run(256, 300, 2048, generate.fbm(256, seed=5), "fbm 256")

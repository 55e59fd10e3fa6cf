"""python tools/time_erode.py <edge> <particles> [batches]: device-resident World::erode timing (parallel mode)."""
import sys

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

n, particles = int(sys.argv[1]), int(sys.argv[2])
w = swe.World(n)
if len(sys.argv) > 3:
    w.set_option(swe.OPT_CALL_BATCHES, int(sys.argv[3]))
w.generate_noise(1, 8); w.normalize(); w.srand(1)
for _ in range(2):
    w.erode(particles)
w.sync()
s0 = w.stats().steps
ms = 0.0
for _ in range(3):
    w.erode(particles); ms += w.stats().last_erode_ms
st = w.stats()
print(f"n={n} particles={particles}: {ms / 3:.2f} ms/call, {(st.steps - s0) / ms * 1e3:.4e} steps/s", flush=True)

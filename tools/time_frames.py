"""Latency of the reference's own frame: World::erode(8 * tilesize) (SimpleWindErosion.cpp:217) at several
map sizes, parallel mode, device-resident."""
import sys

sys.path.insert(0, ".")
import simplewinderosion_b200 as swe  # noqa: E402

pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
print("PDL", pdl)
for n in (512, 1024, 4096, 16384):
    w = swe.World(n)
    w.set_option(swe.OPT_PDL, pdl)
    w.generate_noise(1, 8); w.normalize(); w.srand(1)
    cycles = 8 * n
    for _ in range(3):
        w.erode(cycles)
    w.sync()
    s0 = w.stats().steps
    tot = 0.0
    for _ in range(5):
        w.erode(cycles); tot += w.stats().last_erode_ms
    steps = (w.stats().steps - s0) / 5
    print(f"n={n}: erode({cycles}) {tot / 5:.2f} ms/frame, {steps / (tot / 5) * 1e3:.3e} steps/s")
    w.close()

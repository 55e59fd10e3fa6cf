"""Host-side timing of the phases of one end-to-end step (upload, erode, frame read-back): python tools/e2e_phases.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simplewinderosion_b200 as swe  # noqa: E402

n, particles = 4096, 1 << 22
w = swe.World(n)
if len(sys.argv) > 1 and sys.argv[1] == "torch":       # what bench.py does: the engine works on a torch stream
    import torch
    stream = torch.cuda.Stream()
    w.set_stream(stream.cuda_stream)
w.generate_noise(1, 8); w.normalize(); w.srand(1)


def pinned(shape):
    ptr = C.c_void_p()
    nbytes = int(np.prod(shape)) * 4
    swe._capi.check(swe.lib.swe_host_alloc(C.byref(ptr), nbytes))
    return np.frombuffer((C.c_float * (nbytes // 4)).from_address(ptr.value), dtype=np.float32).reshape(shape)


fr = [pinned((5, n, n)), pinned((5, n, n))]
fr[0][0][...] = w.download()
cur = 0
t_all = time.perf_counter()
for it in range(5):
    t = [time.perf_counter()]
    w.upload(fr[cur][0]); t.append(time.perf_counter())
    w.erode(particles, swe.MODE_PARALLEL); t.append(time.perf_counter())
    t.append(time.perf_counter())
    cur ^= 1
    w.download_frame_begin(fr[cur]); t.append(time.perf_counter())
    w.download_frame_wait(False); t.append(time.perf_counter())
    if it == 4: w.download_frame_wait(True)      # (otherwise the flow fields arrive during the next step)
    t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print(f"step {it}: upload {d[0]:.2f} erode-call {d[1]:.2f} sync {d[2]:.2f} begin {d[3]:.2f} wait-height {d[4]:.2f} wait-fields {d[5]:.2f} "
          f"total {1e3 * (t[-1] - t[0]):.2f} ms", flush=True)
print(f"5 steps: {1e3 * (time.perf_counter() - t_all) / 5:.2f} ms per step; last device erode {w.stats().last_erode_ms:.2f} ms")

#!/usr/bin/env python
"""bench.py -- headline benchmark of the erode-cycle hot path (BASELINE.json metric:
wind particle-steps/s @4096^2 map; cascade % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA arm
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU arm

A "step" is one World::erode call (track reset + all particles + EMA) in parallel mode on
the 4096^2 dune-field workload with 4,194,304 particles (BASELINE.json configs[2], in its
reference-pinned form: single height layer; rock/abrasion are engine extensions absent from
the reference snapshot). One JSON line is printed by rank 0.

  value     particle-steps/s with the map resident in HBM (CUDA events on the launch stream,
            barrier + synchronize on both sides, max over ranks)
  e2e       the same metric through the host-buffer C-ABI: every step uploads the height
            plane from pinned host memory, erodes, and downloads height + the four flow
            fields (what the reference's renderer reads each frame)
  roofline  the flight kernel (single GPU: tile_fly_kernel, one launch per super-round; stripes:
            colour_persistent_kernel, one cooperative launch per call -- Wind::fly + erosion +
            World::cascade of every particle step): 104 algorithmic B/step (3x3
            heights read + written, four tracks read + written) / its device time, against
            MEASURED_PEAKS.json hbm_gbs; cascade_roofline is the same for the whole-map cascade
            sweep (8 B/cell/sweep) on the 8192^2 random-slope workload
  schedule  the headline runs the tile-resident scheduler (SWE_OPT_SCHEDULE = 2, parity-green
            against the sequential oracle at this density, tests/test_gpu_tiles.py), on one GPU and
            on the stripes of a multi-GPU run (tests/test_gpu_stripes.py); the colour-queue scheduler
            (equally parity-green) and round 1's lock-step schedule, which exports +24 % more sand
            than the reference, are timed beside it as `colour_queue_schedule` and
            `lockstep_all_at_once`
  cpu_baseline  the unmodified reference (oracle/_ref) timed on one host core on a bounded
            sample of the same workload (the reference is single-threaded, world.hpp:100)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MAP = 4096
PARTICLES = 4 * 1024 * 1024
CASCADE_N = 8192
BYTES_PER_STEP = 104.0          # 9x4 read + 9x4 write + 4 tracks x (4+4)
BYTES_PER_CELL_SWEEP = 8.0      # read h + write h
METRIC = "wind particle-steps/sec @4096^2 map"
UNIT = "particle-steps/s"
WORKLOAD = ("cfg3-pinned: 4096x4096 dune field = map::init's eight FastNoiseLite OpenSimplex2-FBm layers (seed 1, "
            "cellpool.h:388-428), 4194304 particles per World::erode, parallel mode, single height layer "
            "(rock/abrasion extension off)")


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set
    full capture (profiles/traffic.json, written by tools/make_profiles.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["bytes_per_launch"].get(kernel)
    except Exception:
        return None


def measured_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.p, self.lines = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU reference workers (oracle/_ref = the unmodified reference compiled headless; test
# infrastructure executed here only as the reported baseline, never as the product)
# ------------------------------------------------------------------------------------------
def _ref_lib_path(n: int) -> str | None:
    p = os.path.join(ROOT, "oracle", "_ref", f"libswe_ref_{n}.so")
    return p if os.path.exists(p) else None


def _cpu_worker(kind, n, map_path, cycles, n_steps, barrier, queue, core):
    """One single-threaded replica of the reference erode loop on its own copy of the map."""
    import numpy as np
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    hm = np.load(map_path)
    if kind == "reference":
        L = C.CDLL(_ref_lib_path(n))
        L.ref_set_plane.argtypes = [C.c_int, C.c_void_p]
        L.ref_erode_counted.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
        L.ref_init(1)
        L.ref_set_plane(0, np.ascontiguousarray(hm).ctypes.data)
        st, lost = (C.c_longlong * 3)(), C.c_double()

        def step():
            L.ref_erode_counted(cycles, st, C.byref(lost))
            return int(st[0])
    else:
        from tests.oracle_lib import PortWorld
        w = PortWorld(n)
        w.set_plane(0, hm)
        w.srand(1)

        def step():
            return int(w.erode(cycles).steps)
    for _ in range(n_steps):
        barrier.wait()
        t0 = time.monotonic()
        s = step()
        t1 = time.monotonic()
        queue.put((s, t0, t1))


def run_cpu_reference(n: int, map_path: str, cycles: int, n_steps: int, replicas: int):
    """Returns (kind, per-step list of (total_steps, seconds))."""
    import multiprocessing as mp
    kind = "reference" if _ref_lib_path(n) else "port"
    ctx = mp.get_context("spawn")
    barrier, queue = ctx.Barrier(replicas + 1), ctx.Queue()
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    procs = [ctx.Process(target=_cpu_worker, daemon=True,
                         args=(kind, n, map_path, cycles, n_steps, barrier, queue, cores[i % len(cores)]))
             for i in range(replicas)]
    for p in procs:
        p.start()
    out = []
    for _ in range(n_steps):
        barrier.wait()
        rs = [queue.get() for _ in range(replicas)]
        out.append((sum(r[0] for r in rs), max(r[2] for r in rs) - min(r[1] for r in rs)))
    for p in procs:
        p.join(timeout=30)
    return kind, out


def host_threads_for_reference(n: int) -> int:
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    per_replica = 9 * 4 * n * n + 3 * 4 * n * n + (256 << 20)   # AoS pool + staging + runtime
    try:
        import psutil
        by_mem = int(psutil.virtual_memory().available * 0.5 // per_replica)
    except Exception:
        by_mem = 8
    return max(1, min(ncpu, by_mem, 64))


def make_map(n: int, device: int | None = None):
    """The dune-field input of SURVEY 8(d) cfg3/cfg4: quad::map::init's eight layers of FastNoiseLite
    OpenSimplex2 FBm (cellpool.h:388-411, seed 1) normalised to [0, 0.5] (cellpool.h:416-428).
    Our arm generates it on its GPU (swe_generate_noise, bitwise identical to the host library);
    the reference arm computes the same map with the oracle port on the host cores."""
    if device is not None:
        import simplewinderosion_b200 as swe
        w = swe.World(n, device=device)
        w.generate_noise(1, 8)
        w.normalize()
        hm = w.download()
        w.close()
        return hm
    from tests.oracle_lib import noise_map
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    return noise_map(n, seed=1, layers=8, workers=max(1, min(ncpu, 64)))


def save_map_tmp(hm) -> str:
    import tempfile
    import numpy as np
    fd, path = tempfile.mkstemp(prefix="swe_bench_map_", suffix=".npy")
    os.close(fd)
    np.save(path, hm)
    return path


# ------------------------------------------------------------------------------------------
def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hm = make_map(N_MAP)
    path = save_map_tmp(hm)
    replicas = host_threads_for_reference(N_MAP)
    cycles = args.ref_cycles     # bounded sample of the 4,194,304-particle call, per replica
    try:
        kind, steps = run_cpu_reference(N_MAP, path, cycles, args.warmup + args.steps, replicas)
    finally:
        os.unlink(path)
    timed = steps[args.warmup:]
    tot_steps = sum(s for s, _ in timed)
    tot_time = sum(t for _, t in timed)
    value = tot_steps / tot_time
    sample = (f"{replicas} independent single-threaded replicas of the reference erode loop "
              f"(the reference is non-reentrant, world.hpp:100), each World::erode({cycles}) per step "
              f"on its own copy of the 4096^2 map")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": replicas, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def cascade_roofline(swe, device: int, peak: float, peak_src: str, quick: bool) -> dict:
    from simplewinderosion_b200 import generate
    n = CASCADE_N
    w = swe.World(n, device=device)
    w.upload(generate.random_slopes(n, seed=42))
    bytes_per_sweep = BYTES_PER_CELL_SWEEP * n * n
    first, total_ms, sweeps = [], 0.0, 0
    ex = w.max_excess()
    for i in range(3):                       # warm-up sweeps are part of the relaxation: timed too
        w.cascade_sweep(1); ms = w.stats().last_sweep_ms
        first.append(ms); total_ms += ms; sweeps += 1
    max_sweeps = 40 if quick else 400
    while sweeps < max_sweeps:
        k = 8
        ex = w.cascade_sweep(k, want_excess=True)
        ms = w.stats().last_sweep_ms
        total_ms += ms; sweeps += k
        last_ms = ms / k
        if ex < 1e-4:
            break
    w.cascade_sweep(8); relaxed_ms = w.stats().last_sweep_ms / 8
    # a map at rest (every slope below the repose angle): the pure cost of proving that nothing moves
    w.upload(generate.flat_sand(n, seed=1))
    w.cascade_sweep(2)
    w.cascade_sweep(8); rest_ms = w.stats().last_sweep_ms / 8
    w.close()
    achieved = bytes_per_sweep * sweeps / (total_ms * 1e-3) / 1e9
    return {"bound": "hbm", "workload": f"cfg5: {n}x{n} U(0,1) random-slope map relaxed until max excess < 1e-4",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": ncu_traffic("cascade_sweep8192"), "sweeps": sweeps,
            "final_max_excess": ex,
            "ms_per_sweep_mean": total_ms / sweeps, "ms_first_sweep": first[0],
            "ms_per_sweep_relaxed": relaxed_ms,
            "frac_relaxed": bytes_per_sweep / (relaxed_ms * 1e-3) / 1e9 / peak,
            "ms_per_sweep_at_rest": rest_ms,
            "frac_at_rest": bytes_per_sweep / (rest_ms * 1e-3) / 1e9 / peak}


def dune_formation_extra(swe, device: int, steps: int) -> dict:
    """BASELINE.json configs[1]: 1024x1024 flat-sand dune formation, 1,048,576 particles per
    World::erode (height = 0.25 + 1e-3 U(-1,1)). Reported beside the headline."""
    import torch
    from simplewinderosion_b200 import analysis, generate
    n, particles = 1024, 1 << 20
    w = swe.World(n, device=device)
    w.upload(generate.flat_sand(n, seed=42))
    w.srand(1)
    stream = torch.cuda.Stream()
    w.set_stream(stream.cuda_stream)
    for _ in range(3):
        w.erode(particles, swe.MODE_PARALLEL)
    w.sync()
    s0 = w.stats().steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            w.erode(particles, swe.MODE_PARALLEL)
        ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    done = w.stats().steps - s0
    h = w.download()
    w.close()
    return {"workload": f"cfg2: {n}x{n} flat sand (0.25 + 1e-3 U(-1,1)), {particles} particles per World::erode, "
                        f"{3 + steps} calls from the flat bed",
            "value": done / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "dominant_wavelength_cells": analysis.dominant_wavelength(h),
            "height_std": float(h.std())}


def layered_extra(swe, device: int, n: int, particles: int, hm, steps: int) -> dict:
    """BASELINE.json configs[2] as named: dune field with a rock pyramid obstacle and abrasion on.
    Rock/abrasion do not exist in the reference snapshot, so this runs the engine's layered-terrain
    EXTENSION (parity unpinned) and is reported beside the headline, not as the headline."""
    import numpy as np
    import torch
    from simplewinderosion_b200 import generate
    rock = generate.pyramid(n, apex=0.5)
    height = np.maximum(hm, rock).astype(np.float32)
    w = swe.World(n, device=device)
    w.set(abrasion=0.1)
    w.upload(height)
    w.upload_rock(rock)
    w.srand(1)
    for _ in range(2):
        w.erode(particles, swe.MODE_PARALLEL)
    w.sync()
    s0 = w.stats().steps
    ms = 0.0
    for _ in range(steps):
        w.erode(particles, swe.MODE_PARALLEL)
        ms += w.stats().last_erode_ms
    out = {"workload": f"cfg3 as named: {n}x{n} FBm dune field over a rock pyramid (base n/4, apex 0.5), abrasion 0.1, "
                       f"{particles} particles per World::erode; layered-terrain extension, parity unpinned",
           "value": (w.stats().steps - s0) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps}
    w.close()
    torch.cuda.synchronize()
    return out


def cfg4_strong(swe, rank: int, world: int, device: int, barrier, max_over_ranks, sum_over_ranks) -> dict:
    """BASELINE.json configs[3]: a 16384 x 16384 map row-tiled over the GPUs, 16,777,216 particles per World::erode
    (0.0625 per cell), generated on the GPUs (map::init's noise layers, collective normalisation). With N > 1 rank 0
    also times the SAME call on its single GPU, so that the strong-scaling speed-up is measured inside one run."""
    import torch
    n, particles, steps = 16384, 1 << 24, 2

    def timed(w):
        w.srand(1)
        w.erode(particles, swe.MODE_PARALLEL)          # warm-up
        w.sync()
        s0 = w.stats().steps
        stream = torch.cuda.Stream()
        w.set_stream(stream.cuda_stream)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        return s0, stream, ev0, ev1

    one = one_cq = None
    if rank == 0:
        # the same call on ONE GPU: on the tile-resident schedule (the fastest single-GPU path) and, with N > 1, also on
        # the colour-queue schedule the stripes run, so that both speed-ups can be read off one run
        for sched in ([swe.SCHEDULE_TILES] if world == 1 else [swe.SCHEDULE_TILES, swe.SCHEDULE_COLOUR_QUEUES]):
            w1 = swe.World(n, device=device)
            w1.set_option(swe.OPT_SCHEDULE, sched)
            w1.generate_noise(1, 8); w1.normalize()
            s0, stream, ev0, ev1 = timed(w1)
            with torch.cuda.stream(stream):
                ev0.record(stream)
                for _ in range(steps):
                    w1.erode(particles, swe.MODE_PARALLEL)
                ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            r = {"value": (w1.stats().steps - s0) / (ms * 1e-3), "ms_per_step": ms / steps,
                 "schedule": "tile-resident" if sched == swe.SCHEDULE_TILES else "colour queues"}
            w1.close()
            if sched == swe.SCHEDULE_TILES:
                one = r
            else:
                one_cq = r
    out = {"workload": f"cfg4: {n}x{n} FBm dune field, {particles} particles per World::erode (0.0625 per cell), "
                       f"tile-resident schedule{'' if world == 1 else ' on row stripes'}, {steps} timed calls after 1 warm-up",
           "unit": UNIT, "n_gpus": world}
    if world == 1:
        out.update(one)
        return out
    from simplewinderosion_b200 import stripes
    barrier()
    w = stripes.StripedWorld(n, n, rank, world, particles, device=device)
    w.connect()
    w.generate_noise(1, 8)
    w.normalize_collective()
    w.sync_halos()
    s0, stream, ev0, ev1 = timed(w)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            w.erode(particles, swe.MODE_PARALLEL)
        ev1.record(stream)
    stream.synchronize()          # (see run_ours: never an NCCL barrier while stripe erodes are in flight)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    st = w.stats()
    total = sum_over_ranks(float(st.steps - s0))
    dropped = sum_over_ranks(float(st.queue_dropped + st.stripe_overflow))
    w.close()
    out.update({"value": total / (ms * 1e-3), "ms_per_step": ms / steps, "dropped_particles": int(dropped),
                "schedule": "tile-resident" if st.last_super_rounds > 0 else "colour queues"})
    if rank == 0:
        out["one_gpu_same_run"] = one                         # the same call on one GPU (tile-resident schedule)
        out["one_gpu_colour_queues"] = one_cq                 # ... and on the colour-queue schedule, for reference
        out["speedup_vs_1gpu"] = out["value"] / one["value"]
    return out


def colour_queue_extra(swe, device: int, n: int, particles: int, hm, steps: int) -> dict:
    """The colour-queue scheduler (SWE_OPT_SCHEDULE = 0; the schedule of the multi-GPU stripes and of periodic maps)
    on the headline call, timed beside the tile-resident headline."""
    w = swe.World(n, device=device)
    w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_COLOUR_QUEUES)
    w.upload(hm)
    w.srand(1)
    for _ in range(2):
        w.erode(particles, swe.MODE_PARALLEL)
    w.sync()
    s0 = w.stats().steps
    ms = 0.0
    for _ in range(steps):
        w.erode(particles, swe.MODE_PARALLEL)
        ms += w.stats().last_erode_ms
    out = {"schedule": "colour queues: one persistent cooperative kernel, grid barrier per micro-round (parity-green, "
                       "tests/test_gpu_density.py)",
           "value": (w.stats().steps - s0) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "micro_rounds_per_step": w.stats().last_micro_rounds}
    w.close()
    return out


def lockstep_extra(swe, device: int, n: int, particles: int, hm, steps: int) -> dict:
    """Round 1's schedule (SWE_OPT_SCHEDULE = 1): every particle of the call steps once per round, then one
    stamped cascade over the touched cells. Timed beside the headline because it is what round 1 reported;
    it is NOT parity-green at this density (tests/test_gpu_density.py, profiles/r1_stat_parity_density.txt)."""
    w = swe.World(n, device=device)
    w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_LOCKSTEP)
    w.upload(hm)
    w.srand(1)
    for _ in range(2):
        w.erode(particles, swe.MODE_PARALLEL)
    w.sync()
    s0 = w.stats().steps
    ms = 0.0
    for _ in range(steps):
        w.erode(particles, swe.MODE_PARALLEL)
        ms += w.stats().last_erode_ms
    out = {"schedule": "lock-step rounds, all particles in flight at once (round 1)",
           "value": (w.stats().steps - s0) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "parity": "fails: +24 % exported sand, erosion-field correlation 0.80 against the sequential oracle at 0.25 "
                     "particles per cell"}
    w.close()
    return out


def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import simplewinderosion_b200 as swe

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    out_stream = sys.stdout
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the first communicator is
        # created, so everything else that lands on fd 1 during the run is sent to stderr
        out_stream = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = measured_peak_gbs()
    n, particles = args.size, args.particles
    hm = make_map(n, device=local)

    if world == 1:
        w = swe.World(n, device=local)
        w.set_option(swe.OPT_SCHEDULE, swe.SCHEDULE_TILES)
    else:
        # weak scaling: every GPU owns one n x n stripe of an (n*world) x n map (copies of the
        # single-GPU map stacked along the wind axis) and the call flies particles*world particles.
        # Neighbours exchange ghost rows and hand over particles by P2P kernels (no NCCL on the
        # data path; torch.distributed only ships the CUDA IPC handles).
        from simplewinderosion_b200 import stripes
        if args.strong:   # cfg4-style: ONE size x size map cut into stripes, `particles` in total
            w = stripes.StripedWorld(n, n, rank, world, particles, device=local)
        else:
            w = stripes.StripedWorld(n * world, n, rank, world, particles * world, device=local)
        w.connect()
    stream = torch.cuda.Stream()
    w.set_stream(stream.cuda_stream)
    if world > 1 and args.strong:
        w.upload_global(hm)
    else:
        w.upload(hm)
    if world > 1:
        w.sync_halos()
    stripe_check = None
    if world > 1:
        # the real-peer bitwise check (the 1-GPU test box skips tests/test_gpu_stripes.py::test_real_multi_gpu): three
        # striped whole-map cascade sweeps against a single engine sweeping the same global map on this GPU
        gmap = hm if args.strong else np.concatenate([hm] * world, axis=0)
        ref = swe.World(gmap.shape[0], n, device=local)
        ref.upload(gmap)
        ref.cascade_sweep(3)
        w.cascade_sweep(3)
        lo, hi = w.bounds[rank]
        same = bool(np.array_equal(w.download().view(np.uint32), ref.download()[lo:hi].view(np.uint32)))
        ref.close()
        stripe_check = "bitwise" if sum_over_ranks(0.0 if same else 1.0) == 0.0 else "MISMATCH"
        if args.strong:
            w.upload_global(hm)
        else:
            w.upload(hm)
        w.sync_halos()
    w.srand(1)
    # GLOBAL particle count of one World::erode call
    particles_call = particles if (world == 1 or args.strong) else particles * world
    rows_local = w.nx

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(args.warmup):
        w.erode(particles_call, swe.MODE_PARALLEL)
    w.sync()
    st0 = w.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fly_ms = 0.0
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            w.erode(particles_call, swe.MODE_PARALLEL)
        ev1.record(stream)
    # the erodes are asynchronous and, on stripes, their persistent kernels wait for the other ranks ON THE DEVICE while
    # holding every SM: drain them before the NCCL barrier, whose kernel could otherwise be scheduled between two ranks'
    # launches and wait for a peer that is itself waiting for this GPU (the engine's watchdog would turn that into an error)
    stream.synchronize()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    st1 = w.stats()
    steps_done = st1.steps - st0.steps
    launches = st1.kernel_launches - st0.kernel_launches
    total_steps = sum_over_ranks(float(steps_done))
    total_particles = sum_over_ranks(float(st1.particles - st0.particles))
    overflow_total = sum_over_ranks(float(st1.stripe_overflow))
    value = total_steps / (dev_ms * 1e-3)

    # device time of the particle phase (CUDA events on the launch stream around the flight kernel(s) of each
    # call, swe_stats.last_fly_ms): an extra pass outside the timed region, one stats read-back per call
    lockstep = False              # stripes wired as a ring run the colour-queue schedule too
    if lockstep:
        w.set_option(swe.OPT_KERNEL_TIMING, 1)
    round_ms = casc_ms = fly_ms = erode_ms = 0.0
    fly_steps = 0
    micro_rounds = super_rounds = 0
    for _ in range(args.steps):
        s_before = w.stats().steps
        w.erode(particles_call, swe.MODE_PARALLEL)
        s = w.stats()
        round_ms += s.last_round_ms; casc_ms += s.last_cascade_ms; fly_ms += s.last_fly_ms; erode_ms += s.last_erode_ms
        fly_steps += s.steps - s_before
        micro_rounds += s.last_micro_rounds
        super_rounds += s.last_super_rounds
    if lockstep:
        w.set_option(swe.OPT_KERNEL_TIMING, 0)
    rounds_per_call = w.params.max_age + 2
    achieved = BYTES_PER_STEP * fly_steps / (fly_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI ---------------------------------------------
    def pinned(shape):
        ptr = C.c_void_p()
        nbytes = int(np.prod(shape)) * 4
        swe._capi.check(swe.lib.swe_host_alloc(C.byref(ptr), nbytes))
        buf = (C.c_float * (nbytes // 4)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.float32).reshape(shape), ptr

    # two pinned frame buffers: the height plane of frame k is back before the host hands the map to step k + 1, its four
    # flow-field planes (snapshotted on the device) arrive on a side stream WHILE step k + 1 erodes (swe_download_frame_begin)
    frames = [pinned((5, rows_local, n)), pinned((5, rows_local, n))]
    cur = 0
    frames[cur][0][0][...] = w.download()

    def e2e_step():
        nonlocal cur
        # the host owns the canonical map between calls: the height plane of the previous read-back is what goes up
        w.upload(frames[cur][0][0])
        if world > 1:
            w.sync_halos()
        w.erode(particles_call, swe.MODE_PARALLEL)
        cur ^= 1
        w.download_frame_begin(frames[cur][0])   # height + massflow + momentum x/y/z: what the renderer reads each frame
        w.download_frame_wait(False)             # (height plane: the next step uploads it)

    e2e_step()
    w.download_frame_wait(True)
    s0 = w.stats().steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    w.download_frame_wait(True)                  # the last frame's flow fields are inside the timed region
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_steps = sum_over_ranks(float(w.stats().steps - s0))
    e2e_value = e2e_steps / e2e_s
    mean_h = float(frames[cur][0][0].mean())
    fields_ok = bool(np.isfinite(frames[cur][0][1:]).all() and float(frames[cur][0][1].max()) > 0.0)
    swe.lib.swe_host_free(frames[0][1])
    p_out = frames[1][1]
    swe.lib.swe_host_free(p_out)
    w.close()

    # ---- BASELINE cfg4: ONE 16384^2 map, 16,777,216 particles per call, cut into `world` row stripes (strong scaling)
    cfg4 = None
    if not args.skip_cfg4 and not args.strong and n == N_MAP:
        cfg4 = cfg4_strong(swe, rank, world, local, barrier, max_over_ranks, sum_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD if (n == N_MAP and particles == PARTICLES and not args.strong) else
            f"{n}x{n} FBm dune field, {particles} particles per World::erode, parallel mode",
            "map": [n if args.strong else n * world, n], "particles_per_step": particles_call,
            "parallelism": "single GPU" if world == 1 else
            f"{world} row stripes of a {n if args.strong else n * world}x{n} map ({rows_local} rows and "
            f"~{particles_call // world} particles per GPU), "
            + ("tile-resident schedule on a GLOBAL tiling: every GPU launches the tiles whose centre row it owns; a tile that "
               "straddles a stripe boundary loads / stores the neighbour's rows in the neighbour's memory over NVLink once per "
               "super-round and adds their track vectors there (system-scope reductions); surviving particles whose next tile "
               "belongs to a neighbour are counted and scattered into the neighbour's segment tables (peer atomics and stores); "
               "three barrier kernels per super-round (mailbox words in peer memory), no NCCL on the data path"
               if super_rounds > 0 else
               "one persistent colour-queue kernel per GPU; grid barrier across the GPUs every micro-round (mailbox words in peer "
               "memory over NVLink); 3x3 blocks on a stripe boundary are read and written in the neighbour's memory; particles "
               "crossing a boundary are appended to the neighbour's queue (system-scope atomics)"),
            "stripe_overflow": int(overflow_total),
            **({} if world == 1 else {"exchange": ({
                "barriers_per_step": 3 * (super_rounds / max(1, args.steps) + 1),
                "what": "per super-round and stripe boundary: the rows of the straddling tiles beyond the boundary (<= 2/3 of a "
                        "tile edge x nz floats, read once and written back once), 32 B per particle whose next tile belongs to "
                        "the neighbour, three 8-byte barrier words to every rank; one halo refresh (12 rows) per call",
                "super_rounds_per_step": super_rounds / max(1, args.steps)} if super_rounds > 0 else {
                "per_micro_round_per_rank_bytes": int(8 * world),
                "what": "one 8-byte barrier word to every rank per micro-round; 32 B per particle handed to a neighbour's queue; "
                        "4 B loads/stores of the neighbour's boundary rows by the particles standing next to them; "
                        "one halo refresh (12 rows) per call", "micro_rounds_per_step": micro_rounds / max(1, args.steps)})}),
            "l2": "working set > 600 MiB per GPU (height 64 + track 256 + field 256 MiB + live particle records + claim lattice) "
                  "> 126 MB L2; no flush needed",
            "steps_per_particle": total_steps / max(1.0, total_particles),
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * rows_local * n),
                "d2h_bytes_per_step": int(5 * 4 * rows_local * n), "ms_per_step": 1e3 * e2e_s / args.steps,
                "result_check_mean_height": mean_h, "result_check_flow_fields_arrived": fields_ok,
                "pipelining": "height plane D2H on the engine's stream; the four flow-field planes of frame k are snapshotted on "
                              "the device and copied on a side stream while step k+1 erodes (all inside the timed region)"},
        "gpu_launches": int(launches),
        "roofline": ({
            "bound": "hbm", "kernel": "tile_fly_kernel (Wind::fly + erosion + World::cascade per particle step; one CTA per 120x120 tile, "
                                      "heights and particles in shared memory)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic("tile_fly"), "peak_source": peak_src,
            "algorithmic_bytes_per_step": BYTES_PER_STEP,
            "launches_per_step": super_rounds / args.steps,
            "kernel_ms_per_launch": fly_ms / max(1, super_rounds),
            "algorithmic_bytes_per_launch": BYTES_PER_STEP * fly_steps / max(1, super_rounds),
            "share_of_step": {"tile_fly_kernel + scan + scatter (one super-round each)": fly_ms / erode_ms,
                              "track reset + EMA": 1.0 - fly_ms / erode_ms},
            "micro_rounds_per_step": micro_rounds / args.steps,
            "note": "104 B/step is SURVEY 8(d)'s figure for flight + cascade, which this one kernel does. `traffic` is the DRAM "
                    "traffic of ONE mid-call launch under ncu (the launches of a call differ in size: the injection ramps "
                    "up and down); per particle step it is ~4x BELOW the algorithmic bytes because a tile's heights are "
                    "read and written once per super-round, not per step. The kernel is bound by instruction issue and "
                    "the dependent chain of a step, not by HBM (DESIGN.md 4)"}
            if super_rounds > 0 else {
            "bound": "hbm", "kernel": "colour_persistent_kernel (Wind::fly + erosion + World::cascade per particle step)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic("colour_persistent"), "peak_source": peak_src,
            "algorithmic_bytes_per_step": BYTES_PER_STEP, "launches_per_step": 1,
            "kernel_ms_per_launch": fly_ms / args.steps,
            "share_of_step": {"colour_persistent_kernel": fly_ms / erode_ms, "track reset + EMA": 1.0 - fly_ms / erode_ms},
            "micro_rounds_per_step": micro_rounds / args.steps,
            "note": "104 B/step is SURVEY 8(d)'s figure for flight + cascade, which this one kernel does; it is bound by "
                    "dependent-access latency and the grid barrier per micro-round, not by HBM bandwidth (DESIGN.md 4)"}
            if not lockstep else {
            "bound": "hbm", "kernel": "par_round_kernel + cascade_tile_stamped_kernel + cascade_commit_kernel (+ stripe exchange)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_step": BYTES_PER_STEP,
            "launches_per_step": 6 * rounds_per_call,
            "share_of_step": {"par_round_kernel": round_ms / max(fly_ms, 1e-9),
                              "halo exchange+hand-off+cascade": casc_ms / max(fly_ms, 1e-9)}}),
    }
    line["config"]["schedule"] = (
        "lock-step rounds (stripe engines; round 1's schedule: NOT parity-green at this density, DESIGN.md 6)" if lockstep else
        f"tile-resident: 120x120-cell tiles, one CTA per tile for {micro_rounds // max(1, 9 * super_rounds)} rounds of nine colour "
        f"micro-rounds per launch, three "
        f"shifted tilings, interior-only service, in-register cascade, automatic stagger "
        f"({int(8 * particles / (n * n) * rounds_per_call + 0.5)} rounds); parity-green at this density "
        f"(tests/test_gpu_tiles.py)" + ("" if world == 1 else "; on stripes: global tiling, a tile belongs to the rank that owns its "
        "centre row, three cross-GPU barriers per super-round (tests/test_gpu_stripes.py)") if super_rounds > 0 else
        f"colour queues: 9-colour micro-rounds, one particle per cell, in-register cascade, automatic stagger "
        f"({int(8 * particles / (n * n) * rounds_per_call + 0.5)} rounds); parity-green at this density "
        f"(tests/test_gpu_density.py)")
    if world > 1:
        line["stripe_check"] = stripe_check
    if cfg4 is not None:
        line["cfg4_strong"] = cfg4
    if world == 1:
        line["colour_queue_schedule"] = colour_queue_extra(swe, local, n, particles, hm, args.steps)
        line["lockstep_all_at_once"] = lockstep_extra(swe, local, n, particles, hm, args.steps)
    if world == 1 and not args.skip_layered:
        line["layered_extension"] = layered_extra(swe, local, n, particles, hm, args.steps)
        line["dune_formation_cfg2"] = dune_formation_extra(swe, local, args.steps)
    if not args.skip_cascade and world == 1:
        line["cascade_roofline"] = cascade_roofline(swe, local, peak, peak_src, args.quick)
        # the other half of BASELINE's metric ("cascade % HBM roofline"), also as an entry of `roofline`
        c = line["cascade_roofline"]
        line["roofline"]["cascade_stencil"] = {
            "bound": "hbm", "kernel": "cascade_tile_kernel (whole-map nine-colour sweep, 8 B/cell)", "achieved": c["achieved"],
            "peak": c["peak"], "unit": "GB/s", "frac": c["frac"], "frac_relaxed": c["frac_relaxed"],
            "frac_at_rest": c["frac_at_rest"], "traffic": c["traffic"]}
    if world == 1 and not args.skip_cpu:
        path = save_map_tmp(hm)
        try:
            kind, steps = run_cpu_reference(n, path, args.cpu_cycles, 1, 1)
        finally:
            os.unlink(path)
        s, t = steps[0]
        line["cpu_baseline"] = {
            "value": s / t, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"one World::erode({args.cpu_cycles}) = {s} particle-steps on the same {n}^2 map, "
                      f"1 thread (the reference is single-threaded), {t:.1f} s"}
    print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=N_MAP)
    ap.add_argument("--particles", type=int, default=PARTICLES)
    ap.add_argument("--cpu-cycles", type=int, default=32768, help="particles of the 1-core CPU baseline sample")
    ap.add_argument("--ref-cycles", type=int, default=4096, help="particles per replica per step, --impl reference")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-cascade", action="store_true")
    ap.add_argument("--skip-layered", action="store_true")
    ap.add_argument("--skip-cfg4", action="store_true", help="do not run the 16384^2 strong-scaling case (BASELINE cfg4)")
    ap.add_argument("--quick", action="store_true", help="shorter cascade relaxation")
    ap.add_argument("--strong", action="store_true",
                    help="N>1: cut ONE size x size map into N stripes (cfg4: --size 16384 --particles 16777216) "
                         "instead of the default weak scaling (one size x size stripe per GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

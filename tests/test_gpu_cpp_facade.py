"""GPU: the header-only C++ façade (include/swe_world.hpp) that re-creates the reference's
World/Wind names over the C-ABI: compile examples/headless_main.cpp -- a headless restatement
of the reference's main loop (SimpleWindErosion.cpp:27-42,211-231) -- with g++, run it in
deterministic mode and compare with the oracle bit for bit."""
import os
import subprocess

import pytest

from tests.oracle_lib import PortWorld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "headless")
    lib = os.path.join(ROOT, "simplewinderosion_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "headless_main.cpp"), "-L", lib, "-lswe_b200",
                           f"-Wl,-rpath,{lib}", "-o", exe])
    return exe


def test_facade_compiles_without_gpu(tmp_path):
    """CPU: the façade is plain C++17 over the C-ABI; without a device it fails loudly."""
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "64", "1", "1", "8", "0"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_main_loop_matches_oracle(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.check_output([exe, "128", "1", "2", "1024", "0"], text=True)
    got = dict(kv.split("=") for kv in out.split())
    ora = PortWorld(128)
    ora.init_gaussian()
    ora.srand(1)
    steps = sum(ora.erode(1024).steps for _ in range(2))
    assert int(got["steps"]) == steps
    assert got["hash"] == f"{ora.hash_height():016x}"
    import numpy as np
    assert np.float32(got["centre_height"]) == ora.plane(0)[32, 64]
    assert float(got["centre_momentumx"]) == pytest.approx(float(ora.plane(2)[32, 64]), rel=1e-7)


@pytest.mark.gpu
def test_cpp_main_loop_parallel_mode_runs(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.check_output([exe, "256", "3", "2", "2048", "1"], text=True)
    got = dict(kv.split("=") for kv in out.split())
    assert int(got["steps"]) > 100000 and int(got["oob"]) + int(got["age"]) == 4096


@pytest.mark.gpu
def test_cpp_init_with_noise_layers_matches_oracle(tmp_path):
    """map::init with the reference's commented noise layers enabled: Gaussian hill on the host,
    eight FastNoiseLite layers and the normalisation on the device, then a deterministic erode."""
    exe = _build(tmp_path)
    out = subprocess.check_output([exe, "64", "7", "1", "256", "0", "8"], text=True)
    got = dict(kv.split("=") for kv in out.split())
    raw = PortWorld(64)
    raw.hill_raw()
    raw.noise_stack(7, 8)
    raw.normalize()
    raw.srand(7)
    steps = raw.erode(256).steps
    assert int(got["steps"]) == steps
    assert got["hash"] == f"{raw.hash_height():016x}"


def _build_compat(tmp_path):
    exe = str(tmp_path / "compat")
    lib = os.path.join(ROOT, "simplewinderosion_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "helpers", "frame_loop_compat.cpp"), "-L", lib, "-lswe_b200",
                           f"-Wl,-rpath,{lib}", "-o", exe])
    return exe


def test_reference_caller_lines_compile_verbatim(tmp_path):
    """CPU: the reference application's own lines (SimpleWindErosion.cpp:11-12,26-42,211-231, quoted verbatim in
    tests/helpers/frame_loop_compat.cpp) compile unchanged against include/swe_world.hpp: cellpool.reserve(quad::area),
    World::map.init(vertexpool, cellpool, SEED), world.erode(8*quad::tilesize), `for(auto& node: world.map.nodes)
    updatenode(vertexpool, node)`, world.map.get(p)->get(p)->momentumx, quad::res."""
    exe = _build_compat(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "64", "1", "1"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_frame_loop_runs_unchanged_and_matches_oracle(tmp_path):
    """GPU: that verbatim frame loop, run in deterministic mode, leaves the map bitwise equal to the oracle; the
    vertex pool received the reference's updatenode records and the wind-map lambda read current momentum values
    out of the host cell pool (World::autosync)."""
    import math
    import numpy as np
    exe = _build_compat(tmp_path)
    n, seed, frames = 128, 3, 2
    out = subprocess.check_output([exe, str(n), str(seed), str(frames)], text=True)
    got = dict(kv.split("=") for kv in out.strip().splitlines()[-1].split())
    ora = PortWorld(n)
    ora.init_gaussian()
    ora.srand(seed)
    steps = sum(ora.erode(8 * n).steps for _ in range(frames))
    assert int(got["steps"]) == steps                           # (read before the probe particle flies)
    assert got["hash"] == f"{ora.hash_height():016x}"
    cx, cz = n // 4, n // 2
    h = ora.plane(0)[cx, cz]
    assert np.float32(got["cell_height"]) == h
    assert np.float32(got["vertex_y"]) == np.float32(80.0) * h
    assert float(got["normal_y"]) == pytest.approx(float(ora.normal(cx, cz)[1]), rel=1e-6)
    assert float(got["tangent_x"]) == 1.0
    assert float(got["wind_r"]) == pytest.approx(0.5 * (1.0 + math.erf(2.0 * float(ora.plane(2)[cx, cz]))), rel=1e-5)
    assert got["probe_alive"] == "1" and got["probe_age"] == "1"

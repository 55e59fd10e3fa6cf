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

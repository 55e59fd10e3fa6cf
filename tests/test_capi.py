"""CPU: the C-ABI library loads and exports every symbol include/swe.h declares; host-side
argument checking and the no-GPU failure path (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import simplewinderosion_b200 as swe
from simplewinderosion_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "swe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 30
    raw = C.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/swe.h but not exported"
        assert s in _capi.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_capi.PROTOTYPES) == set(syms)


def test_only_swe_symbols_are_exported():
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", _capi.LIB_PATH], text=True)
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert names and all(n.startswith("swe_") for n in names), [n for n in names if not n.startswith("swe_")][:5]


def test_abi_version_and_defaults():
    assert swe.lib.swe_abi_version() == 3
    p = swe.default_params()
    # wind.hpp:23-26, world.hpp:30-32, cellpool.h:184, wind.hpp:13
    assert p.max_age == 512 and p.boundarylayer == 2.0 and p.mapscale == 80
    assert abs(p.suspension - 0.05) < 1e-9 and abs(p.gravity - 0.1) < 1e-8
    assert abs(p.lrate - 0.1) < 1e-8 and abs(p.maxdiff - 0.005) < 1e-9 and p.settling == 1.0
    assert list(p.pspeed) == [1.0, 0.0, 0.0]
    assert C.sizeof(swe.Params) == 64 and C.sizeof(swe.WindState) == 32
    assert swe.CELL_DTYPE.itemsize == 36   # quad::cell, cellpool.h:226-240


def test_argument_errors_are_reported():
    h = C.c_void_p()
    assert swe.lib.swe_create(C.byref(h), 0, 16, None, 0) == -1
    assert b"map size" in swe.lib.swe_last_error()
    p = swe.default_params()
    p.boundarylayer = 0.0
    assert swe.lib.swe_create(C.byref(h), 16, 16, C.byref(p), 0) == -1
    assert b"boundarylayer" in swe.lib.swe_last_error()
    assert swe.lib.swe_params_default(None) == -1
    assert swe.lib.swe_erode(None, 1, 0) == -1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(swe.SweError) as ei:
        swe.World(16, 16)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_references_oracle():
    """The product path must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "simplewinderosion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "swe_oracle" not in text and "oracle_lib" not in text and "libswe_ref" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle/', text), f
    for f in ("swe.h", "swe_world.hpp"):
        path = os.path.join(ROOT, "include", f)
        if os.path.exists(path):
            assert not re.search(r'#include\s+"[^"]*oracle', open(path).read())

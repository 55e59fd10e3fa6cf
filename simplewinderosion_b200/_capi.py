"""ctypes binding of the C-ABI declared in include/swe.h (libswe_b200.so).

The library is the product; this file only loads it and declares the prototypes. There is
no fallback: if the shared library is missing or its ABI version does not match, importing
the package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWE_B200_LIB") or os.path.join(_HERE, "libswe_b200.so")   # override: kernel-variant experiments

ABI_VERSION = 3


class SweError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"swe error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """struct swe_params (include/swe.h)"""
    _fields_ = [
        ("max_age", C.c_int32),
        ("boundarylayer", C.c_float),
        ("suspension", C.c_float),
        ("gravity", C.c_float),
        ("lrate", C.c_float),
        ("maxdiff", C.c_float),
        ("settling", C.c_float),
        ("mapscale", C.c_int32),
        ("pspeed", C.c_float * 3),
        ("abrasion", C.c_float),
        ("reserved", C.c_int32 * 4),
    ]


class WindState(C.Structure):
    """struct swe_wind (include/swe.h) == struct Wind (wind.hpp:4-21) as plain data"""
    _fields_ = [
        ("age", C.c_int32),
        ("sediment", C.c_float),
        ("pos", C.c_float * 3),
        ("speed", C.c_float * 3),
    ]


class Stats(C.Structure):
    """struct swe_stats (include/swe.h)"""
    _fields_ = [
        ("particles", C.c_uint64),
        ("steps", C.c_uint64),
        ("exits_oob", C.c_uint64),
        ("exits_age", C.c_uint64),
        ("sediment_lost", C.c_double),
        ("cascade_cells", C.c_uint64),
        ("erode_calls", C.c_uint64),
        ("last_erode_ms", C.c_double),
        ("last_sweep_ms", C.c_double),
        ("kernel_launches", C.c_uint64),
        ("last_fly_ms", C.c_double),
        ("last_round_ms", C.c_double),
        ("last_cascade_ms", C.c_double),
        ("stripe_overflow", C.c_uint64),
        ("last_micro_rounds", C.c_uint64),
        ("queue_dropped", C.c_uint64),
        ("last_super_rounds", C.c_uint64),
    ]


class StripeLink(C.Structure):
    """struct swe_stripe_link (include/swe.h)"""
    _fields_ = [("height", C.c_void_p), ("stamp", C.c_void_p), ("arena", C.c_void_p),
                ("x0", C.c_int32), ("inbox_capacity", C.c_int32),
                ("queue_a", C.c_void_p), ("queue_b", C.c_void_p), ("queue_count", C.c_void_p), ("persist", C.c_void_p),
                ("queue_cap", C.c_int32), ("reserved", C.c_int32)]


STRIPE_IPC_BYTES = 512
STATE_BYTES = 256

# name -> (restype, argtypes); must list every symbol include/swe.h declares
_E = C.c_void_p
_F = C.POINTER(C.c_float)
PROTOTYPES = {
    "swe_abi_version": (C.c_int, []),
    "swe_tile_geometry": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "swe_tile_owned_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "swe_last_error": (C.c_char_p, []),
    "swe_params_default": (C.c_int, [C.POINTER(Params)]),
    "swe_create": (C.c_int, [C.POINTER(_E), C.c_int, C.c_int, C.POINTER(Params), C.c_int]),
    "swe_destroy": (None, [_E]),
    "swe_set_params": (C.c_int, [_E, C.POINTER(Params)]),
    "swe_get_params": (C.c_int, [_E, C.POINTER(Params)]),
    "swe_get_size": (C.c_int, [_E, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "swe_set_option": (C.c_int, [_E, C.c_int, C.c_int]),
    "swe_stripe_create": (C.c_int, [C.POINTER(_E), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(Params), C.c_int]),
    "swe_stripe_export": (C.c_int, [_E, C.POINTER(StripeLink)]),
    "swe_stripe_export_ipc": (C.c_int, [_E, C.c_void_p, C.c_size_t]),
    "swe_stripe_open_ipc": (C.c_int, [_E, C.c_void_p, C.c_size_t, C.POINTER(StripeLink)]),
    "swe_stripe_connect": (C.c_int, [_E, C.POINTER(StripeLink), C.POINTER(StripeLink)]),
    "swe_stripe_stored_rows": (C.c_int, [_E, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "swe_stripe_connect_ring": (C.c_int, [_E, C.c_int, C.c_int, C.POINTER(StripeLink), C.c_int]),
    "swe_stripe_sync_halos": (C.c_int, [_E]),
    "swe_stripe_rows": (C.c_int, [_E, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "swe_stripe_group_erode": (C.c_int, [C.POINTER(_E), C.c_int, C.c_int]),
    "swe_stripe_group_sweep": (C.c_int, [C.POINTER(_E), C.c_int, C.c_int]),
    "swe_stripe_group_sync_halos": (C.c_int, [C.POINTER(_E), C.c_int]),
    "swe_srand": (C.c_int, [_E, C.c_uint32]),
    "swe_export_state": (C.c_int, [_E, C.c_void_p, C.c_size_t]),
    "swe_import_state": (C.c_int, [_E, C.c_void_p, C.c_size_t]),
    "swe_upload_plane": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_ssize_t]),
    "swe_download_plane": (C.c_int, [_E, C.c_int, C.c_void_p, C.c_ssize_t]),
    "swe_download_frame": (C.c_int, [_E, C.c_void_p]),
    "swe_download_frame_begin": (C.c_int, [_E, C.c_void_p]),
    "swe_download_frame_wait": (C.c_int, [_E, C.c_int]),
    "swe_get_option": (C.c_int, [_E, C.c_int, C.POINTER(C.c_int)]),
    "swe_import_stats": (C.c_int, [_E, C.POINTER(Stats)]),
    "swe_upload_cells": (C.c_int, [_E, C.c_void_p]),
    "swe_download_cells": (C.c_int, [_E, C.c_void_p]),
    "swe_upload_rock": (C.c_int, [_E, C.c_void_p, C.c_ssize_t]),
    "swe_download_rock": (C.c_int, [_E, C.c_void_p, C.c_ssize_t]),
    "swe_clear_rock": (C.c_int, [_E]),
    "swe_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "swe_host_free": (C.c_int, [C.c_void_p]),
    "swe_erode": (C.c_int, [_E, C.c_int, C.c_int]),
    "swe_cascade": (C.c_int, [_E, C.c_int, C.c_int]),
    "swe_cascade_sweep": (C.c_int, [_E, C.c_int, _F]),
    "swe_cascade_converge": (C.c_int, [_E, C.c_float, C.c_int, C.POINTER(C.c_int), _F]),
    "swe_trace": (C.c_int, [_E, C.c_float, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_int),
                            C.POINTER(WindState)]),
    "swe_wind_init": (C.c_int, [C.POINTER(WindState), C.c_float, C.c_float]),
    "swe_wind_fly": (C.c_int, [_E, C.POINTER(WindState), C.POINTER(C.c_int)]),
    "swe_normal": (C.c_int, [_E, C.c_int, C.c_int, _F]),
    "swe_height": (C.c_int, [_E, C.c_int, C.c_int, _F]),
    "swe_download_normals": (C.c_int, [_E, C.c_void_p]),
    "swe_download_vertices": (C.c_int, [_E, C.c_void_p]),
    "swe_write_vertices_device": (C.c_int, [_E, C.c_void_p]),
    "swe_generate_noise": (C.c_int, [_E, C.c_int, C.c_int, C.c_int]),
    "swe_height_range": (C.c_int, [_E, _F, _F]),
    "swe_normalize_height": (C.c_int, [_E, C.c_float, C.c_float]),
    "swe_get_stats": (C.c_int, [_E, C.POINTER(Stats)]),
    "swe_reset_stats": (C.c_int, [_E]),
    "swe_sum_plane": (C.c_int, [_E, C.c_int, C.POINTER(C.c_double)]),
    "swe_max_excess": (C.c_int, [_E, _F]),
    "swe_hash_height": (C.c_int, [_E, C.POINTER(C.c_uint64)]),
    "swe_sync": (C.c_int, [_E]),
    "swe_set_stream": (C.c_int, [_E, C.c_void_p]),
    "swe_get_stream": (C.c_void_p, [_E]),
    "swe_device_height": (C.c_void_p, [_E]),
    "swe_device_field": (C.c_void_p, [_E]),
    "swe_device_track": (C.c_void_p, [_E]),
}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This package has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    got = lib.swe_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {got}, expected {ABI_VERSION}")
    return lib


lib = load()


def check(rc: int) -> None:
    if rc != 0:
        raise SweError(rc, lib.swe_last_error().decode("utf-8", "replace"))

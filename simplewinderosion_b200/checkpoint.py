"""Checkpoint / dump format. The reference has no persistence at all (its map
lives in one `new cell[area]`, cellpool.h:139-143, and dies with the process); this is the format a
long dune-formation run needs to stop and resume, and what the analysis tools read.

One file, little-endian:

    bytes 0..7    magic  b"SWEB200\\0"
    bytes 8..11   uint32 format version (1)
    bytes 12..15  uint32 header length L
    bytes 16..    L bytes of UTF-8 JSON: nx, nz, row0, global_nx, planes (names, in file order),
                  params (swe_params fields), state (hex of the SWE_STATE_BYTES run-state blob:
                  libc-rand replica, seed, erode call index), stats, has_rock, user
    padding       zeros up to the next multiple of 64
    planes        each nx*nz float32, index x*nz + z (math::flatten, source/include/math.h:11-12),
                  in `planes` order: the nine quad::cell members (cellpool.h:226-240), then "rock"
                  if the layered extension is on

`save` reads the planes through the C-ABI (swe_download_plane), `load` creates an engine and puts
them back, including the run state, so a deterministic run resumed from a checkpoint continues
bitwise identically to the uninterrupted one. A stripe engine saves/loads its owned rows (one file
per rank); `read_planes` gives the arrays to tools without touching a GPU."""
from __future__ import annotations

import ctypes as C
import json
import os
import struct

import numpy as np

from . import world as _world
from ._capi import STATE_BYTES, Params, Stats, check, lib

MAGIC = b"SWEB200\0"
VERSION = 1
PLANE_NAMES = ["height", "massflow", "momentumx", "momentumy", "momentumz",
               "massflow_track", "momentumx_track", "momentumy_track", "momentumz_track"]
# engine options that change what a parallel run computes (swe_option values)
_OPTIONS = {"cascade_interval": _world.OPT_CASCADE_INTERVAL, "call_batches": _world.OPT_CALL_BATCHES,
            "schedule": _world.OPT_SCHEDULE, "stagger_rounds": _world.OPT_STAGGER_ROUNDS,
            "tile_rounds": _world.OPT_TILE_ROUNDS, "tile_shape": _world.OPT_TILE_SHAPE}
_PARAM_FIELDS = ["max_age", "boundarylayer", "suspension", "gravity", "lrate", "maxdiff", "settling",
                 "mapscale", "abrasion"]


def _params_to_dict(p: Params) -> dict:
    d = {f: getattr(p, f) for f in _PARAM_FIELDS}
    d["pspeed"] = [float(v) for v in p.pspeed]
    return d


def _params_from_dict(d: dict) -> Params:
    p = _world.default_params()
    for f in _PARAM_FIELDS:
        if f in d:
            setattr(p, f, d[f])
    if "pspeed" in d:
        p.pspeed[:] = [float(v) for v in d["pspeed"]]
    return p


def save(w: "_world.World", path: str, user: dict | None = None, with_rock: bool | None = None) -> dict:
    """Write the engine's planes and run state to `path`; returns the header."""
    state = (C.c_ubyte * STATE_BYTES)()
    check(lib.swe_export_state(w._h, state, STATE_BYTES))
    planes = [w.download(k) for k in range(_world.NUM_PLANES)]
    names = list(PLANE_NAMES)
    if with_rock is None:
        try:
            rock = w.download_rock()
        except _world.SweError:
            rock = None
    else:
        rock = w.download_rock() if with_rock else None
    if rock is not None:
        planes.append(rock)
        names.append("rock")
    st = w.stats()
    header = {
        "nx": w.nx, "nz": w.nz, "row0": w.row0, "global_nx": w.global_nx, "planes": names,
        "params": _params_to_dict(w.params), "state": bytes(state).hex(), "has_rock": rock is not None,
        "stats": {"particles": st.particles, "steps": st.steps, "exits_oob": st.exits_oob,
                  "exits_age": st.exits_age, "sediment_lost": st.sediment_lost, "erode_calls": st.erode_calls,
                  "cascade_cells": st.cascade_cells},
        "options": {name: w.get_option(opt) for name, opt in _OPTIONS.items()},
        "user": user or {},
    }
    blob = json.dumps(header, sort_keys=True).encode("utf-8")
    # written beside the target and renamed into place: a crash mid-write never destroys the previous checkpoint
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", VERSION, len(blob)))
        f.write(blob)
        f.write(b"\0" * (-(16 + len(blob)) % 64))
        for a in planes:
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)
    return header


def read_header(path: str) -> tuple[dict, int]:
    """(header, byte offset of the first plane)."""
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) != 16 or head[:8] != MAGIC:
            raise ValueError(f"{path}: not a SWEB200 checkpoint")
        version, n = struct.unpack("<II", head[8:])
        if version != VERSION:
            raise ValueError(f"{path}: format version {version}, this reader knows {VERSION}")
        header = json.loads(f.read(n).decode("utf-8"))
    off = 16 + n
    return header, off + (-off % 64)


def read_planes(path: str) -> tuple[dict, dict]:
    """(header, {plane name: (nx, nz) float32 array}) without a GPU (memory-mapped)."""
    header, off = read_header(path)
    nx, nz = header["nx"], header["nz"]
    out = {}
    for i, name in enumerate(header["planes"]):
        out[name] = np.memmap(path, dtype="<f4", mode="r", offset=off + i * nx * nz * 4, shape=(nx, nz))
    return header, out


def load(path: str, device: int = 0, into: "_world.World | None" = None, sync_halos: bool = False) -> "_world.World":
    """Recreate the engine of a checkpoint (or refill `into`, e.g. a connected stripe engine): planes, rock,
    run state, the options that change what a parallel run computes, and the counters (so that the mass balance
    sum(height) + sediment_lost stays continuous across the resume). A stripe engine gets its owned rows only:
    its ghost rows must be refreshed before the next erode -- pass sync_halos=True when every rank loads
    collectively (one process per GPU), or call StripeGroup.sync_halos() for in-process stripes."""
    header, planes = read_planes(path)
    if into is None:
        if header["row0"] != 0 or header["global_nx"] != header["nx"]:
            raise ValueError("stripe checkpoint: create the StripedWorld, connect it and pass it as `into`")
        w = _world.World(header["nx"], header["nz"], params=_params_from_dict(header["params"]), device=device)
    else:
        w = into
        if (w.nx, w.nz, w.row0, w.global_nx) != (header["nx"], header["nz"], header["row0"], header["global_nx"]):
            raise ValueError("checkpoint geometry does not match the engine")
        w.params = _params_from_dict(header["params"])
    for k, name in enumerate(PLANE_NAMES):
        w.upload(np.ascontiguousarray(planes[name]), k)
    if header.get("has_rock"):
        w.upload_rock(np.ascontiguousarray(planes["rock"]))
    state = (C.c_ubyte * STATE_BYTES).from_buffer_copy(bytes.fromhex(header["state"]))
    check(lib.swe_import_state(w._h, state, STATE_BYTES))
    for name, value in header.get("options", {}).items():
        if name in _OPTIONS:
            w.set_option(_OPTIONS[name], int(value))
    hs = header.get("stats")
    if hs:
        st = Stats()
        for k, v in hs.items():
            setattr(st, k, v)
        check(lib.swe_import_stats(w._h, C.byref(st)))
    if sync_halos and hasattr(w, "sync_halos"):
        w.sync_halos()
    return w

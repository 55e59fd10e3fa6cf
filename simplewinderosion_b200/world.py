"""Host-side mirror of the reference interface for the erode cycle, over the C-ABI.

Names follow the reference: ``World.erode(cycles)`` / ``World.cascade(pos)`` /
``World.lrate, maxdiff, settling`` (source/world.hpp:9-32), ``World.map.height(p) /
normal(p) / oob(p)`` (source/cellpool.h:432-480), ``Wind(pos)``, ``Wind.fly()``,
``Wind.maxAge, boundarylayer, suspension, gravity`` (source/wind.hpp:4-26). The reference
keeps all of this in process-global statics; here a ``World`` instance owns one engine
(one GPU, or one stripe of a multi-GPU map). All map arithmetic happens in the CUDA library;
numpy is only the host container.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Params, Stats, SweError, WindState, check, lib

HEIGHT, MASSFLOW, MOMENTUMX, MOMENTUMY, MOMENTUMZ = 0, 1, 2, 3, 4
MASSFLOW_TRACK, MOMENTUMX_TRACK, MOMENTUMY_TRACK, MOMENTUMZ_TRACK = 5, 6, 7, 8
NUM_PLANES = 9
MODE_DETERMINISTIC, MODE_PARALLEL = 0, 1
OPT_CASCADE_VARIANT, OPT_CASCADE_INTERVAL, OPT_KERNEL_TIMING, OPT_CUDA_GRAPH, OPT_PDL, OPT_CALL_BATCHES = 1, 2, 3, 4, 5, 6
OPT_SCHEDULE, OPT_STAGGER_ROUNDS, OPT_PERSISTENT, OPT_PERIODIC, OPT_TILE_ROUNDS, OPT_TILE_SHAPE = 7, 8, 9, 10, 11, 12
SCHEDULE_COLOUR_QUEUES, SCHEDULE_LOCKSTEP, SCHEDULE_TILES = 0, 1, 2

# numpy view of the reference's 36-byte AoS record, quad::cell (cellpool.h:226-240)
CELL_DTYPE = np.dtype([(n, np.float32) for n in (
    "height", "massflow", "momentumx", "momentumy", "momentumz",
    "massflow_track", "momentumx_track", "momentumy_track", "momentumz_track")])

_PARAM_ALIASES = {  # reference spelling -> swe_params field
    "maxAge": "max_age", "boundarylayer": "boundarylayer", "suspension": "suspension",
    "gravity": "gravity", "lrate": "lrate", "maxdiff": "maxdiff", "settling": "settling",
    "mapscale": "mapscale", "abrasion": "abrasion",
}


def default_params() -> Params:
    p = Params()
    check(lib.swe_params_default(C.byref(p)))
    return p


class _Map:
    """quad::map accessors (cellpool.h:432-480) on the device-resident map."""

    def __init__(self, world: "World"):
        self._w = world

    def oob(self, p) -> bool:
        x, z = int(p[0]), int(p[1])
        return x < 0 or z < 0 or x >= self._w.global_nx or z >= self._w.nz

    def height(self, p) -> float:
        """map.height(p): 0.0 for out-of-bounds, like the reference (cellpool.h:472-476)."""
        if self.oob(p):
            return 0.0
        out = C.c_float()
        check(lib.swe_height(self._w._h, int(p[0]), int(p[1]), C.byref(out)))
        return out.value

    def normal(self, p) -> np.ndarray:
        out = (C.c_float * 3)()
        check(lib.swe_normal(self._w._h, int(p[0]), int(p[1]), out))
        return np.array(out[:], dtype=np.float32)


class World:
    """class World (world.hpp:9-32) with a run-time map size."""

    def __init__(self, nx: int, nz: int | None = None, params: Params | None = None, device: int = 0,
                 stripe: tuple[int, int] | None = None, max_particles: int = 0):
        """stripe=(own_lo, own_hi): this engine is one row stripe of an nx-by-nz map spread over
        several GPUs (see stripes.py); host transfers then cover the owned rows only and
        ``self.nx`` is the number of owned rows (``self.global_nx`` the map's)."""
        nz = nx if nz is None else nz
        self._h = C.c_void_p()
        p = params if params is not None else default_params()
        self.global_nx, self.nz, self.device = int(nx), int(nz), int(device)
        if stripe is None:
            check(lib.swe_create(C.byref(self._h), int(nx), int(nz), C.byref(p), int(device)))
            self.row0, self.nx = 0, int(nx)
        else:
            lo, hi = int(stripe[0]), int(stripe[1])
            check(lib.swe_stripe_create(C.byref(self._h), int(nx), int(nz), lo, hi, int(max_particles),
                                        C.byref(p), int(device)))
            self.row0, self.nx = lo, hi - lo
        self.stripe = stripe
        self.map = _Map(self)
        self.SEED = 0

    # -- lifetime --------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.swe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- tunables (World::lrate = ..., Wind::maxAge = ...) -----------------------------------
    @property
    def params(self) -> Params:
        p = Params()
        check(lib.swe_get_params(self._h, C.byref(p)))
        return p

    @params.setter
    def params(self, p: Params) -> None:
        check(lib.swe_set_params(self._h, C.byref(p)))

    def set(self, **kw) -> None:
        p = self.params
        for k, v in kw.items():
            f = _PARAM_ALIASES.get(k, k)
            if f == "pspeed":
                p.pspeed[:] = [float(x) for x in v]
            elif hasattr(p, f):
                setattr(p, f, v)
            else:
                raise AttributeError(k)
        self.params = p

    def __getattr__(self, name):  # world.lrate, world.maxdiff, world.settling ...
        f = _PARAM_ALIASES.get(name)
        if f is None:
            raise AttributeError(name)
        return getattr(self.params, f)

    def set_option(self, option: int, value: int) -> None:
        check(lib.swe_set_option(self._h, int(option), int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int()
        check(lib.swe_get_option(self._h, int(option), C.byref(v)))
        return v.value

    def srand(self, seed: int) -> None:
        """srand(SEED) (SimpleWindErosion.cpp:27-34)."""
        self.SEED = int(seed)
        check(lib.swe_srand(self._h, int(seed) & 0xFFFFFFFF))

    # -- map transfer --------------------------------------------------------------------------
    def upload(self, height: np.ndarray, plane: int = HEIGHT) -> None:
        a = np.ascontiguousarray(height, dtype=np.float32)
        if a.size != self.nx * self.nz:
            raise ValueError(f"expected {self.nx}x{self.nz} values, got {a.shape}")
        check(lib.swe_upload_plane(self._h, plane, a.ctypes.data, 4))

    def download(self, plane: int = HEIGHT, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.nx, self.nz), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == self.nx * self.nz
        check(lib.swe_download_plane(self._h, plane, out.ctypes.data, 4))
        return out

    def download_frame(self, out: np.ndarray | None = None) -> np.ndarray:
        """What the reference's frame loop reads after World::erode (SimpleWindErosion.cpp:219-230): the (5, nx, nz)
        stack height, massflow, momentumx, momentumy, momentumz in one call and one synchronisation."""
        if out is None:
            out = np.empty((5, self.nx, self.nz), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == 5 * self.nx * self.nz
        check(lib.swe_download_frame(self._h, out.ctypes.data))
        return out

    def download_frame_begin(self, out: np.ndarray) -> None:
        """Pipelined form of download_frame: returns at once; the height plane arrives on the engine's stream, the four
        flow-field planes on a side stream while the next erode runs. `out` (5, nx, nz) should be pinned memory and
        must stay alive until download_frame_wait(True)."""
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == 5 * self.nx * self.nz
        check(lib.swe_download_frame_begin(self._h, out.ctypes.data))

    def download_frame_wait(self, all_planes: bool = True) -> None:
        """Block until the height plane (all_planes=False) or all five planes of the frame in flight have arrived."""
        check(lib.swe_download_frame_wait(self._h, 1 if all_planes else 0))

    # layered-terrain extension (bedrock under the sand; not in the reference snapshot)
    def upload_rock(self, rock: np.ndarray) -> None:
        a = np.ascontiguousarray(rock, dtype=np.float32)
        if a.size != self.nx * self.nz:
            raise ValueError(f"expected {self.nx}x{self.nz} values, got {a.shape}")
        check(lib.swe_upload_rock(self._h, a.ctypes.data, 4))

    def download_rock(self) -> np.ndarray:
        out = np.empty((self.nx, self.nz), dtype=np.float32)
        check(lib.swe_download_rock(self._h, out.ctypes.data, 4))
        return out

    def clear_rock(self) -> None:
        check(lib.swe_clear_rock(self._h))

    def upload_cells(self, cells: np.ndarray) -> None:
        """cells: structured array of CELL_DTYPE (the reference's cellpool memory)."""
        assert cells.dtype == CELL_DTYPE and cells.flags.c_contiguous and cells.size == self.nx * self.nz
        check(lib.swe_upload_cells(self._h, cells.ctypes.data))

    def download_cells(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.nx, self.nz), dtype=CELL_DTYPE)
        assert out.dtype == CELL_DTYPE and out.flags.c_contiguous
        check(lib.swe_download_cells(self._h, out.ctypes.data))
        return out

    def upload_strided(self, plane: int, base_ptr: int, stride_bytes: int) -> None:
        check(lib.swe_upload_plane(self._h, plane, base_ptr, stride_bytes))

    def download_strided(self, plane: int, base_ptr: int, stride_bytes: int) -> None:
        check(lib.swe_download_plane(self._h, plane, base_ptr, stride_bytes))

    def normals(self) -> np.ndarray:
        out = np.empty((self.nx, self.nz, 3), dtype=np.float32)
        check(lib.swe_download_normals(self._h, out.ctypes.data))
        return out

    def vertices(self) -> np.ndarray:
        """The (nx, nz, 12) Vertex records updatenode writes each frame (cellpool.h:302-321):
        position, normal, tangent, bitangent."""
        out = np.empty((self.nx, self.nz, 12), dtype=np.float32)
        check(lib.swe_download_vertices(self._h, out.ctypes.data))
        return out

    def write_vertices_device(self, device_ptr: int) -> None:
        """Same records straight into device memory (a CUDA-mapped vertex buffer), asynchronously."""
        check(lib.swe_write_vertices_device(self._h, C.c_void_p(device_ptr)))

    # -- generate (quad::map::init on the device) -------------------------------------------------
    def generate_noise(self, seed: int, layers: int = 8, tile: int = 0) -> None:
        """height += the FastNoiseLite OpenSimplex2-FBm layers of map::init (cellpool.h:388-411);
        bitwise identical to the host library. tile = coordinate divisor (quad::tilesize), 0 = nz."""
        check(lib.swe_generate_noise(self._h, int(seed), int(layers), int(tile)))

    def height_range(self) -> tuple[float, float]:
        """(min, max) folded from 0.0f like cellpool.h:416-423."""
        lo, hi = C.c_float(), C.c_float()
        check(lib.swe_height_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def normalize(self, lo: float | None = None, hi: float | None = None) -> None:
        """height = 0.5*((height - min)/(max - min)) (cellpool.h:424-428)."""
        if lo is None or hi is None:
            lo, hi = self.height_range()
        check(lib.swe_normalize_height(self._h, float(lo), float(hi)))

    # -- hot path --------------------------------------------------------------------------------
    def erode(self, cycles: int, mode: int = MODE_PARALLEL) -> None:
        """World::erode(cycles) (world.hpp:36-77)."""
        check(lib.swe_erode(self._h, int(cycles), int(mode)))

    def cascade(self, pos) -> None:
        """World::cascade(vec2 pos) (world.hpp:79-153); pos truncates to ints (world.hpp:103)."""
        check(lib.swe_cascade(self._h, int(pos[0]), int(pos[1])))

    def cascade_sweep(self, sweeps: int = 1, want_excess: bool = False):
        ex = C.c_float()
        check(lib.swe_cascade_sweep(self._h, int(sweeps), C.byref(ex) if want_excess else None))
        return ex.value if want_excess else None

    def cascade_converge(self, tol: float, max_sweeps: int) -> tuple[int, float]:
        done, ex = C.c_int(), C.c_float()
        check(lib.swe_cascade_converge(self._h, float(tol), int(max_sweeps), C.byref(done), C.byref(ex)))
        return done.value, ex.value

    def trace(self, x: float, z: float, max_records: int = 600):
        """Wind(vec2(x,z)); while(wind.fly()) ... recording every completed step."""
        rec = np.zeros((max_records, 8), dtype=np.float32)
        steps, fin = C.c_int(), WindState()
        check(lib.swe_trace(self._h, float(x), float(z), int(max_records), rec.ctypes.data,
                            C.byref(steps), C.byref(fin)))
        return rec[:min(steps.value, max_records)], steps.value, fin

    # -- diagnostics ------------------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        check(lib.swe_get_stats(self._h, C.byref(s)))
        return s

    def reset_stats(self) -> None:
        check(lib.swe_reset_stats(self._h))

    def sum_plane(self, plane: int = HEIGHT) -> float:
        out = C.c_double()
        check(lib.swe_sum_plane(self._h, plane, C.byref(out)))
        return out.value

    def max_excess(self) -> float:
        out = C.c_float()
        check(lib.swe_max_excess(self._h, C.byref(out)))
        return out.value

    def hash_height(self) -> int:
        out = C.c_uint64()
        check(lib.swe_hash_height(self._h, C.byref(out)))
        return out.value

    def sync(self) -> None:
        check(lib.swe_sync(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        check(lib.swe_set_stream(self._h, cuda_stream))


class Wind:
    """struct Wind (wind.hpp:4-21): ``w = Wind(world, (x, z)); while w.fly(): ...``"""

    def __init__(self, world: World, pos):
        self._world = world
        self._s = WindState()
        check(lib.swe_wind_init(C.byref(self._s), float(pos[0]), float(pos[1])))

    age = property(lambda self: self._s.age)
    sediment = property(lambda self: self._s.sediment)
    pos = property(lambda self: np.array(self._s.pos[:], dtype=np.float32))
    speed = property(lambda self: np.array(self._s.speed[:], dtype=np.float32))

    def fly(self) -> bool:
        alive = C.c_int()
        check(lib.swe_wind_fly(self._world._h, C.byref(self._s), C.byref(alive)))
        return bool(alive.value)


__all__ = [
    "World", "Wind", "Params", "Stats", "WindState", "SweError", "default_params", "CELL_DTYPE",
    "HEIGHT", "MASSFLOW", "MOMENTUMX", "MOMENTUMY", "MOMENTUMZ", "MASSFLOW_TRACK",
    "MOMENTUMX_TRACK", "MOMENTUMY_TRACK", "MOMENTUMZ_TRACK", "NUM_PLANES",
    "MODE_DETERMINISTIC", "MODE_PARALLEL", "OPT_CASCADE_VARIANT", "OPT_CASCADE_INTERVAL", "OPT_KERNEL_TIMING", "OPT_CUDA_GRAPH", "OPT_PDL", "OPT_CALL_BATCHES",
    "OPT_SCHEDULE", "OPT_STAGGER_ROUNDS", "OPT_PERSISTENT", "OPT_PERIODIC", "SCHEDULE_COLOUR_QUEUES", "SCHEDULE_LOCKSTEP", "SCHEDULE_TILES", "OPT_TILE_ROUNDS", "OPT_TILE_SHAPE",
]

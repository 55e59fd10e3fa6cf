"""TEST INFRASTRUCTURE: ctypes access to the checker --
   PortWorld : oracle/_build/libswe_oracle.so  (our plain-C restatement, oracle/swe_oracle.c)
   RefWorld  : oracle/_ref/libswe_ref_<N>.so   (the UNMODIFIED reference headers, compiled
               headless by oracle/Makefile; one library per map edge, one map per process)
Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_LIB = os.path.join(ROOT, "oracle", "_build", "libswe_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class OParams(C.Structure):
    _fields_ = [("max_age", C.c_int), ("boundarylayer", C.c_float), ("suspension", C.c_float),
                ("gravity", C.c_float), ("lrate", C.c_float), ("maxdiff", C.c_float),
                ("settling", C.c_float), ("mapscale", C.c_int), ("pspeed", C.c_float * 3)]


class ORand(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]


class OWorld(C.Structure):
    _fields_ = [("nx", C.c_int), ("nz", C.c_int), ("plane", C.POINTER(C.c_float) * 9),
                ("p", OParams), ("rng", ORand), ("rock", C.POINTER(C.c_float)), ("periodic", C.c_int)]


class OStats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("exits_oob", C.c_longlong), ("exits_age", C.c_longlong),
                ("sediment_lost", C.c_double)]


class OParticle(C.Structure):
    _fields_ = [("age", C.c_int), ("sediment", C.c_float), ("pos", C.c_float * 3),
                ("speed", C.c_float * 3)]


_port = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_LIB):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        L = C.CDLL(PORT_LIB)
        W = C.POINTER(OWorld)
        L.swo_create.restype = W
        L.swo_create.argtypes = [C.c_int, C.c_int]
        L.swo_destroy.argtypes = [W]
        L.swo_set_rock.argtypes = [W, C.c_void_p]
        L.swo_srand.argtypes = [C.POINTER(ORand), C.c_uint]
        L.swo_rand.argtypes = [C.POINTER(ORand)]
        L.swo_rand.restype = C.c_int
        L.swo_expf.argtypes = [C.c_float]
        L.swo_expf.restype = C.c_float
        L.swo_init_gaussian.argtypes = [W]
        L.swo_hill_raw.argtypes = [W]
        L.swo_normal.argtypes = [W, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.swo_fly.argtypes = [W, C.POINTER(OParticle)]
        L.swo_fly.restype = C.c_int
        L.swo_particle_at.argtypes = [C.POINTER(OParticle), C.c_float, C.c_float]
        L.swo_cascade.argtypes = [W, C.c_int, C.c_int]
        L.swo_erode.argtypes = [W, C.c_int, C.POINTER(OStats)]
        L.swo_cascade_sweep.argtypes = [W, C.c_int]
        L.swo_max_excess.argtypes = [W]
        L.swo_max_excess.restype = C.c_float
        L.swo_trace.argtypes = [W, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.swo_trace.restype = C.c_int
        L.swo_hash_height.argtypes = [W]
        L.swo_hash_height.restype = C.c_uint64
        L.swo_fnl_single3.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
        L.swo_fnl_single3.restype = C.c_float
        L.swo_fnl_getnoise.argtypes = [C.c_float] * 4
        L.swo_fnl_getnoise.restype = C.c_float
        L.swo_noise_stack.argtypes = [W, C.c_int, C.c_int, C.c_int]
        L.swo_normalize.argtypes = [W]
        L.swo_vertices.argtypes = [W, C.c_void_p]
        L.swo_noise_rows.argtypes = [C.c_void_p] + [C.c_int] * 6
        _port = L
    return _port


class PortWorld:
    """The plain-C restatement (oracle/swe_oracle.c)."""

    def __init__(self, nx: int, nz: int | None = None, periodic: bool = False):
        nz = nx if nz is None else nz
        self.L = port()
        self.w = self.L.swo_create(nx, nz)
        assert self.w, "swo_create failed"
        self.nx, self.nz = nx, nz
        self.w.contents.periodic = int(periodic)

    def __del__(self):
        try:
            self.L.swo_destroy(self.w)
        except Exception:
            pass

    @property
    def params(self) -> OParams:
        return self.w.contents.p

    def set(self, **kw):
        for k, v in kw.items():
            k = {"maxAge": "max_age"}.get(k, k)
            setattr(self.w.contents.p, k, v)

    def plane(self, k: int) -> np.ndarray:
        """numpy VIEW of plane k (0=height ... 8=momentumz_track); the view keeps this
        world alive."""
        addr = C.addressof(self.w.contents.plane[k].contents)
        buf = (C.c_float * (self.nx * self.nz)).from_address(addr)
        buf._owner = self
        return np.frombuffer(buf, dtype=np.float32).reshape(self.nx, self.nz)

    def set_plane(self, k: int, a) -> None:
        self.plane(k)[...] = np.asarray(a, dtype=np.float32).reshape(self.nx, self.nz)

    def set_rock(self, rock) -> None:
        """Layered extension (our own specification, parity unpinned): bedrock plane or None."""
        if rock is None:
            self.L.swo_set_rock(self.w, None)
        else:
            a = np.ascontiguousarray(rock, dtype=np.float32).reshape(self.nx, self.nz)
            self.L.swo_set_rock(self.w, a.ctypes.data)

    def srand(self, seed: int) -> None:
        self.L.swo_srand(C.byref(self.w.contents.rng), seed)

    def rand(self) -> int:
        return self.L.swo_rand(C.byref(self.w.contents.rng))

    def init_gaussian(self) -> None:
        self.L.swo_init_gaussian(self.w)

    def hill_raw(self) -> None:
        """the Gaussian hill before normalisation (cellpool.h:378-385)"""
        self.L.swo_hill_raw(self.w)

    def erode(self, cycles: int) -> OStats:
        st = OStats()
        self.L.swo_erode(self.w, cycles, C.byref(st))
        return st

    def cascade(self, x: int, z: int) -> None:
        self.L.swo_cascade(self.w, x, z)

    def cascade_sweep(self, sweeps: int = 1) -> None:
        self.L.swo_cascade_sweep(self.w, sweeps)

    def max_excess(self) -> float:
        return self.L.swo_max_excess(self.w)

    def normal(self, x: int, z: int) -> np.ndarray:
        out = (C.c_float * 3)()
        self.L.swo_normal(self.w, x, z, out)
        return np.array(out[:], dtype=np.float32)

    def trace(self, x: float, z: float, max_records: int = 600):
        rec = np.zeros((max_records, 8), dtype=np.float32)
        n = self.L.swo_trace(self.w, x, z, max_records, rec.ctypes.data)
        return rec[:min(n, max_records)], n

    def hash_height(self) -> int:
        return self.L.swo_hash_height(self.w)

    def vertices(self) -> np.ndarray:
        """(nx, nz, 12): the Vertex records of updatenode (cellpool.h:302-321)."""
        out = np.empty((self.nx, self.nz, 12), dtype=np.float32)
        self.L.swo_vertices(self.w, out.ctypes.data)
        return out

    def noise_stack(self, seed: int, layers: int = 8, tile: int | None = None) -> None:
        """height += the FastNoiseLite layer stack of cellpool.h:392-409."""
        self.L.swo_noise_stack(self.w, seed, layers, self.nz if tile is None else tile)

    def normalize(self) -> None:
        self.L.swo_normalize(self.w)

    def get_noise(self, frequency: float, x: float, y: float, z: float) -> float:
        return self.L.swo_fnl_getnoise(frequency, x, y, z)


def ref_available(n: int, periodic: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libswe_ref_{'p' if periodic else ''}{n}.so"))


class RefWorld:
    """The unmodified reference (oracle/_ref). ONE instance per size per process: the
    reference keeps its map in static globals (world.hpp:12-19)."""
    _libs: dict = {}

    def __init__(self, n: int, seed: int = 1, periodic: bool = False):
        """periodic: the build with the reference's commented periodic-boundary blocks swapped in
        (cellpool.h:96-113,447-465; oracle/Makefile)."""
        key = ("p" if periodic else "") + str(n)
        if key not in RefWorld._libs:
            L = C.CDLL(os.path.join(REF_DIR, f"libswe_ref_{key}.so"))
            L.ref_hash.restype = C.c_ulonglong
            L.ref_sum_height.restype = C.c_double
            L.ref_set_plane.argtypes = [C.c_int, C.c_void_p]
            L.ref_get_plane.argtypes = [C.c_int, C.c_void_p]
            L.ref_erode_counted.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
            L.ref_trace.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p]
            L.ref_trace.restype = C.c_int
            L.ref_cascade.argtypes = [C.c_float, C.c_float]
            L.ref_normal.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
            L.ref_set_params.argtypes = [C.c_int] + [C.c_float] * 6
            if hasattr(L, "ref_get_noise"):
                L.ref_get_noise.argtypes = [C.c_float] * 4
                L.ref_get_noise.restype = C.c_float
                L.ref_noise_stack.argtypes = [C.c_int, C.c_int]
            RefWorld._libs[key] = L
        self.L = RefWorld._libs[key]
        assert self.L.ref_size() == n
        self.n = self.nx = self.nz = n
        self.L.ref_init(seed)  # srand(seed) + Gaussian-hill map (reset if already initialised)
        self.L.ref_set_params(512, 2.0, 0.05, 0.1, 0.1, 0.005, 1.0)

    def plane(self, k: int) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float32)
        self.L.ref_get_plane(k, out.ctypes.data)
        return out

    def set_plane(self, k: int, a) -> None:
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(self.n, self.n)
        self.L.ref_set_plane(k, a.ctypes.data)

    def set(self, maxAge=512, boundarylayer=2.0, suspension=0.05, gravity=0.1, lrate=0.1,
            maxdiff=0.005, settling=1.0):
        self.L.ref_set_params(maxAge, boundarylayer, suspension, gravity, lrate, maxdiff, settling)

    def srand(self, seed: int) -> None:
        self.L.ref_srand(seed)

    def rand(self) -> int:
        return self.L.ref_rand()

    def erode(self, cycles: int) -> None:
        self.L.ref_erode(cycles)

    def erode_counted(self, cycles: int):
        st = (C.c_longlong * 3)()
        lost = C.c_double()
        self.L.ref_erode_counted(cycles, st, C.byref(lost))
        return {"steps": st[0], "exits_oob": st[1], "exits_age": st[2], "sediment_lost": lost.value}

    def cascade(self, x: int, z: int) -> None:
        self.L.ref_cascade(float(x), float(z))

    def cascade_sweep(self, sweeps: int = 1) -> None:
        self.L.ref_cascade_sweep(sweeps)

    def normal(self, x: int, z: int) -> np.ndarray:
        out = (C.c_float * 3)()
        self.L.ref_normal(x, z, out)
        return np.array(out[:], dtype=np.float32)

    def trace(self, x: float, z: float, max_records: int = 600):
        rec = np.zeros((max_records, 8), dtype=np.float32)
        n = self.L.ref_trace(x, z, max_records, rec.ctypes.data)
        return rec[:min(n, max_records)], n

    def hash_height(self) -> int:
        return self.L.ref_hash()

    def sum_height(self) -> float:
        return self.L.ref_sum_height()

    def vertices(self) -> np.ndarray:
        """quad::updatenode run by the reference itself; the harness' Vertexpool stub records it."""
        out = np.empty((self.n, self.n, 12), dtype=np.float32)
        self.L.ref_updatenode.argtypes = [C.c_void_p]
        self.L.ref_updatenode(out.ctypes.data)
        return out

    def noise_stack(self, seed: int, layers: int = 8) -> None:
        self.L.ref_noise_stack(seed, layers)

    def normalize(self) -> None:
        self.L.ref_normalize()

    def get_noise(self, frequency: float, x: float, y: float, z: float) -> float:
        return self.L.ref_get_noise(frequency, x, y, z)


def _noise_band(args):
    x_lo, x_hi, nz, seed, layers, tile = args
    band = np.zeros((x_hi - x_lo, nz), dtype=np.float32)
    port().swo_noise_rows(band.ctypes.data, x_lo, x_hi, nz, seed, layers, tile)
    return band


def noise_map(nx: int, nz: int | None = None, seed: int = 1, layers: int = 8, workers: int = 1) -> np.ndarray:
    """The map::init noise layers on a zero map, normalised (cellpool.h:388-428), computed by the
    oracle port on `workers` host threads (row bands are independent)."""
    nz = nx if nz is None else nz
    edges = np.linspace(0, nx, max(1, min(workers, nx)) * 4 + 1).astype(int)
    jobs = [(int(a), int(b), nz, seed, layers, nz) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    if workers <= 1:
        bands = [_noise_band(j) for j in jobs]
    else:
        from concurrent.futures import ThreadPoolExecutor   # ctypes drops the GIL inside the C call
        port()
        with ThreadPoolExecutor(workers) as pool:
            bands = list(pool.map(_noise_band, jobs))
    h = np.concatenate(bands, axis=0)
    mn, mx = np.float32(min(0.0, h.min())), np.float32(max(0.0, h.max()))
    return (np.float32(0.5) * ((h - mn) / (mx - mn))).astype(np.float32)


def fnv1a64_f32(a: np.ndarray) -> int:
    """FNV-1a-64 over fp32 bit patterns, vectorised-free reference."""
    h = 1469598103934665603
    for w in np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel().tolist():
        h = ((h ^ w) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def mt_uniform_map(n: int, seed: int = 42, nz: int | None = None) -> np.ndarray:
    """height = U(0,1) map for cascade tests. (The survey's probe used std::mt19937 +
    uniform_real_distribution<float>; numpy's MT19937 float path differs, so fixtures that
    must match the probe hash are generated by oracle/_ref tooling instead.)"""
    nz = n if nz is None else nz
    rs = np.random.RandomState(seed)
    return rs.random_sample((n, nz)).astype(np.float32)

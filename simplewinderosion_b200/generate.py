"""Host-side map generation (stays on the host, like the reference's quad::map::init,
source/cellpool.h:339-430). Nothing here is on the accelerated path.

* ``gaussian_hill``  -- the reference's live initial condition (cellpool.h:378-385,416-428)
* ``fbm``            -- 8-octave fractal noise stack in the spirit of the reference's
                        commented-out block (cellpool.h:388-411: octave o has frequency 2^o and
                        amplitude 0.6^(o+1)), normalised to [0, 0.5] like cellpool.h:416-428.
                        The lattice noise is our own smooth value noise (the reference's
                        FastNoiseLite/libnoise generators are third-party and not reproduced),
                        so these maps are synthetic inputs of the right character, not a
                        bit-parity target.
* ``flat_sand``      -- 0.25 + 1e-3 * U(-1,1): the flat-bed dune-formation input
* ``random_slopes``  -- U(0,1): the cascade stress input
"""
from __future__ import annotations

import numpy as np


def normalise_half(h: np.ndarray) -> np.ndarray:
    """cellpool.h:416-428: min/max start at 0.0, height = 0.5*(h-min)/(max-min)."""
    mn = min(0.0, float(h.min()))
    mx = max(0.0, float(h.max()))
    return (0.5 * ((h - mn) / (mx - mn))).astype(np.float32)


def gaussian_hill(n: int) -> np.ndarray:
    """exp(-d^2 * n * 0.2) around (n/4, n/2) in tile units, then normalised to [0, 0.5]."""
    x = (np.arange(n, dtype=np.float32) / np.float32(n))[:, None]
    z = (np.arange(n, dtype=np.float32) / np.float32(n))[None, :]
    cx, cz = np.float32(n // 4) / np.float32(n), np.float32(n // 2) / np.float32(n)
    d = np.sqrt((x - cx) ** 2 + (z - cz) ** 2, dtype=np.float32)
    h = np.exp((-d * d * np.float32(n)).astype(np.float64) * 0.2).astype(np.float32)
    return normalise_half(h)


def _value_noise(nx: int, nz: int, cells: int, rs: np.random.RandomState) -> np.ndarray:
    """Smooth (quintic-interpolated) lattice noise in [-1, 1] with `cells` lattice cells
    across the longer map edge."""
    span = max(nx, nz)
    gx = int(np.ceil(nx * cells / span)) + 2
    gz = int(np.ceil(nz * cells / span)) + 2
    lat = rs.uniform(-1.0, 1.0, size=(gx, gz)).astype(np.float32)
    fx = np.arange(nx, dtype=np.float32) * (cells / span)
    fz = np.arange(nz, dtype=np.float32) * (cells / span)
    ix, iz = fx.astype(np.int64), fz.astype(np.int64)
    tx, tz = fx - ix, fz - iz
    tx = tx * tx * tx * (tx * (tx * 6 - 15) + 10)
    tz = tz * tz * tz * (tz * (tz * 6 - 15) + 10)
    a = lat[ix][:, iz]; b = lat[ix + 1][:, iz]
    c = lat[ix][:, iz + 1]; d = lat[ix + 1][:, iz + 1]
    tx = tx[:, None]; tz = tz[None, :]
    top = a + (b - a) * tx
    bot = c + (d - c) * tx
    return (top + (bot - top) * tz).astype(np.float32)


def fbm(nx: int, nz: int | None = None, seed: int = 1, octaves: int = 8, gain: float = 0.6,
        base_cells: int = 1) -> np.ndarray:
    nz = nx if nz is None else nz
    rs = np.random.RandomState(seed)
    h = np.zeros((nx, nz), dtype=np.float32)
    amp, cells = gain, base_cells
    for _ in range(octaves):
        h += np.float32(amp) * _value_noise(nx, nz, cells, rs)
        amp *= gain
        cells *= 2
    return normalise_half(h)


def flat_sand(nx: int, nz: int | None = None, seed: int = 42, level: float = 0.25,
              noise: float = 1e-3) -> np.ndarray:
    nz = nx if nz is None else nz
    rs = np.random.RandomState(seed)
    return (level + noise * (2.0 * rs.random_sample((nx, nz)) - 1.0)).astype(np.float32)


def random_slopes(nx: int, nz: int | None = None, seed: int = 42) -> np.ndarray:
    nz = nx if nz is None else nz
    rs = np.random.RandomState(seed)
    return rs.random_sample((nx, nz)).astype(np.float32)


def pyramid(nx: int, nz: int | None = None, apex: float = 0.5) -> np.ndarray:
    """Square-based rock pyramid, base n/4 centred on the map (the obstacle of the
    4096^2 configuration; an engine extension -- the reference snapshot has no rock layer)."""
    nz = nx if nz is None else nz
    x = np.abs(np.arange(nx, dtype=np.float32) - nx / 2)[:, None]
    z = np.abs(np.arange(nz, dtype=np.float32) - nz / 2)[None, :]
    r = np.maximum(x / (nx / 8), z / (nz / 8))
    return np.maximum(0.0, apex * (1.0 - r)).astype(np.float32)

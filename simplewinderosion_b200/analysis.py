"""Host-side analysis helpers for the statistical parity checks of parallel mode (numpy only;
not on the accelerated path): height histogram and dune-wavelength spectrum."""
from __future__ import annotations

import numpy as np


def height_histogram(h: np.ndarray, bins: int = 64, lo: float | None = None, hi: float | None = None):
    """Normalised histogram (sums to 1) of the height plane over [lo, hi]."""
    h = np.asarray(h, dtype=np.float64)
    lo = float(h.min()) if lo is None else lo
    hi = float(h.max()) if hi is None else hi
    hist, edges = np.histogram(h, bins=bins, range=(lo, hi))
    return hist / max(1, hist.sum()), edges


def histogram_distance(h_a: np.ndarray, h_b: np.ndarray, bins: int = 64) -> float:
    """Total-variation distance (0..1) between the two height distributions on a common range."""
    lo = float(min(h_a.min(), h_b.min()))
    hi = float(max(h_a.max(), h_b.max()))
    pa, _ = height_histogram(h_a, bins, lo, hi)
    pb, _ = height_histogram(h_b, bins, lo, hi)
    return 0.5 * float(np.abs(pa - pb).sum())


def wind_spectrum(h: np.ndarray) -> np.ndarray:
    """Power spectrum along the wind direction (x = first index), averaged over z; the mean of
    every column is removed first. Index k = cycles per map length. Dunes show as a peak."""
    h = np.asarray(h, dtype=np.float64)
    d = h - h.mean(axis=0, keepdims=True)
    f = np.fft.rfft(d, axis=0)
    return (np.abs(f) ** 2).mean(axis=1) / h.shape[0]


def detrend_along_wind(h: np.ndarray, degree: int = 3) -> np.ndarray:
    """Remove the large-scale along-wind trend (a polynomial fitted to the z-averaged profile): on a closed
    map sand piles up / drains along x, which otherwise puts all spectral power into the lowest bins."""
    h = np.asarray(h, dtype=np.float64)
    x = np.arange(h.shape[0], dtype=np.float64)
    p = np.polyfit(x, h.mean(axis=1), degree)
    return h - np.polyval(p, x)[:, None]


def dominant_wavelength(h: np.ndarray, kmin: int = 4, detrend: bool = True) -> float:
    """Wavelength (cells) of the strongest spectral peak along the wind. The profile is detrended first and the
    bins k < kmin (map-scale structure, not dunes) are ignored -- without that the answer is just n/2 or n/3."""
    p = wind_spectrum(detrend_along_wind(h) if detrend else h)
    kmin = max(1, min(int(kmin), len(p) - 1))
    k = kmin + int(np.argmax(p[kmin:]))
    return h.shape[0] / k


def spectrum_similarity(h_a: np.ndarray, h_b: np.ndarray, kmin: int = 1, kmax: int | None = None) -> float:
    """Correlation of log power spectra over k in [kmin, kmax]."""
    pa, pb = wind_spectrum(h_a), wind_spectrum(h_b)
    kmax = min(len(pa), len(pb)) - 1 if kmax is None else kmax
    a = np.log(pa[kmin:kmax + 1] + 1e-30)
    b = np.log(pb[kmin:kmax + 1] + 1e-30)
    return float(np.corrcoef(a, b)[0, 1])

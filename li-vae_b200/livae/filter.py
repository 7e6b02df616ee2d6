"""livae.filter -- FFT pre-processing of whole micrographs (reference src/livae/filter.py).

One-shot per image at dataset construction, float64, off the training hot path:
kept on the host with scipy's FFT so the cached images are the reference's to the last bit
(tests/test_dropin_cpu.py compares against vectors produced by the reference's own filter.py).
The three radial filters share one helper: keep the centred-spectrum annulus lo <= r <= hi.
"""
from __future__ import annotations

import numpy as np
from scipy import fft as _fft

__all__ = ["fft_spectra", "normalize_image", "lowpass_filter", "highpass_filter", "bandpass_filter"]


def _as_image(image) -> np.ndarray:
    a = np.asarray(image)
    if a.ndim != 2:
        raise ValueError(f"Expected a 2D array, got shape {a.shape}")
    return a.astype(np.float64, copy=False)


def _centred_spectrum(a: np.ndarray) -> np.ndarray:
    return _fft.fftshift(_fft.fft2(a))


def _annulus_filter(image, lo: float | None, hi: float | None) -> np.ndarray:
    """real(ifft2(spectrum * [lo <= r <= hi])), r = distance from the centre bin (rows//2, cols//2)
    (filter.py:25-39, 147-232)"""
    a = _as_image(image)
    rows, cols = a.shape
    yy, xx = np.ogrid[:rows, :cols]
    r = np.sqrt((xx - cols // 2) ** 2 + (yy - rows // 2) ** 2)
    keep = np.ones(a.shape, dtype=bool) if lo is None else (r >= lo)
    if hi is not None:
        keep &= r <= hi
    return np.real(_fft.ifft2(_fft.ifftshift(_centred_spectrum(a) * keep)))


def fft_spectra(image):
    """-> (magnitude, phase) of the centred 2-D FFT (filter.py:42-74)"""
    f = _centred_spectrum(_as_image(image))
    return np.abs(f), np.angle(f)


def normalize_image(image) -> np.ndarray:
    """min-max to [0,1]; a constant image maps to zeros (filter.py:77-108)"""
    a = np.asarray(image, dtype=np.float64)
    lo = float(a.min())
    span = float(np.ptp(a))
    if span == 0.0:
        return np.zeros_like(a)
    return (a - lo) / span


def lowpass_filter(image, cutoff_radius: float) -> np.ndarray:
    return _annulus_filter(image, None, cutoff_radius)


def highpass_filter(image, cutoff_radius: float) -> np.ndarray:
    return _annulus_filter(image, cutoff_radius, None)


def bandpass_filter(image, low_cutoff: float, high_cutoff: float) -> np.ndarray:
    if high_cutoff <= low_cutoff:
        raise ValueError("high_cutoff must be greater than low_cutoff")
    return _annulus_filter(image, low_cutoff, high_cutoff)

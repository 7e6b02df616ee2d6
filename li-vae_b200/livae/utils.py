"""livae.utils -- lattice constant, HDF5 loading, checkpoint key clean-up (reference src/livae/utils.py).

Host-side, one-shot helpers the reference's scripts import (`from livae.utils import load_image_from_h5`,
scripts/train_rvae.py:24, train_vae.py:24, pretrain_stn.py:16).  h5py is imported lazily so that
`import livae` works where it is not installed (it is absent from the GPU image); calling
load_image_from_h5 without it raises ImportError with that explanation.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
from scipy.ndimage import gaussian_filter
from scipy.signal import find_peaks

from .filter import fft_spectra

__all__ = ["estimate_lattice_constant", "load_image_from_h5", "clean_state_dict"]


def estimate_lattice_constant(image, min_atom_size: float = 10.0, max_atom_size: float = 60.0,
                              prominence_factor: float = 0.1) -> float:
    """Lattice spacing in pixels from the first prominent ring of the radially averaged FFT magnitude of
    the background-subtracted image; 15.0 when no ring stands out (utils.py:23-108)."""
    n = image.shape[0]
    flat = np.asarray(image, dtype=np.float64) - gaussian_filter(image, sigma=n * 0.005).astype(np.float64)
    mag, _ = fft_spectra(flat)
    yy, xx = np.ogrid[:n, :n]
    ring = np.sqrt((xx - n // 2) ** 2 + (yy - n // 2) ** 2).astype(np.int32).ravel()
    total = np.bincount(ring, mag.ravel(), minlength=n)
    count = np.bincount(ring, minlength=n)
    count[count == 0] = 1
    profile = total / count
    r_lo = max(2, int(n / max_atom_size))
    r_hi = min(len(profile) - 1, int(n / min_atom_size))
    window = profile[r_lo:r_hi + 1]
    peaks, _ = find_peaks(window, prominence=np.max(window) * prominence_factor)
    if len(peaks) == 0:
        return 15.0
    return n / (peaks[0] + r_lo)


def load_image_from_h5(file_path, dataset_name: str | None = None) -> np.ndarray:
    """2-D dataset of an HDF5 file: `dataset_name` (full path, else first dataset with that base name), else the
    largest 2-D dataset, preferring ones called image / data / HAADF (utils.py:111-185)."""
    try:
        import h5py
    except ImportError as e:                                     # pragma: no cover - depends on the image
        raise ImportError("livae.utils.load_image_from_h5 needs h5py, which is not installed here") from e
    path = Path(file_path)
    with h5py.File(path, "r") as f:
        found: list[tuple[str, tuple[int, ...]]] = []
        f.visititems(lambda name, obj: found.append((name, tuple(int(s) for s in obj.shape)))
                     if isinstance(obj, h5py.Dataset) else None)
        chosen = None
        if dataset_name is not None:
            if dataset_name in f:
                chosen = dataset_name
            else:
                base = Path(dataset_name).name
                chosen = next((n for n, _ in found if Path(n).name == base), None)
        if chosen is None:
            planar = [(n, s) for n, s in found if len(s) == 2]
            if not planar:
                raise KeyError(f"No 2D datasets found in HDF5 file: {path}")
            # stable sort, descending by (preferred name, area): ties keep the file's visiting order
            planar.sort(key=lambda it: (Path(it[0]).name in ("image", "data", "HAADF"), it[1][0] * it[1][1]),
                        reverse=True)
            chosen = planar[0][0]
        return f[chosen][:]


def clean_state_dict(state_dict):
    """strip the `_orig_mod.` prefixes torch.compile adds to checkpoint keys (utils.py:188-196)"""
    return {k.replace("_orig_mod.", ""): v for k, v in state_dict.items()}

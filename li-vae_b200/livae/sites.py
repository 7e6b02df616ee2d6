"""livae.sites -- where the patches are: peak detection and lattice-site extrapolation on whole micrographs
(reference data.py:28-75 generate_lattice_grid, 119-148 get_clean_peaks, 176-202 / 328-345 pre-processing and
edge exclusion, 349-460 adaptive lattice sites).

One-shot host work at dataset construction, vectorised numpy / scipy instead of the
reference's per-atom Python loops; plain module (numpy + scipy only, no package imports at module level) so that
tests/golden/make_golden.py can load it by path next to the reference package.
"""
from __future__ import annotations

import numpy as np

__all__ = ["generate_lattice_grid", "peak_local_max", "get_clean_peaks", "adaptive_sites", "preprocess_image",
           "inside_margin"]

def generate_lattice_grid(image_shape, lattice_spacing: float, offset=(0, 0)) -> np.ndarray:
    """(y, x) points of a hexagonal grid: rows `lattice_spacing` apart, points 2*dx apart within a row with
    dx = spacing*sqrt(3)/2, odd rows shifted by dx (data.py:28-75).  The coordinates are built by repeated
    addition like the reference's while-loops so that the `< h` / `< w` cut-offs fall on the same points."""
    h, w = image_shape
    y_off, x_off = offset
    dx = lattice_spacing * np.sqrt(3) / 2
    pts = []
    y, row = y_off, 0
    while y < h:
        x = x_off + dx if row & 1 else x_off
        while x < w:
            pts.append((y, x))
            x += 2 * dx
        y += lattice_spacing
        row += 1
    return np.array(pts)


def peak_local_max(image: np.ndarray, min_distance: int = 1, threshold_rel: float | None = None) -> np.ndarray:
    """Integer (row, col) of local maxima, brightest first: pixels equal to the maximum of their
    (2*min_distance+1)^2 neighbourhood, above max(image.min(), threshold_rel*image.max()), not within
    `min_distance` of the border, and at least `min_distance` (Chebyshev) from any brighter accepted peak.
    Stands in for skimage.feature.peak_local_max(image, min_distance, threshold_rel) with its defaults
    (exclude_border=True, p_norm=inf) -- scikit-image is not installed in this image, so this restatement of its
    documented algorithm is NOT pinned against skimage itself."""
    from scipy.ndimage import maximum_filter
    from scipy.spatial import cKDTree
    img = np.asarray(image)
    d = max(int(min_distance), 0)
    size = 2 * d + 1
    thr = img.min()
    if threshold_rel is not None:
        thr = max(thr, threshold_rel * img.max())
    mask = (maximum_filter(img, size=size, mode="nearest") == img) & (img > thr)
    if d > 0:
        mask[:d, :] = False; mask[-d:, :] = False; mask[:, :d] = False; mask[:, -d:] = False
    rc = np.argwhere(mask)
    if len(rc) == 0:
        return rc.reshape(0, 2)
    rc = rc[np.argsort(-img[rc[:, 0], rc[:, 1]], kind="stable")]
    if d > 0 and len(rc) > 1:
        # plateaus / equal neighbours: keep the first (brightest) of every group closer than min_distance
        pairs = cKDTree(rc).query_pairs(r=d, p=np.inf, output_type="ndarray")
        if len(pairs):
            close = np.abs(rc[pairs[:, 0]] - rc[pairs[:, 1]]).max(1) < d
            pairs = pairs[close]
            keep = np.ones(len(rc), dtype=bool)
            for i, j in pairs[np.argsort(pairs[:, 0], kind="stable")]:          # i < j: i is the brighter one
                if keep[i]:
                    keep[j] = False
            rc = rc[keep]
    return rc


def get_clean_peaks(img: np.ndarray, min_distance: int = 5, threshold_rel: float = 0.01) -> np.ndarray:
    """peaks, each moved to the argmax of its 5x5 neighbourhood (data.py:119-148)"""
    peaks = peak_local_max(img, min_distance=min_distance, threshold_rel=threshold_rel)
    if len(peaks) == 0:
        return np.array([])
    h, w = img.shape
    # 5x5 windows clipped at the borders: take the row-major first maximum inside each clipped window
    pad = np.pad(img, 2, mode="constant", constant_values=-np.inf)
    win = np.lib.stride_tricks.sliding_window_view(pad, (5, 5))[peaks[:, 0], peaks[:, 1]].reshape(len(peaks), 25)
    k = np.argmax(win, axis=1)
    return np.stack([peaks[:, 0] - 2 + k // 5, peaks[:, 1] - 2 + k % 5], axis=1)


def preprocess_image(img: np.ndarray) -> np.ndarray:
    from livae.filter import bandpass_filter, normalize_image
    return normalize_image(bandpass_filter(img, 20, 100))               # data.py:176-179, 328-331


def inside_margin(c: np.ndarray, shape, margin: int) -> np.ndarray:
    return ((c[:, 0] >= margin) & (c[:, 0] <= shape[0] - margin) & (c[:, 1] >= margin) & (c[:, 1] <= shape[1] - margin))


_PAIRS_I, _PAIRS_J = np.triu_indices(6, k=1)          # (0,1), (0,2), ... in the reference's loop order


def adaptive_sites(img: np.ndarray, atoms: np.ndarray, lattice_spacing: float, half_patch: int,
                   detection_threshold: float = 0.6):
    """Lattice sites of one image from its detected atoms (data.py:349-460), vectorised:
    every atom predicts 8 neighbour sites from the most independent pair (v1, v2) of the vectors to its 6 nearest
    atoms (+-v1, +-v2, +-(v1+v2), +-(v1-v2)); all predictions closer than 0.35*spacing (single linkage) collapse
    to their centroid; a site is labelled 1 if an atom lies within detection_threshold*spacing of it.
    Ordering matches the reference: atoms first, then each atom's kept predictions; clusters by first member.
    -> (sites float64 [n,2], labels int [n])"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    atoms = np.asarray(atoms, dtype=np.float64).reshape(-1, 2)
    n = len(atoms)
    if n == 0:
        return np.zeros((0, 2)), np.zeros((0,), dtype=np.int64)
    tree = cKDTree(atoms)
    preds, owner = [atoms], [np.arange(n) * 9]
    k = min(7, n)
    if k >= 3:
        _, nb = tree.query(atoms, k=k)
        vec = atoms[nb[:, 1:]] - atoms[:, None, :]                             # [n, k-1, 2]
        pi, pj = (_PAIRS_I, _PAIRS_J) if k == 7 else np.triu_indices(k - 1, k=1)
        a, b = vec[:, pi], vec[:, pj]
        na, nb_ = np.linalg.norm(a, axis=2), np.linalg.norm(b, axis=2)
        ok = (na >= 1e-6) & (nb_ >= 1e-6)
        indep = np.where(ok, np.abs(a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]) / np.where(ok, na * nb_, 1.0), -2.0)
        best = np.argmax(indep, axis=1)                                        # first maximum, as `>` keeps it
        has = indep[np.arange(n), best] > -1
        v1, v2 = a[np.arange(n), best], b[np.arange(n), best]
        offs = np.stack([v1, -v1, v2, -v2, v1 + v2, -(v1 + v2), v1 - v2, v2 - v1], axis=1)     # [n, 8, 2]
        cand = atoms[:, None, :] + offs
        keep = has[:, None] & (cand[..., 0] >= half_patch) & (cand[..., 0] <= img.shape[0] - half_patch) & \
            (cand[..., 1] >= half_patch) & (cand[..., 1] <= img.shape[1] - half_patch)
        ai, oi = np.nonzero(keep)                                              # row-major: atom, then offset order
        preds.append(cand[ai, oi])
    pts = np.concatenate(preds)
    m = len(pts)
    pairs = cKDTree(pts).query_pairs(r=lattice_spacing * 0.35, output_type="ndarray")
    graph = coo_matrix((np.ones(len(pairs), dtype=np.int8), (pairs[:, 0], pairs[:, 1])), shape=(m, m))
    _, comp = connected_components(graph, directed=False)
    # clusters in the order of their first member (dict insertion order over i = 0..m-1 in the reference)
    first = np.full(comp.max() + 1, m, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(m))
    rank = np.empty_like(first)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    cid = rank[comp]
    cnt = np.bincount(cid).astype(np.float64)
    sites = np.stack([np.bincount(cid, pts[:, 0]) / cnt, np.bincount(cid, pts[:, 1]) / cnt], axis=1)
    dist, _ = tree.query(sites)
    return sites, (dist < lattice_spacing * detection_threshold).astype(np.int64)



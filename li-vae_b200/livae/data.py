"""Device-side patch pipeline: the per-item work of the reference's datasets (`PatchDataset`,
`AdaptiveLatticeDataset`, `PairedAdaptiveLatticeDataset` `__getitem__`, reference data.py:211-250, 478-560,
617-735) and `default_transform` (data.py:78-116), done for a whole batch by CUDA kernels on images resident
in HBM instead of by DataLoader worker processes.

The dataset classes of the reference (`PatchDataset`, `AdaptiveLatticeDataset`, `PairedAdaptiveLatticeDataset`)
are here with their constructor signatures, so scripts/train_rvae.py, train_vae.py and pretrain_stn.py build their
DataLoaders unchanged.  Construction (band-pass, lattice constant, peaks, lattice-site extrapolation: one-shot host
work, data.py:176-202, 299-473) runs on the host in vectorised numpy/scipy; `__getitem__` is the device path:
  * in the main process an item is produced by the CUDA kernels and handed back as CPU tensors like the
    reference's (`ds[i]`, DataLoader(num_workers=0));
  * in a DataLoader WORKER process (no CUDA after fork) an item is only its RECIPE -- flat index plus the draws the
    reference would have made from Python's `random`, in its order -- which the default collate function (a handler
    registered in torch's `default_collate_fn_map`) packs into a `RecipeBatch`; the loader's pin-memory step
    (`pin_memory=True` in all three scripts) runs in the main process and calls `RecipeBatch.pin_memory()`, where
    the whole batch is gathered / augmented / rotated / normalised on the GPU.  The script receives the usual
    `(patch, rotated, angle)` batch, already resident on the device.  Worker processes never touch the library.
`DevicePatchSource.from_dataset` also accepts a dataset object built by the reference itself.

Random numbers: the reference draws from Python's global `random` per item (scale, [angle], hflip, vflip,
shift_x, shift_y, then the pair angle).  The draws here are made on the host IN THE SAME ORDER from the same
global `random`, so `random.seed(s)` followed by items i0, i1, ... yields the reference's patches (to the fp32
tolerance of the bilinear resampling); only the pixel work moves to the device.  There is no CPU path.
"""
from __future__ import annotations

import math
import random
from typing import Iterable, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset, get_worker_info
from torch.utils.data._utils.collate import default_collate_fn_map

from livae import ops
from livae.sites import (adaptive_sites, generate_lattice_grid, get_clean_peaks, inside_margin as _inside,
                         peak_local_max, preprocess_image as _preprocess)

__all__ = ["PatchDataset", "AdaptiveLatticeDataset", "PairedAdaptiveLatticeDataset", "default_transform",
           "generate_lattice_grid", "get_clean_peaks", "peak_local_max", "draw_transform_params",
           "DevicePatchSource", "DevicePatchLoader", "PatchRecipe", "RecipeBatch"]


_FAST_DRAW_MIN = 64         # batches at least this large parse the generator's word stream instead of calling `random`
_DRAW_SLACK = 256           # words drawn beyond the expected need of a batch (a short block is redrawn twice as long)


def _draw(n: int, flip_prob: float, jitter_amount: int, rotation: bool, transform: bool, pair_angle: bool):
    """n consecutive items' draws from Python's `random` in the reference's order: per item the default_transform
    draws (scale, [angle], hflip, vflip, shift_x, shift_y; data.py:85-114) if `transform`, then the pair angle
    (data.py:695) if `pair_angle`.  Runs on the host once per batch: large batches take `_draw_stream` (the same
    numbers and the same final generator state, bit for bit, without 7 interpreter calls per patch -- 4.6 ms per
    2048-patch batch on the thread that also launches the training step)."""
    if n >= _FAST_DRAW_MIN and (transform or pair_angle) and type(random._inst) is random.Random:
        return _draw_stream(n, flip_prob, jitter_amount, rotation, transform, pair_angle)
    return _draw_calls(n, flip_prob, jitter_amount, rotation, transform, pair_angle)


def _draw_stream(n: int, flip_prob: float, jitter_amount: int, rotation: bool, transform: bool, pair_angle: bool):
    """`_draw_calls` restated on the 32-bit output stream of Python's Mersenne Twister, which numpy's legacy
    RandomState reproduces from the same 624-word state:
      random()       two words a, b -> ((a >> 5) * 2^26 + (b >> 6)) / 2^53                       (CPython _randommodule.c)
      uniform(a, b)  a + (b - a) * random()
      randint(-j, j) -j + _randbelow(2j + 1): one word >> (32 - k) per attempt, k = bit_length(2j + 1), repeated while
                     the value is >= 2j + 1 -- the only variable-length draw, so item boundaries are found by walking the
                     positions of the accepted words (two lookups per item)
    Afterwards `random`'s state is set to exactly the words consumed, so later draws continue the reference's stream."""
    pre = ((4 if rotation else 2) + 4) if transform else 0         # words before shift_x: scale, [angle], hflip, vflip
    post = 2 if pair_angle else 0
    width = 2 * jitter_amount + 1
    variable = transform and jitter_amount > 0
    kbits = width.bit_length()
    version, internal, gauss_next = random.getstate()
    key0, pos0 = np.array(internal[:624], dtype=np.uint32), internal[624]
    rs = np.random.RandomState()
    per_item = pre + post + (2 * 3 if variable else 0)              # expected 16 / 9 attempts per shift at j = 4
    m = max(n * per_item + _DRAW_SLACK, pre + post + 4)
    while True:
        rs.set_state(("MT19937", key0, pos0))
        w = rs.randint(0, 1 << 32, size=m, dtype=np.uint32)
        if not variable:
            starts = np.arange(n, dtype=np.int64) * (pre + post)
            jx = jy = None
            used = n * (pre + post)
            break
        ok = w < np.uint32(width << (32 - kbits))                # (w >> (32 - kbits)) < width: the attempt is accepted
        count = np.cumsum(ok, dtype=np.int32)                    # accepted words in [0, i]
        acc = np.append(np.flatnonzero(ok), [m, m])              # their positions (+ sentinels)
        cv, av = memoryview(count), memoryview(acc)
        starts, i, lim = [0] * n, 0, m - pre - post - 2
        for t in range(n):
            if i > lim:
                break
            starts[t] = i
            # shift_x = first accepted word at or after i + pre, shift_y = the next accepted one, then the pair angle
            i = av[cv[i + pre - 1] + 1] + 1 + post
        else:
            if i <= m:
                starts = np.asarray(starts, dtype=np.int64)
                rank = count[starts + (pre - 1)]
                jx, jy = acc[rank], acc[rank + 1]
                used = i
                break
        m *= 2                                                                  # ran out of words: draw a longer block

    def unit(at):                                                              # random() from the words at `at`, `at` + 1
        return ((w[at] >> np.uint32(5)).astype(np.float64) * 67108864.0
                + (w[at + 1] >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)

    p = None
    if transform:
        c = 2
        scale = 0.9 + (1.1 - 0.9) * unit(starts)
        if rotation:
            angle = 0.0 + 360.0 * unit(starts + c)
            c += 2
        else:
            angle = np.full(n, np.nan)
        flags = (unit(starts + c) < flip_prob).astype(np.int32) | ((unit(starts + c + 2) < flip_prob).astype(np.int32) << 1)
        shift = np.zeros((n, 2), dtype=np.int32)
        if variable:
            shift[:, 1] = (w[jx] >> np.uint32(32 - kbits)).astype(np.int32) - jitter_amount        # shift_x is drawn first
            shift[:, 0] = (w[jy] >> np.uint32(32 - kbits)).astype(np.int32) - jitter_amount
        p = {"scale": scale.astype(np.float32), "angle": angle, "flags": flags, "shift": shift}
    pair = None
    if pair_angle:
        at = (jy + 1) if variable else starts + pre
        pair = 0.0 + 360.0 * unit(at)
    rs.set_state(("MT19937", key0, pos0))
    if used:
        rs.randint(0, 1 << 32, size=used, dtype=np.uint32)
    _, key1, pos1 = rs.get_state()[:3]
    random.setstate((version, tuple(key1.tolist()) + (int(pos1),), gauss_next))
    return p, pair


def _draw_calls(n: int, flip_prob: float, jitter_amount: int, rotation: bool, transform: bool, pair_angle: bool):
    """the draws as the reference makes them: one call into `random` per number"""
    u, r, ri = random.uniform, random.random, random.randint
    scale, angle, flags, shift, pair = [], [], [], [], []
    for _ in range(n):
        if transform:
            scale.append(u(0.9, 1.1))
            if rotation:
                angle.append(u(0, 360))
            f = 1 if r() < flip_prob else 0
            if r() < flip_prob:
                f |= 2
            flags.append(f)
            if jitter_amount > 0:
                sx = ri(-jitter_amount, jitter_amount)
                sy = ri(-jitter_amount, jitter_amount)
                shift.append((sy, sx))
            else:
                shift.append((0, 0))
        if pair_angle:
            pair.append(u(0, 360))
    p = None
    if transform:
        p = {"scale": np.asarray(scale, np.float32),
             "angle": np.asarray(angle, np.float64) if rotation else np.full(n, np.nan),
             "flags": np.asarray(flags, np.int32).reshape(n), "shift": np.asarray(shift, np.int32).reshape(n, 2)}
    return p, (np.asarray(pair, np.float64) if pair_angle else None)


def draw_transform_params(n: int, flip_prob: float = 0.5, jitter_amount: int = 4, rotation: bool = False):
    """n consecutive default_transform draws from Python's `random`, in the reference's order (data.py:85-114).
    -> dict of host numpy arrays: scale f32 [n], angle f64 [n] (nan if not rotation), flags i32 [n], shift i32 [n,2]"""
    return _draw(n, flip_prob, jitter_amount, rotation, True, False)[0]


def _apply_transform(big: torch.Tensor, p: dict, rotation: bool) -> torch.Tensor:
    dev = big.device
    scale = torch.from_numpy(p["scale"]).to(dev)
    flags = torch.from_numpy(p["flags"]).to(dev)
    shift = torch.from_numpy(p["shift"]).to(dev)
    if not rotation:
        return ops.augment(big, scale, flags, shift)
    # scale -> rotate -> flips + roll (data.py:85-114): the rotation sits between the two halves of the fused kernel
    zeros_i = torch.zeros_like(flags)
    t = ops.augment(big, scale, zeros_i, torch.zeros_like(shift))
    t = ops.rotate_crop(t, big.shape[-1], torch.from_numpy(p["angle"]).to(dev))
    return ops.augment(t, scale, flags | 4, shift)


def default_transform(patch: torch.Tensor, flip_prob: float = 0.5, jitter_amount: int = 4,
                      rotation: bool = False) -> torch.Tensor:
    """Reference signature (data.py:78-83).  `patch`: CUDA float32 [C,S,S] (one item, C == 1) or [N,1,S,S]
    (a batch: one set of draws per patch, item order)."""
    if not patch.is_cuda:
        raise RuntimeError("livae.data.default_transform: CUDA tensor required; there is no CPU path")
    single = patch.dim() == 3
    x = patch.unsqueeze(0) if single else patch
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[-1] != x.shape[-2]:
        raise ValueError("default_transform: expected [1,S,S] or [N,1,S,S]")
    x = x.contiguous().float()
    out = _apply_transform(x, draw_transform_params(x.shape[0], flip_prob, jitter_amount, rotation), rotation)
    return out[0] if single else out


class DevicePatchSource:
    """Images + site lists resident on one GPU; batches of dataset items by index.

    images: sequence of equally sized 2-D arrays (float64 like the reference caches them, or float32) or a
    [n_img,H,W] tensor.  coords: per-image [n_i,2] arrays of (cy, cx) -- float sites (`sample_coords`) or integer
    peaks (`atom_coords`).  transform: None or `default_transform` (anything else is refused: the device
    kernels implement exactly that function)."""

    def __init__(self, images, coords: Sequence, patch_size: int = 128, padding: int = 32,
                 transform=None, device="cuda"):
        if transform not in (None, default_transform) and getattr(transform, "__name__", "") != "default_transform":
            raise ValueError("DevicePatchSource: transform must be None or default_transform")
        self.transform = transform
        self.patch_size, self.padding = int(patch_size), int(padding)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DevicePatchSource: a CUDA device is required; there is no CPU path")
        if isinstance(images, torch.Tensor):
            imgs = images
        else:
            imgs = torch.from_numpy(np.stack([np.asarray(i) for i in images]))
        if imgs.dim() != 3 or imgs.dtype not in (torch.float32, torch.float64):
            raise ValueError("DevicePatchSource: images must be [n_img,H,W] float32/float64")
        self.images = imgs.to(dev).contiguous()
        self.counts = [len(c) for c in coords]
        self._offsets = np.concatenate([[0], np.cumsum(self.counts)])
        flat = [np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in coords]
        self._yx = np.concatenate(flat) if flat else np.zeros((0, 2))
        self._img = np.repeat(np.arange(len(coords), dtype=np.int32), self.counts)
        self.device = dev
        self._slots = [(None, None)] * 3          # pinned staging slots (buffer, event of the copy that used it)
        self._slot_i = 0

    @classmethod
    def from_dataset(cls, ds, device="cuda"):
        """from a reference dataset object (PatchDataset: .atom_coords; Adaptive*: .sample_coords)"""
        coords = getattr(ds, "sample_coords", None)
        if coords is None:
            coords = ds.atom_coords
        return cls(ds.images, coords, ds.patch_size, ds.padding, getattr(ds, "transform", None), device)

    def __len__(self):
        return int(self._offsets[-1])

    def _lookup(self, indices):
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        if idx.size and (idx.min() < 0 or idx.max() >= len(self)):
            raise IndexError(f"Index out of range for dataset of size {len(self)}")      # data.py:217-220
        return self._img[idx], self._yx[idx]

    def _upload(self, cols):
        """All per-item numbers of a batch (site, draws) as ONE float64 [n,k] array through a pinned staging slot
        and one asynchronous copy: a pageable .to(device) per array would synchronise the host with the compute
        stream six times per batch and stop the CPU from running ahead of the training step."""
        n = cols[0].shape[0]
        host = np.concatenate([np.asarray(c, dtype=np.float64).reshape(n, -1) for c in cols], axis=1)
        k = host.shape[1]
        slot = self._slot_i % len(self._slots)
        self._slot_i += 1
        buf, ev = self._slots[slot]
        if buf is None or buf.shape[0] < n or buf.shape[1] != k:
            buf = torch.empty((max(n, 1), k), dtype=torch.float64, pin_memory=True)
        elif ev is not None:
            ev.synchronize()                     # the copy that last used this slot has finished
        buf.numpy()[:n] = host
        dev = buf[:n].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._slots[slot] = (buf, ev)
        return dev

    # --- PatchDataset.__getitem__ (data.py:211-250) -----------------------------------------------------
    def patch_batch(self, indices, draws=None) -> torch.Tensor:
        """transform=None: the bit-exact integer crop.  With default_transform the reference crops P + 2*padding,
        applies transform(patch_big, rotation=True) and centre-crops to P (data.py:240-248); same here."""
        img, yx = self._lookup(indices)
        if self.transform is None:
            d = self._upload([img, np.round(yx)])
            return ops.patch_gather(self.images, d.to(torch.int32).contiguous(), self.patch_size)
        p = draws[0] if draws is not None else draw_transform_params(len(img), rotation=True)
        d = self._upload([img, np.round(yx), p["scale"], p["flags"], p["shift"], p["angle"]])
        S = self.patch_size + 2 * self.padding
        big = ops.patch_gather(self.images, d[:, :3].to(torch.int32).contiguous(), S)
        q = self._params_on_device(d, 3)
        zeros_i = torch.zeros_like(q["flags"])
        # scale -> rotate -> flips + roll (data.py:85-114): the rotation sits between the two halves of `augment`
        t = ops.augment(big, q["scale"], zeros_i, torch.zeros_like(q["shift"]))
        t = ops.rotate_crop(t, S, d[:, 7].contiguous())
        t = ops.augment(t, q["scale"], q["flags"] | 4, q["shift"])
        return ops.rotate_crop(t, self.patch_size, None, normalise=False)

    def _big(self, d):
        S = self.patch_size + 2 * self.padding
        roi = self.patch_size + max(16, 2 * self.padding)
        return ops.patch_gather_roi(self.images, d[:, 0].to(torch.int32).contiguous(), d[:, 1:3].contiguous(), S, roi)

    @staticmethod
    def _params_on_device(d, c0):
        """columns c0.. of the uploaded array: scale, flags, shift_y, shift_x"""
        return {"scale": d[:, c0].float().contiguous(), "flags": d[:, c0 + 1].to(torch.int32).contiguous(),
                "shift": d[:, c0 + 2:c0 + 4].to(torch.int32).contiguous()}

    # --- AdaptiveLatticeDataset.__getitem__ (data.py:478-560) -------------------------------------------
    def adaptive_batch(self, indices, draws=None) -> torch.Tensor:
        img, yx = self._lookup(indices)
        cols = [img, yx]
        if self.transform is not None:
            p = draws[0] if draws is not None else draw_transform_params(len(img))
            cols += [p["scale"], p["flags"], p["shift"]]
        d = self._upload(cols)
        big = self._big(d)
        if self.transform is not None:
            q = self._params_on_device(d, 3)
            big = ops.augment(big, q["scale"], q["flags"], q["shift"])
        return ops.rotate_crop(big, self.patch_size, None, normalise=True)

    # --- PairedAdaptiveLatticeDataset.__getitem__ (data.py:617-735) -------------------------------------
    def paired_batch(self, indices, angles_deg: Optional[Iterable[float]] = None, draws=None):
        """-> (patch [N,1,P,P], rotated [N,1,P,P], angle_rad float32 [N]) as the default collate of the
        reference's items gives them.  Draw order per item: transform draws, then the pair angle (drawn even when
        angles_deg overrides it, so the random stream advances as in the reference).  `draws` = (transform draws,
        pair angles) made elsewhere (a DataLoader worker's recipes)."""
        img, yx = self._lookup(indices)
        n = len(img)
        p, ang = draws if draws is not None else _draw(n, 0.5, 4, False, self.transform is not None, True)
        if angles_deg is not None:
            ang = np.asarray(list(angles_deg), dtype=np.float64)
        cols = [img, yx, ang]
        if p is not None:
            cols += [p["scale"], p["flags"], p["shift"]]
        d = self._upload(cols)
        big = self._big(d)
        if p is not None:
            q = self._params_on_device(d, 4)
            big = ops.augment(big, q["scale"], q["flags"], q["shift"])
        ang_dev = d[:, 3].contiguous()
        patch = ops.rotate_crop(big, self.patch_size, None, normalise=True)
        rotated = ops.rotate_crop(big, self.patch_size, ang_dev, normalise=True)
        return patch, rotated, torch.deg2rad(ang_dev).float()


class DevicePatchLoader:
    """Minimal DataLoader stand-in over a DevicePatchSource: yields device batches in the shapes
    `train_rvae_one_epoch` / `train_one_epoch` unpack (train.py:315-324, 67-75).  Shuffling uses a numpy
    Generator (documented deviation: torch's DataLoader shuffles with the torch generator)."""

    def __init__(self, source: DevicePatchSource, batch_size: int, mode: str = "paired", shuffle: bool = True,
                 drop_last: bool = True, seed: int = 0, indices: Optional[Sequence[int]] = None):
        if mode not in ("paired", "adaptive", "patch"):
            raise ValueError("mode must be paired, adaptive or patch")
        self.source, self.batch_size, self.mode = source, int(batch_size), mode
        self.shuffle, self.drop_last = shuffle, drop_last
        self._rng = np.random.default_rng(seed)
        self._indices = np.arange(len(source)) if indices is None else np.asarray(indices, dtype=np.int64)

    def __len__(self):
        n = len(self._indices)
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)

    def __iter__(self):
        order = self._rng.permutation(self._indices) if self.shuffle else self._indices
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            if self.mode == "paired":
                yield self.source.paired_batch(idx)
            elif self.mode == "adaptive":
                yield self.source.adaptive_batch(idx)
            else:
                yield self.source.patch_batch(idx)


# ------------------------------------------------------------------------------------------------------------
# DataLoader-worker recipes
# ------------------------------------------------------------------------------------------------------------
_REGISTRY: dict = {}          # dataset key -> dataset object (main process; workers inherit a copy they never use)
_NEXT_KEY = [0]


class PatchRecipe:
    """What a DataLoader worker returns for one item instead of pixels: the flat index and the item's random draws
    (made from the worker's own `random`, seeded by torch per worker exactly as for the reference's items)."""
    __slots__ = ("key", "index", "t", "angle")

    def __init__(self, key, index, t, angle):
        self.key, self.index, self.t, self.angle = key, int(index), t, angle

    def __reduce__(self):
        return (PatchRecipe, (self.key, self.index, self.t, self.angle))


class RecipeBatch:
    """A collated batch of recipes.  `pin_memory()` -- called by the DataLoader's pin-memory thread in the MAIN
    process -- materialises it on the GPU; so does `materialise()` for loops that got it unpinned."""

    def __init__(self, key, indices, t, angles):
        self.key, self.indices, self.t, self.angles = key, indices, t, angles

    def __len__(self):
        return len(self.indices)

    def materialise(self):
        ds = _REGISTRY.get(self.key)
        if ds is None:
            raise RuntimeError("livae.data: recipe batch for a dataset that does not live in this process")
        return ds._batch(self.indices, (self.t, self.angles))

    def pin_memory(self):
        return self.materialise()


def _collate_recipes(batch, *, collate_fn_map=None):
    first = batch[0]
    t = None
    if first.t is not None:
        t = {k: np.concatenate([r.t[k] for r in batch]) for k in first.t}
    ang = None if first.angle is None else np.asarray([r.angle for r in batch], dtype=np.float64)
    return RecipeBatch(first.key, np.asarray([r.index for r in batch], dtype=np.int64), t, ang)


default_collate_fn_map[PatchRecipe] = _collate_recipes


class _DeviceDataset(Dataset):
    """shared item plumbing of the three datasets"""
    _kind = "patch"

    def _register(self):
        self._key = _NEXT_KEY[0]
        _NEXT_KEY[0] += 1
        _REGISTRY[self._key] = self
        self._src = None

    def _coords(self):
        return self.sample_coords if hasattr(self, "sample_coords") else self.atom_coords

    def __len__(self) -> int:
        return int(sum(len(c) for c in self._coords()))

    def source(self, device=None) -> DevicePatchSource:
        """the images + sites on the GPU (uploaded on first use)"""
        if self._src is None:
            if get_worker_info() is not None:
                raise RuntimeError("livae.data: DataLoader workers must not touch the GPU")
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            self._src = DevicePatchSource(self.images, self._coords(), self.patch_size, self.padding,
                                          self.transform, dev)
        return self._src

    def _draws(self, n):
        has_t = self.transform is not None
        if self._kind == "patch":
            return (draw_transform_params(n, rotation=True) if has_t else None), None
        return _draw(n, 0.5, 4, False, has_t, self._kind == "paired")

    def _batch(self, indices, draws=None):
        src = self.source()
        if draws is None:
            draws = self._draws(len(indices))
        if self._kind == "paired":
            return list(src.paired_batch(indices, draws=draws))
        if self._kind == "adaptive":
            return src.adaptive_batch(indices, draws=draws)
        return src.patch_batch(indices, draws=draws)

    def _check(self, idx):
        n = len(self)
        if idx < 0:
            idx += n                    # Python indexing convenience; the reference walks the lists forward only
        if not 0 <= idx < n:
            raise IndexError(f"Index {idx} out of range for dataset of size {n}")        # data.py:217-220
        return idx

    def __getitem__(self, idx):
        idx = self._check(int(idx))
        if get_worker_info() is not None:
            t, ang = self._draws(1)
            return PatchRecipe(self._key, idx, t, None if ang is None else float(ang[0]))
        out = self._batch([idx])
        if self._kind == "paired":
            patch, rotated, ang = out
            return patch[0].cpu(), rotated[0].cpu(), float(ang[0])
        return out[0].cpu()

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_src"] = None               # device tensors stay in the process that owns the GPU
        return st

    # ---- inspection plots of the reference's dataset classes (data.py:252-289, 562-612): host-side, matplotlib ----
    def _window(self, img_idx, coords, size, offset):
        """the displayed region of image `img_idx` and the sites that fall inside it, in the region's own coordinates
        -> (image view, sites, boolean mask over `coords`)"""
        img = self.images[img_idx]
        coords = np.asarray(coords).reshape(-1, 2)
        if size is None:
            return img, coords, np.ones(len(coords), dtype=bool)
        y0, x0 = offset
        inside = ((coords[:, 0] >= y0) & (coords[:, 0] < y0 + size) & (coords[:, 1] >= x0) & (coords[:, 1] < x0 + size))
        return img[y0:y0 + size, x0:x0 + size], coords[inside] - np.array([y0, x0]), inside

    @staticmethod
    def _show(img, groups, figsize, marker_size):
        import matplotlib.pyplot as plt            # not needed anywhere else in the package
        plt.figure(figsize=figsize)
        plt.imshow(img, cmap="gray")
        for pts in groups:
            if len(pts) > 0:
                plt.scatter(pts[:, 1], pts[:, 0], s=marker_size, c="red", marker="o", alpha=0.8)
        plt.axis("off")
        plt.show()


class PatchDataset(_DeviceDataset):
    """Patches centred on detected atoms (reference data.py:151-250): same constructor, attributes
    (`images`, `atom_coords`, `patch_size`, `padding`, `transform`) and item semantics."""
    _kind = "patch"

    def __init__(self, images, patch_size: int, padding: int = 4, transform=default_transform):
        from .utils import estimate_lattice_constant
        self.patch_size, self.padding, self.transform = patch_size, padding, transform
        print("Preprocessing images (caching)...")
        self.images = [_preprocess(im) for im in images]
        self.atom_coords = []
        margin = patch_size // 2 + padding
        for img in self.images:
            spacing = estimate_lattice_constant(img)
            coords = get_clean_peaks(img, min_distance=int(spacing * 0.15)).reshape(-1, 2)
            ok = _inside(coords, img.shape, margin)
            print(f"Detected {len(coords)} atoms, {int(ok.sum())} after edge exclusion.")
            self.atom_coords.append(coords[ok])
        self._register()

    def plot_peaks(self, img_idx: int, size: int | None = None, offset: tuple[int, int] = (0, 0)) -> None:
        """detected atoms over image `img_idx`, optionally only a size x size region at `offset` = (y, x)
        (reference data.py:252-289)"""
        img, pts, _ = self._window(img_idx, self.atom_coords[img_idx], size, offset)
        # the reference scatters unconditionally: an empty region still draws (an empty) scatter
        import matplotlib.pyplot as plt
        plt.figure(figsize=(6, 6))
        plt.imshow(img, cmap="gray")
        plt.scatter(pts[:, 1], pts[:, 0], s=30, c="red", marker="o", alpha=0.8)
        plt.axis("off")
        plt.show()


class AdaptiveLatticeDataset(_DeviceDataset):
    """Patches on every lattice site extrapolated from the detected atoms (reference data.py:292-560)."""
    _kind = "adaptive"

    def __init__(self, images, patch_size: int, padding: int = 48, transform=default_transform,
                 detection_threshold: float = 0.6):
        from .utils import estimate_lattice_constant
        self.patch_size, self.padding, self.transform = patch_size, padding, transform
        self.detection_threshold = detection_threshold
        self.images = [_preprocess(im) for im in images]
        self.sample_coords, self.labels = [], []
        half = patch_size // 2 + padding
        for img in self.images:
            spacing = estimate_lattice_constant(img)
            atoms = get_clean_peaks(img, min_distance=int(spacing * 0.15)).reshape(-1, 2)
            atoms = atoms[_inside(atoms, img.shape, half)]
            sites, labels = adaptive_sites(img, atoms, spacing, half, detection_threshold)
            print(f"Adaptive lattice: {len(sites)} unique sites - "
                  f"{int((labels == 1).sum())} with atoms, {int((labels == 0).sum())} empty sites")
            self.sample_coords.append(sites)
            self.labels.append(labels)
        self._register()

    def plot_lattice(self, img_idx: int, size: int | None = None, offset: tuple[int, int] = (0, 0)) -> None:
        """lattice sites over image `img_idx` -- sites with an atom, then empty sites, as two scatters in the same style
        (reference data.py:562-612)"""
        img, pts, inside = self._window(img_idx, self.sample_coords[img_idx], size, offset)
        labels = np.asarray(self.labels[img_idx])[inside]
        self._show(img, [pts[labels == 1], pts[labels == 0]], (8, 8), 50)

    @classmethod
    def from_sites(cls, images, sample_coords, patch_size: int, padding: int = 48, transform=default_transform,
                   labels=None):
        """a dataset over GIVEN images (already pre-processed, float32/64 [H,W]) and per-image float (y, x) sites:
        skips the site finding (synthetic lattices with analytic sites)"""
        self = cls.__new__(cls)
        self.patch_size, self.padding, self.transform = patch_size, padding, transform
        self.detection_threshold = 0.6
        self.images = list(images)
        self.sample_coords = [np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in sample_coords]
        self.labels = labels if labels is not None else [np.ones(len(c), dtype=np.int64) for c in self.sample_coords]
        self._register()
        return self


class PairedAdaptiveLatticeDataset(AdaptiveLatticeDataset):
    """(patch, patch rotated by a random angle, angle in radians) per lattice site (reference data.py:617-735)"""
    _kind = "paired"

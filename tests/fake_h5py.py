"""An in-memory stand-in for the few h5py calls `load_image_from_h5` makes (File as a context manager, `in`, item access
with `[:]`, `visititems`, the Dataset class) -- h5py is not installed in this image.  `FILES[path] = {dataset path: array}`;
groups are implied by the slashes in the dataset paths and visited depth-first in insertion order like h5py's default."""
import sys
import types

import numpy as np

FILES: dict = {}


class Dataset:
    def __init__(self, arr):
        self._a = np.asarray(arr)
        self.shape = self._a.shape

    def __getitem__(self, key):
        return self._a[key]


class Group:
    pass


class File:
    def __init__(self, path, mode="r"):
        self._d = FILES[str(path)]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __contains__(self, name):
        return name.strip("/") in self._d

    def __getitem__(self, name):
        return Dataset(self._d[name.strip("/")])

    def visititems(self, fn):
        seen = set()
        for name, arr in self._d.items():
            parts = name.split("/")
            for k in range(1, len(parts)):
                g = "/".join(parts[:k])
                if g not in seen:
                    seen.add(g)
                    if fn(g, Group()) is not None:
                        return
            if fn(name, Dataset(arr)) is not None:
                return


def install():
    mod = types.ModuleType("h5py")
    mod.File, mod.Dataset, mod.Group = File, Dataset, Group
    sys.modules["h5py"] = mod
    return mod


def cases():
    """(file contents, dataset_name argument, expected dataset path or the exception type)"""
    a = np.arange(12.0).reshape(3, 4)
    big = np.ones((8, 9))
    cube = np.zeros((2, 3, 4))
    return [
        ({"x/raw": a, "y/big": big}, "x/raw", "x/raw"),                       # full path given
        ({"x/raw": a, "y/big": big}, "/x/raw", "x/raw"),
        ({"x/raw": a, "y/raw": big}, "elsewhere/raw", "x/raw"),               # base name: the first one visited
        ({"x/raw": a, "y/big": big}, "missing", "y/big"),                     # unknown name: auto-detection
        ({"x/raw": a, "y/big": big}, None, "y/big"),                          # largest 2-D dataset
        ({"y/big": big, "x/HAADF": a}, None, "x/HAADF"),                      # preferred names beat area
        ({"p/data": a, "q/image": a, "r/other": big}, None, "p/data"),        # tie among preferred: visiting order
        ({"s/one": a, "t/two": a.T.copy()}, None, "s/one"),                   # equal areas: visiting order
        ({"c/cube": cube, "m/meta": np.arange(5)}, None, KeyError),           # nothing 2-D
        ({"c/cube": cube, "x/raw": a}, None, "x/raw"),
    ]

"""Import the UNMODIFIED reference package from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference does not
exist on the GPU box; everything that needs it is skipped there and the
committed vectors under tests/golden/ stand in for it.

The reference's ``livae/__init__.py`` imports matplotlib, skimage and h5py
(data.py:6,11; utils.py:5; train.py:15), none of which is installed here and
none of which the hot path uses; empty stub modules are registered before the
import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "livae"))


def load():
    """Return the reference ``livae`` package under the module name ``livae`` --
    callers must not have the product's ``livae`` imported in the same process."""
    if not available():
        raise RuntimeError("reference not present at /root/reference")
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "skimage", "skimage.feature"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sk = sys.modules["skimage.feature"]
    if not hasattr(sk, "peak_local_max"):
        sk.peak_local_max = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
    sys.modules["skimage"].feature = sk
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "livae" in sys.modules and not getattr(sys.modules["livae"], "__file__", "").startswith(REF_SRC):
        raise RuntimeError("a different 'livae' is already imported in this process")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    return importlib.import_module("livae")

"""Shared helpers for the parity tests."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 64


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def sample_idx(numel, name):
    h = int(hashlib.sha256(name.encode()).hexdigest()[:8], 16)
    rng = np.random.default_rng(h)
    return rng.integers(0, numel, size=min(SAMPLE, numel))


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_grads_against_golden(grads, gold, rtol, atol_frac=1e-6, names=None):
    """grads: dict name -> tensor.  Compares norm, sum-free sampled entries."""
    worst = 0.0
    for k, g in grads.items():
        if names is not None and k not in names:
            continue
        if "gnorm/" + k not in gold.files:
            continue
        g = torch.as_tensor(g).detach().double().cpu().reshape(-1)
        gn = float(gold["gnorm/" + k])
        samp = g[torch.from_numpy(sample_idx(g.numel(), k))].numpy()
        want = gold["gsamp/" + k]
        err = np.linalg.norm(samp - want) / max(np.linalg.norm(want), atol_frac * max(gn, 1e-30))
        nerr = abs(float(g.norm()) - gn) / max(gn, 1e-30)
        worst = max(worst, err, nerr)
        assert err <= rtol, f"{k}: sampled-grad rel err {err:.3e} > {rtol}"
        assert nerr <= rtol, f"{k}: grad-norm rel err {nerr:.3e} > {rtol}"
    return worst

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
        tol = rtol[k] if isinstance(rtol, dict) else rtol
        assert err <= tol, f"{k}: sampled-grad rel err {err:.3e} > {tol:.3e}"
        assert nerr <= tol, f"{k}: grad-norm rel err {nerr:.3e} > {tol:.3e}"
    return worst


def fp32_noise_floor(step_fn, params, *inputs, **kw):
    """How reproducible the REFERENCE algorithm's own fp32 gradients are: relative L2 distance, per
    parameter, between the oracle step evaluated in float32 and in float64 on the same inputs.
    ReLU / max-pool / bilinear-cell decisions flip under 1e-7 perturbations, so end-to-end
    gradients of the full step are only defined to this level (1e-3..1e-2 for the STN/encoder at
    P=128); single-kernel parity is tested at 1e-4 in test_gpu_kernels.py."""
    f64 = lambda t: t.double() if isinstance(t, torch.Tensor) and t.is_floating_point() else t
    _, g32 = step_fn(params, *inputs, **kw)
    _, g64 = step_fn({k: v.double() for k, v in params.items()}, *[f64(t) for t in inputs], **kw)
    out = {}
    for k in g32:
        d = float(g64[k].norm())
        out[k] = float((g32[k].double() - g64[k]).norm()) / d if d > 0 else 0.0
    return out


def grad_tolerances(floor, base=1e-3, factor=3.0, flip_tol=0.15):
    """Per-parameter relative-L2 tolerance for END-TO-END gradients.

    decoder.*: max(base, factor * fp32 noise floor).  encoder.* (STN + encoder convs + heads): the
    STN amplifies 1e-7 rounding differences to ~4e-6 rad in theta, i.e. ~2e-5 in the rotated input
    of the encoder (measured on B200 vs the CPU oracle); that is enough to flip the sign of a few
    near-zero ReLU pre-activations / max-pool winners, and each flip moves the gradients of all
    layers BELOW it by ~1/sqrt(#active units) (1-5 % at B <= 8).  The same layers match the CPU
    to 1e-6 when fed identical inputs (test_encoder_chain_strict), so end-to-end they are only held
    to `flip_tol`; the reference's own fp32-vs-fp64 gradients differ by the same mechanism."""
    out = {}
    for k, v in floor.items():
        t = max(base, factor * v)
        if k.startswith("encoder."):
            t = max(t, flip_tol)
        out[k] = t
    return out

"""Flat-buffer optimiser for the B200 path: all parameters live in ONE contiguous fp32 buffer and
all gradients in another, so that gradient all-reduce (NCCL), global-norm clipping (train.py:396)
and the Adam/AdamW update (scripts/train_rvae.py:157-159, scripts/train_vae.py:142) are each a
single launch over 9-12 MB instead of 30 per-parameter launches.  Semantics are torch.optim's.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["FlatAdamW"]


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (decoupled=True) / Adam (decoupled=False) on flattened parameters.

    After construction every p.data is a view into `flat_param` and every p.grad a view into
    `flat_grad` (autograd accumulates in place into them), state_dict keys of the MODEL are unchanged.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FlatAdamW: no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW: parameters must be on a CUDA device (no CPU path)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        n = sum(p.numel() for p in params)
        # keep every parameter 16-byte aligned inside the flat buffers (float4 kernels)
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.numel = n
        with torch.no_grad():
            for p, o in zip(params, offs):
                self.flat_param[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + p.numel()].view(p.shape)
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self._params = params
        self._offs = offs

    def zero_grad(self, set_to_none: bool = False):
        # gradients are persistent views; "None" would detach them from the flat buffer
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offs):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None, gscale=None):
        g = self.param_groups[0]
        ops.adamw_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_dev,
                   g["lr"], g["betas"], g["eps"], g["weight_decay"], decoupled=g["decoupled"], gscale=gscale)

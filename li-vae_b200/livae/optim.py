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

    After construction every p.data is a view into `flat_param`; after `sync_grads()` (called by `step`) every
    p.grad is a view into `flat_grad`.  state_dict keys of the MODEL are unchanged.  Deviation from torch.optim: a
    parameter that received no gradient is updated as with a zero gradient (moments decay, weight decay applies);
    torch skips such tensors.  Pass only the parameters that train (as scripts/train_rvae.py:143-159 does).
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
        self._views = [self.flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)]
        self._synced = True

    def zero_grad(self, set_to_none: bool = True):
        """Gradients are dropped, not zeroed: autograd then hands each parameter's gradient tensor over without an
        accumulation kernel (AccumulateGrad adds in place only into an existing .grad; that was ~40 tiny `add`
        launches per step into the flat views), and `sync_grads` packs them into `flat_grad` in one multi-tensor
        copy.  `set_to_none` is accepted for torch.optim compatibility; both values behave the same."""
        for p in self._params:
            p.grad = None
        self._synced = False

    @torch.no_grad()
    def sync_grads(self):
        """Pack the parameters' gradients into `flat_grad` and make every p.grad a view of it (idempotent).  Called
        by `step`; call it yourself before reading or reducing `flat_grad` (clip, all-reduce)."""
        if self._synced:
            return
        dst, src, missing = [], [], False
        for p, v in zip(self._params, self._views):
            g = p.grad
            if g is None:
                missing = True
            elif g.data_ptr() != v.data_ptr():
                dst.append(v); src.append(g.reshape(v.shape) if g.shape != v.shape else g)
        if missing:
            self.flat_grad.zero_()            # parameters that received no gradient contribute zeros
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self._params, self._views):
            p.grad = v
        self._synced = True

    @torch.no_grad()
    def step(self, closure=None, gscale=None):
        self.sync_grads()
        g = self.param_groups[0]
        ops.adamw_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_dev,
                   g["lr"], g["betas"], g["eps"], g["weight_decay"], decoupled=g["decoupled"], gscale=gscale)

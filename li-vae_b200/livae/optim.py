"""Flat-buffer optimiser for the B200 path: all parameters live in ONE contiguous fp32 buffer and
all gradients in another, so that gradient all-reduce (NCCL), global-norm clipping (train.py:396)
and the Adam/AdamW update (scripts/train_rvae.py:143-159, scripts/train_vae.py:142) are each a
single launch over 9-12 MB instead of 30 per-parameter launches.  Semantics are torch.optim's.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["FlatAdamW"]


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (decoupled=True) / Adam (decoupled=False) on flattened parameters.

    `params` is an iterable of tensors or of group dicts, exactly as for torch.optim (scripts/train_rvae.py:143-156
    builds two groups when --stn-lr is given); every group keeps its own lr / betas / eps / weight_decay and is
    updated by its own launch over its contiguous slice of the flat buffers.  After construction every p.data is a
    view into `flat_param`; after `sync_grads()` (called by `step`) every p.grad is a view into `flat_grad`.

    Like torch.optim, a parameter WITHOUT a gradient is skipped entirely -- no weight decay, no moment decay --
    which is what --freeze-stn relies on (scripts/train_rvae.py:184-189 clears requires_grad after the optimiser
    exists): the update is launched per run of consecutive parameters that did receive one.  The step count is
    per optimiser (torch keeps one per parameter; they only differ for a parameter that is frozen and later
    released).  `state_dict()` / `load_state_dict()` carry the moments and the step in torch.optim.AdamW's own
    format (per-parameter `exp_avg`, `exp_avg_sq`, `step`), so checkpoints written by the reference's scripts
    (`optimizer_state`, scripts/train_rvae.py:259-275) resume either way.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, decoupled=True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        flat = [p for g in self.param_groups for p in g["params"]]
        if not flat:
            raise ValueError("FlatAdamW: no parameters")
        dev = flat[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW: parameters must be on a CUDA device (no CPU path)")
        if any(p.device != dev or p.dtype != torch.float32 for p in flat):
            raise RuntimeError("FlatAdamW: all parameters must be float32 on one device")
        # keep every parameter 16-byte aligned inside the flat buffers (float4 kernels)
        offs, total = [], 0
        for p in flat:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.numel = sum(p.numel() for p in flat)
        with torch.no_grad():
            for p, o in zip(flat, offs):
                self.flat_param[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat_param[o:o + p.numel()].view(p.shape)
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self._params = flat
        self._offs = offs
        self._ends = [o + (p.numel() + 3) // 4 * 4 for p, o in zip(flat, offs)]
        self._group_of = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
        self._views = [self.flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(flat, offs)]
        self._active = [True] * len(flat)
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
        by `step`; call it yourself before reading or reducing `flat_grad` (clip, all-reduce).  Parameters that
        received no gradient (frozen, or off this step's graph) are remembered as inactive: their slice of
        `flat_grad` is zeroed (so the global norm ignores them) and `step` leaves them untouched."""
        if self._synced:
            return
        dst, src, idle = [], [], []
        for i, (p, v) in enumerate(zip(self._params, self._views)):
            g = p.grad
            self._active[i] = g is not None
            if g is None:
                idle.append(v)
            elif g.data_ptr() != v.data_ptr():
                dst.append(v); src.append(g.reshape(v.shape) if g.shape != v.shape else g)
        if idle:
            torch._foreach_zero_(idle)
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v, a in zip(self._params, self._views, self._active):
            p.grad = v if a else None
        self._synced = True

    def _runs(self):
        """[(group index, start, end)]: maximal runs of consecutive active parameters of one group"""
        runs = []
        for i, a in enumerate(self._active):
            if not a:
                continue
            gi, s, e = self._group_of[i], self._offs[i], self._ends[i]
            if runs and runs[-1][0] == gi and runs[-1][2] == s and self._active[i - 1]:
                runs[-1][2] = e
            else:
                runs.append([gi, s, e])
        return runs

    @torch.no_grad()
    def step(self, closure=None, gscale=None):
        self.sync_grads()
        runs = self._runs()
        for k, (gi, s, e) in enumerate(runs):
            g = self.param_groups[gi]
            ops.adamw_(self.flat_param[s:e], self.flat_grad[s:e], self.exp_avg[s:e], self.exp_avg_sq[s:e],
                       self.step_dev, g["lr"], g["betas"], g["eps"], g["weight_decay"], decoupled=g["decoupled"],
                       gscale=gscale, inc_step=(k == len(runs) - 1))
        ops.invalidate_weight_packs()          # the raw kernel changed the parameters behind autograd's version counter

    # ---- checkpointing in torch.optim.AdamW's format ---------------------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        step = self.step_dev.detach().clone().reshape(())
        state = {}
        for i, (p, o) in enumerate(zip(self._params, self._offs)):
            n = p.numel()
            state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"])
                                                        for a, b in zip(groups, self.param_groups)):
            raise ValueError("FlatAdamW.load_state_dict: parameter groups do not match")
        for mine, theirs in zip(self.param_groups, groups):
            for k, v in theirs.items():
                if k != "params":
                    mine[k] = v
            mine.setdefault("decoupled", self.defaults["decoupled"])
        steps = []
        ids = [i for g in groups for i in g["params"]]
        for slot, pid in enumerate(ids):
            st = state_dict["state"].get(pid)
            p, o = self._params[slot], self._offs[slot]
            n = p.numel()
            if st is None:
                self.exp_avg[o:o + n].zero_(); self.exp_avg_sq[o:o + n].zero_()
                continue
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(float(st["step"]))
        self.step_dev.fill_(max(steps) if steps else 0.0)

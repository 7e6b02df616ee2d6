"""livae.train -- drop-in for the training-step part of the reference's src/livae/train.py.

Same function names, signatures and metric keys (reference train.py:33-165, 168-278, 286-445,
448-556, 559-573, 670-677).  Differences that do not change results:
  * per-step metrics are accumulated ON DEVICE and read back once per epoch (the reference does
    >= 38 `.item()` syncs per step, train.py:399-427); the reported epoch averages are the same
    quantities;
  * clipping / gradient norm use one fused L2-norm kernel when the parameters' gradients live in a
    livae.optim flat buffer, torch.nn.utils.clip_grad_norm_ otherwise;
  * `scaler` (AMP GradScaler) is accepted for signature compatibility; the B200 path keeps fp32
    master tensors and does its own operand rounding inside the tensor-core kernels, so the scaler
    is not used.
Reference quirks preserved on purpose: `train_canonical_loss` is reported as 0.0
(train.py:310,436: never accumulated); evaluate_rvae reports only the LAST batch
(train.py:521-541 sit outside the loop); clipping is always on (max_norm 20 / 5).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Any

import numpy as np
import os

import torch
import torch.nn as nn

from . import ops

__all__ = [
    "train_one_epoch", "evaluate", "train_rvae_one_epoch", "evaluate_rvae", "rotate_to_canonical",
    "evaluate_rotation_invariance", "log_reconstructions_tensorboard", "compute_atom_position_accuracy",
    "log_scalar_metrics_tensorboard", "rvae_step_loss", "train_rvae_step", "GraphedRvaeStep", "MetricLogger", "compute_psnr",
    "compute_ssim", "get_rotation_stats",
]


class MetricLogger:
    """reference train.py:559-573"""

    def __init__(self):
        self.metrics = defaultdict(list)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.metrics[k].append(v)

    def get_averages(self) -> dict[str, Any]:
        return {k: np.mean(v) for k, v in self.metrics.items()}

    def reset(self):
        self.metrics.clear()


def get_rotation_stats(rotations: torch.Tensor) -> tuple[float, float]:
    """reference train.py:576-580"""
    angles = torch.atan2(rotations[:, 1], rotations[:, 0]) * (180.0 / np.pi)
    return torch.mean(angles).item(), torch.std(angles).item()


def _on_device(*ts):
    """metric helpers accept host tensors like the reference's (tests, notebooks): they are staged to the current CUDA
    device -- the arithmetic still runs in the kernels, there is no CPU implementation"""
    if all(t.is_cuda for t in ts):
        return ts
    dev = next((t.device for t in ts if t.is_cuda), None) or torch.device("cuda", torch.cuda.current_device())
    return tuple(t.to(dev) for t in ts)


def _psnr_dev(img1, img2, max_val: float = 1.0):
    img1, img2 = _on_device(img1.detach().float().contiguous(), img2.detach().float().contiguous())
    mse = ops.elbo_sums(img1, img2)[0] / img1.numel()
    return 20.0 * torch.log10(max_val / torch.sqrt(mse))


def compute_psnr(img1: torch.Tensor, img2: torch.Tensor, max_val: float = 1.0) -> float:
    """reference train.py:583-603 (MSE via the fused reduction kernel)"""
    with torch.no_grad():
        v = _psnr_dev(img1, img2, max_val).item()
    return float("inf") if np.isinf(v) else v


def _ssim_dev(img1, img2, window_size: int = 11, C1: float = 0.01 ** 2, C2: float = 0.03 ** 2):
    """mean box-filter SSIM map (reference train.py:606-667) as one fused pass over the two batches
    (csrc/ssim.cu) instead of five avg_pool2d passes and a dozen elementwise kernels; device scalar."""
    from ._lib import call, lib
    a, b = _on_device(img1.detach().contiguous().float(), img2.detach().contiguous().float())
    ops.require_cuda(a, b)
    assert a.shape == b.shape and a.dim() == 4
    planes, H, W = a.shape[0] * a.shape[1], a.shape[2], a.shape[3]
    ws = torch.empty(max(1, lib().livae_ssim_box_ws_floats(planes, H)), dtype=torch.float32, device=a.device)
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    call("livae_ssim_box", a, b, planes, H, W, window_size, float(C1), float(C2), ws, out)
    return out[0]


def compute_ssim(img1, img2, window_size: int = 11, C1: float = 0.01 ** 2, C2: float = 0.03 ** 2) -> float:
    with torch.no_grad():
        return _ssim_dev(img1, img2, window_size, C1, C2).item()


def rotate_to_canonical(x: torch.Tensor, theta: torch.Tensor, rotation_stn: nn.Module = None) -> torch.Tensor:
    """Rotate a batch to the canonical frame with the predicted angles (reference train.py:670-677):
    grid_sample(x, affine_grid(get_rotation_matrix(theta))) as ONE fused kernel; gradients flow to
    theta (x needs none)."""
    return ops.rot_sample(x, ops.angle_to_cs(theta), 1.0)


def rvae_step_loss(model, criterion, x, x_rotated=None, angle=None, canonical_weight: float = 0.2,
                   elide_dead_encoder: bool = False):
    """forward + loss of one rVAE training step, reference train.py:373-394 (fp32 branch).
    -> (loss, recon_loss, kld_loss, cycle_loss, canonical_loss, outputs)

    train.py:376-377 runs the FULL Encoder on x_rotated although only theta is consumed; that is
    reproduced by default.  elide_dead_encoder=True evaluates only the STN localisation (same
    results, the conv stack's mu/logvar are discarded at the call site)."""
    rotated_recon, canonical_recon, theta, mu, logvar = model(x)
    take = getattr(getattr(model, "encoder", None), "take_canonical", None)
    canonical_input = take(x, theta) if take is not None else None      # always popped: no batch left on the module
    if canonical_weight <= 0:
        canonical_input = None
    theta_rotated = None
    if x_rotated is not None:
        if elide_dead_encoder:
            theta_rotated = model.encoder.theta_only(x_rotated)
        else:
            _, _, theta_rotated = model.encoder(x_rotated)
        if take is not None:
            model.encoder._canonical = None      # drop the rotated pass's stash: nothing will ask for it
    loss, recon_loss, kld_loss, cycle_loss = criterion(rotated_recon, x, mu, logvar, theta, theta_rotated, angle)
    canonical_loss = torch.zeros((), device=loss.device)
    if canonical_weight > 0 and canonical_recon is not None:
        if canonical_input is None:
            canonical_input = rotate_to_canonical(x, theta, model.encoder.rotation_stn)
        canonical_loss = ops.elbo_sums(canonical_recon, canonical_input)[0] / canonical_recon.numel()
        loss = loss + canonical_weight * canonical_loss
    outs = _RvaeOutputs((rotated_recon, canonical_recon, theta, mu, logvar))
    outs.canonical_input = canonical_input if (canonical_weight > 0 and canonical_recon is not None) else None
    return loss, recon_loss, kld_loss, cycle_loss, canonical_loss, outs


class _RvaeOutputs(tuple):
    """the model's 5 outputs; `.canonical_input` additionally carries rotate_to_canonical(x, theta) when the step
    computed it, so the metric block (train.py:419-427) does not resample the batch a second time"""
    canonical_input = None


class DevicePrefetcher:
    """Iterates a DataLoader one batch ahead: batch i+1 is copied host->device on a side stream while step i
    computes, so the copy of a 2048-patch rVAE batch (268 MB, ~5 ms over PCIe 5) is hidden behind the step.
    Yields batches whose tensors are already on `device`; `_unpack_rvae_batch` / `.to(device)` are then no-ops.
    Same role as the reference's pin_memory + non_blocking copies (scripts/train_rvae.py:77-95, train.py:316-339),
    with the overlap made explicit.  Non-tensor items (the Python-float angles of a paired batch) pass through.

    The device side is TWO persistent staging slots (allocated once per batch shape): fresh allocations on a side
    stream made the caching allocator fall back to cudaMalloc -- a device-wide sync -- every step.  A yielded
    batch is therefore only valid until the batch after next is requested (the training loops consume it
    within the step)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self._slots = [{}, {}]

    def _stage(self, item, slot, path, stream):
        if isinstance(item, torch.Tensor):
            if item.device == self.device:
                return item
            buf = slot.get(path)
            if buf is None or buf.shape != item.shape or buf.dtype != item.dtype:
                buf = torch.empty(item.shape, dtype=item.dtype, device=self.device)
                slot[path] = buf
            with torch.cuda.stream(stream):
                buf.copy_(item, non_blocking=True)
            return buf
        if isinstance(item, (list, tuple)):
            return type(item)(self._stage(t, slot, path + (k,), stream) for k, t in enumerate(item))
        return item

    def __iter__(self):
        if self.device.type != "cuda":
            yield from self.loader
            return
        copy_stream = torch.cuda.Stream(self.device)
        it = iter(self.loader)
        n = 0

        def fetch():
            nonlocal n
            batch = next(it)
            # slot n%2 was read by step n-2, which is already queued (or done) on the compute stream
            copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            out = self._stage(batch, self._slots[n % 2], (), copy_stream)
            n += 1
            return out

        try:
            nxt = fetch()
        except StopIteration:
            return
        while True:
            torch.cuda.current_stream(self.device).wait_stream(copy_stream)   # batch i is on the device
            cur = nxt
            try:
                nxt = fetch()          # enqueue the copy of batch i+1 before step i's kernels are launched
            except StopIteration:
                yield cur
                return
            yield cur

    def __len__(self):
        return len(self.loader)


def _materialise(batch):
    """a livae.data.RecipeBatch that reached the loop un-pinned (DataLoader(pin_memory=False)) -> device tensors"""
    return batch.materialise() if hasattr(batch, "materialise") else batch


def _unpack_rvae_batch(batch, device):
    """reference train.py:316-339"""
    batch = _materialise(batch)
    if isinstance(batch, (list, tuple)):
        if len(batch) == 3:
            x, x_rotated, angle = batch
            x = x.to(device, non_blocking=True)
            x_rotated = x_rotated.to(device, non_blocking=True)
            angle = (angle.to(device, non_blocking=True) if isinstance(angle, torch.Tensor)
                     else torch.tensor(angle, dtype=torch.float32, device=device))
            return x, x_rotated, angle.to(torch.float32)
        if len(batch) == 2:
            x, x_rotated = batch
            return x.to(device, non_blocking=True), x_rotated.to(device, non_blocking=True), None
        return batch[0].to(device, non_blocking=True), None, None
    return batch.to(device, non_blocking=True), None, None


def _clip_and_norm(model, optimizer, max_norm):
    """-> device scalar: pre-clip global gradient norm (what clip_grad_norm_ returns)"""
    flat = getattr(optimizer, "flat_grad", None)
    if flat is not None:
        optimizer.sync_grads()
        return ops.l2norm_clip_(flat, max_norm, apply=True)[0]
    return torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)


def _post_clip_norm(pre_norm, max_norm):
    # the reference measures the norm AFTER clipping (train.py:399-405): min(norm, ~max_norm)
    coef = torch.clamp(max_norm / (pre_norm + 1e-6), max=1.0)
    return pre_norm * coef


class _DevAccum:
    """sums of device scalars, read back once"""

    def __init__(self, device):
        self.device = device
        self.sums = {}

    def add(self, **kw):
        for k, v in kw.items():
            v = v.detach().reshape(()) if isinstance(v, torch.Tensor) else torch.tensor(float(v), device=self.device)
            # the first value is copied: under GraphedRvaeStep it is a view of static graph memory that the next replay overwrites
            self.sums[k] = v.clone() if k not in self.sums else self.sums[k] + v

    def averages(self, n):
        if not self.sums:
            return {}
        keys = list(self.sums)
        vals = torch.stack([self.sums[k].to(torch.float32) for k in keys]).cpu().tolist()
        return {k: v / n for k, v in zip(keys, vals)}


def train_rvae_step(model, optimizer, criterion, batch, device, canonical_weight: float = 0.2,
                    max_norm: float = 20.0, reduce_grads=None, elide_dead_encoder: bool = False):
    """ONE batch of the reference's rVAE training loop (train.py:315-397): unpack + host->device
    copy, forward, loss, backward, [gradient all-reduce], clip, optimizer step.
    -> (loss, recon, kld, cycle, canonical, outputs, pre_clip_grad_norm), all device tensors."""
    x, x_rotated, angle = _unpack_rvae_batch(batch, device)
    optimizer.zero_grad(set_to_none=getattr(optimizer, "flat_grad", None) is None)
    loss, recon_l, kld_l, cycle_l, can_l, outs = rvae_step_loss(model, criterion, x, x_rotated, angle,
                                                              canonical_weight, elide_dead_encoder)
    loss.backward()
    if reduce_grads is not None:
        if hasattr(optimizer, "sync_grads"):
            optimizer.sync_grads()            # the all-reduce reads the flat gradient buffer
        reduce_grads()
    pre = _clip_and_norm(model, optimizer, max_norm)
    optimizer.step()
    return x, loss, recon_l, kld_l, cycle_l, can_l, outs, pre


class GraphedRvaeStep:
    """`train_rvae_step` as two CUDA graphs (forward + loss + backward + gradient packing | clip + AdamW), replayed per
    batch.  One training step is ~190 kernel launches issued from Python (4-5 ms of host time): irrelevant at 2048
    patches per GPU (18 ms of GPU work) but the whole cost at 256 (2.5 ms of GPU work), the strong-scaling case of
    SURVEY 8e.  The gradient all-reduce (`reduce_grads`) runs eagerly between the two graphs.  Needs
    livae.optim.FlatAdamW (its flat buffers are the static memory the second graph works on) and batches of one fixed
    shape; the first call captures (after two warm-up runs whose effect on parameters and optimiser state is undone).
    Learning rate, betas, eps, weight decay, the clip norm and the set of trainable parameters are baked into the
    captured launches, so they are compared on every call: after a change (an LR scheduler, --freeze-stn) this object
    refuses to replay -- build a new one (`train_rvae_one_epoch` does).
    Returns what train_rvae_step returns; the tensors are static graph memory, valid until the next call.
    With LIVAE_CUDA_GRAPH=1, `train_rvae_one_epoch` uses this class by itself when the optimiser is a FlatAdamW."""

    def __init__(self, model, optimizer, criterion, device, canonical_weight: float = 0.2, max_norm: float = 20.0,
                 reduce_grads=None, elide_dead_encoder: bool = False):
        if getattr(optimizer, "flat_grad", None) is None:
            raise ValueError("GraphedRvaeStep needs livae.optim.FlatAdamW")
        self.model, self.opt, self.crit, self.device = model, optimizer, criterion, torch.device(device)
        self.cw, self.max_norm, self.reduce, self.elide = canonical_weight, max_norm, reduce_grads, elide_dead_encoder
        self.g_fwd = self.g_opt = None
        self.sig = None

    def _sig(self):
        o = self.opt
        return (tuple((g.get("lr"), tuple(g.get("betas", ())), g.get("eps"), g.get("weight_decay"), g.get("decoupled"))
                      for g in o.param_groups),
                tuple(p.requires_grad for p in self.model.parameters()), self.model.training)

    def _fwd_bwd(self):
        self.opt.zero_grad()
        out = rvae_step_loss(self.model, self.crit, self.sx, self.sxr, self.sang, self.cw, self.elide)
        out[0].backward()
        self.opt.sync_grads()
        return out

    def _update(self):
        pre = _clip_and_norm(self.model, self.opt, self.max_norm)
        self.opt.step()
        return pre

    def _capture(self, x, xr, ang):
        o = self.opt
        self.sig = self._sig()
        self.sx, self.sxr, self.sang = x.clone(), xr.clone(), ang.clone()
        keep = [t.clone() for t in (o.flat_param, o.exp_avg, o.exp_avg_sq, o.step_dev)]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd()
                if self.reduce is not None:
                    self.reduce()
                self._update()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.g_fwd, self.g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd):
            self.outs = self._fwd_bwd()
        with torch.cuda.graph(self.g_opt, pool=self.g_fwd.pool()):
            self.pre = self._update()
        for dst, src in zip((o.flat_param, o.exp_avg, o.exp_avg_sq, o.step_dev), keep):     # undo warm-up + capture
            dst.copy_(src)
        ops.invalidate_weight_packs()

    def accepts(self, x, xr, ang) -> bool:
        """a paired batch of the captured shape (any paired batch before the first capture)"""
        if xr is None or ang is None or ang.dim() != 1:
            return False
        return self.g_fwd is None or (x.shape == self.sx.shape and ang.shape == self.sang.shape)

    def __call__(self, batch):
        x, xr, ang = _unpack_rvae_batch(batch, self.device)
        if xr is None or ang is None:
            raise ValueError("GraphedRvaeStep: paired batches (x, x_rotated, angle) only")
        if self.g_fwd is None:
            self._capture(x, xr, ang)
        elif self.sig != self._sig():
            raise ValueError("GraphedRvaeStep: optimiser hyper-parameters / trainable parameters / train mode changed since "
                             "the capture (they are baked into the graphs): build a new GraphedRvaeStep")
        elif x.shape != self.sx.shape:
            raise ValueError("GraphedRvaeStep: batch shape changed (drop_last=True keeps it fixed)")
        self.sx.copy_(x); self.sxr.copy_(xr); self.sang.copy_(ang)
        self.g_fwd.replay()
        if self.reduce is not None:
            self.reduce()
        self.g_opt.replay()
        ops.invalidate_weight_packs()          # the replay changed the parameters; eager code must re-pack
        loss, recon_l, kld_l, cycle_l, can_l, outs = self.outs
        return self.sx, loss, recon_l, kld_l, cycle_l, can_l, outs, self.pre


def _graph_step_for(model, optimizer, criterion, device, canonical_weight, max_norm, reduce_grads):
    """the GraphedRvaeStep of this (model, optimiser, criterion, ...) combination, kept on the optimiser between epochs;
    None unless LIVAE_CUDA_GRAPH=1 and the optimiser is a FlatAdamW (the reference scripts' torch.optim.AdamW: eager
    launches), or after a failed capture.  A change of the hyper-parameters baked into the graphs replaces the object."""
    if getattr(optimizer, "flat_grad", None) is None or os.environ.get("LIVAE_CUDA_GRAPH", "0") != "1":
        return None
    ent = getattr(optimizer, "_livae_graph_step", None)
    if ent is False:
        return None
    key = (id(model), id(criterion), str(torch.device(device)), float(canonical_weight), float(max_norm), id(reduce_grads))
    if ent is not None and ent[0] == key and (ent[1].sig is None or ent[1].sig == ent[1]._sig()):
        return ent[1]
    gs = GraphedRvaeStep(model, optimizer, criterion, device, canonical_weight, max_norm, reduce_grads)
    optimizer._livae_graph_step = (key, gs)
    return gs


def train_rvae_one_epoch(model, data_loader, optimizer, criterion, metric_logger, device,
                         canonical_weight: float = 0.2, scaler=None, grad_max_norm: float | None = None,
                         reduce_grads=None) -> None:
    """reference train.py:286-445.  `reduce_grads` (not in the reference, keyword only in practice): a callable run
    between backward and clip -- livae.parallel.GradAverager under data parallelism."""
    model.train()
    acc = _DevAccum(device)
    n_batches = 0
    max_norm = grad_max_norm if grad_max_norm is not None else 20.0
    gstep = _graph_step_for(model, optimizer, criterion, device, canonical_weight, max_norm, reduce_grads)
    for batch in DevicePrefetcher(data_loader, device):
        step_out = None
        if gstep is not None:
            if gstep.sig is not None and gstep.sig != gstep._sig():      # e.g. a per-batch LR schedule: a fresh capture
                gstep = _graph_step_for(model, optimizer, criterion, device, canonical_weight, max_norm, reduce_grads)
        if gstep is not None:
            batch = _unpack_rvae_batch(batch, device)
            if gstep.accepts(*batch):              # ragged last batch, unpaired data: the eager step below
                try:
                    step_out = gstep(batch)
                except Exception as e:             # capture failed (never seen; e.g. out of memory for the graph pool)
                    if gstep.g_fwd is not None and gstep.g_opt is not None:
                        raise
                    import warnings
                    warnings.warn(f"livae: CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                                  "continuing with eager launches")
                    optimizer._livae_graph_step = False
                    gstep = None
                    torch.cuda.synchronize()
            if batch[1] is None:
                batch = batch[0]
            elif batch[2] is None:
                batch = batch[:2]
        if step_out is None:
            step_out = train_rvae_step(model, optimizer, criterion, batch, device, canonical_weight, max_norm, reduce_grads)
        x, loss, recon_l, kld_l, cycle_l, _can_l, outs, pre = step_out
        rotated_recon, canonical_recon, theta, mu, logvar = outs
        with torch.no_grad():
            m = dict(train_loss=loss, train_recon_loss=recon_l, train_kld_loss=kld_l, train_cycle_loss=cycle_l,
                     train_psnr=_psnr_dev(rotated_recon, x), train_ssim=_ssim_dev(rotated_recon, x),
                     train_latent_mean_abs=mu.abs().mean(), train_latent_std=torch.exp(0.5 * logvar).mean(),
                     train_grad_norm=_post_clip_norm(pre, max_norm))
            if theta is not None:
                m["train_rotation_std"] = torch.std(theta)
            if canonical_recon is not None:
                canonical_input = getattr(outs, "canonical_input", None)
                if canonical_input is not None:      # the step already formed it and its MSE (train.py:389-393)
                    m["train_canonical_psnr"] = 20.0 * torch.log10(1.0 / torch.sqrt(_can_l))
                else:
                    canonical_input = rotate_to_canonical(x, theta)
                    m["train_canonical_psnr"] = _psnr_dev(canonical_recon, canonical_input)
                m["train_canonical_ssim"] = _ssim_dev(canonical_recon, canonical_input)
            acc.add(**m)
        n_batches += 1
    avg = acc.averages(max(n_batches, 1))
    for k in ("train_rotation_std", "train_canonical_psnr", "train_canonical_ssim"):
        avg.setdefault(k, 0.0)
    avg["train_canonical_loss"] = 0.0   # reference never accumulates it (train.py:310, 436)
    metric_logger.update(**avg)


def train_one_epoch(model, data_loader, optimizer, criterion, metric_logger, device, scaler=None,
                    canonical_weight: float = 0.0) -> None:
    """reference train.py:33-165 (VAE and rVAE outputs both accepted; criterion returns 3 values)"""
    model.train()
    acc = _DevAccum(device)
    n_batches = 0
    canonical_batches = 0
    for x in DevicePrefetcher(data_loader, device):
        x = _materialise(x)
        if isinstance(x, (list, tuple)):
            x = x[0]
        x = x.to(device, non_blocking=True)
        optimizer.zero_grad(set_to_none=getattr(optimizer, "flat_grad", None) is None)
        outputs = model(x)
        canonical_recon = None
        canonical_input = None
        theta = None
        if len(outputs) == 3:
            recon, mu, logvar = outputs
            rotated_recon = recon
            loss, recon_l, kld_l = criterion(recon, x, mu, logvar)
        elif len(outputs) == 5:
            rotated_recon, canonical_recon, theta, mu, logvar = outputs
            # pop the encoder's one-shot stash of its rotated input (it IS rotate_to_canonical(x, theta), which the
            # reference computes here and discards, train.py:85-91): nothing stays pinned on the module
            take = getattr(getattr(model, "encoder", None), "take_canonical", None)
            canonical_input = take(x, theta) if take is not None else None
            loss, recon_l, kld_l = criterion(rotated_recon, x, mu, logvar)
        else:
            raise ValueError(f"Unexpected model output length: {len(outputs)}")
        loss.backward()
        pre = _clip_and_norm(model, optimizer, 5.0)
        optimizer.step()
        with torch.no_grad():
            m = dict(train_loss=loss, train_recon_loss=recon_l, train_kld_loss=kld_l,
                     train_psnr=_psnr_dev(rotated_recon, x), train_ssim=_ssim_dev(rotated_recon, x),
                     train_latent_mean_abs=mu.abs().mean(), train_latent_std=torch.exp(0.5 * logvar).mean(),
                     train_grad_norm=_post_clip_norm(pre, 5.0))
            m["train_rotation_std"] = torch.std(theta) if theta is not None else torch.zeros((), device=device)
            if canonical_recon is not None and theta is not None:
                if canonical_input is None:
                    canonical_input = rotate_to_canonical(x, theta)
                m["train_canonical_psnr"] = _psnr_dev(canonical_recon, canonical_input.detach())
                m["train_canonical_ssim"] = _ssim_dev(canonical_recon, canonical_input.detach())
                canonical_batches += 1
            acc.add(**m)
        n_batches += 1
    avg = acc.averages(max(n_batches, 1))
    if canonical_batches == 0:
        avg.pop("train_canonical_psnr", None)
        avg.pop("train_canonical_ssim", None)
    metric_logger.update(**avg)


def evaluate(model, data_loader, criterion, metric_logger, device, canonical_weight: float = 0.0) -> None:
    """reference train.py:168-278"""
    model.eval()
    acc = _DevAccum(device)
    n_batches = 0
    canonical_batches = 0
    with torch.no_grad():
        for x in data_loader:
            x = _materialise(x)
            if isinstance(x, (list, tuple)):
                x = x[0]
            x = x.to(device)
            outputs = model(x)
            canonical_recon = None
            theta = None
            if len(outputs) == 3:
                rotated_recon, mu, logvar = outputs
            elif len(outputs) == 5:
                rotated_recon, canonical_recon, theta, mu, logvar = outputs
            else:
                raise ValueError(f"Unexpected model output length: {len(outputs)}")
            loss, recon_l, kld_l = criterion(rotated_recon, x, mu, logvar)
            has_can = canonical_recon is not None and theta is not None
            if canonical_weight > 0 and has_can:
                canonical_input = rotate_to_canonical(x, theta)
                loss = loss + canonical_weight * (ops.elbo_sums(canonical_recon, canonical_input)[0]
                                                  / canonical_recon.numel())
            m = dict(val_loss=loss, val_recon_loss=recon_l, val_kld_loss=kld_l,
                     val_psnr=_psnr_dev(rotated_recon, x), val_ssim=_ssim_dev(rotated_recon, x),
                     val_latent_mean_abs=mu.abs().mean(), val_latent_std=torch.exp(0.5 * logvar).mean())
            m["val_rotation_std"] = torch.std(theta) if theta is not None else torch.zeros((), device=device)
            if has_can:
                canonical_input = rotate_to_canonical(x, theta)
                m["val_canonical_psnr"] = _psnr_dev(canonical_recon, canonical_input)
                m["val_canonical_ssim"] = _ssim_dev(canonical_recon, canonical_input)
                canonical_batches += 1
            acc.add(**m)
            n_batches += 1
    avg = acc.averages(max(n_batches, 1))
    if canonical_batches == 0:
        avg.pop("val_canonical_psnr", None)
        avg.pop("val_canonical_ssim", None)
    metric_logger.update(**avg)


def evaluate_rvae(model, data_loader, criterion, metric_logger, device, canonical_weight: float = 0.2) -> None:
    """reference train.py:448-556.  The reference's accumulators sit outside its batch loop, so it
    reports the metrics of the LAST validation batch only; that behaviour is preserved."""
    model.eval()
    last = None
    with torch.no_grad():
        for batch in data_loader:
            x, x_rotated, angle = _unpack_rvae_batch(batch, device)
            loss, recon_l, kld_l, cycle_l, can_l, outs = rvae_step_loss(model, criterion, x, x_rotated, angle,
                                                                      canonical_weight)
            last = (x, loss, recon_l, kld_l, cycle_l, can_l, outs)
        if last is None:
            raise ZeroDivisionError("evaluate_rvae: empty data loader")
        x, loss, recon_l, kld_l, cycle_l, can_l, (rotated_recon, canonical_recon, theta, mu, logvar) = last
        acc = _DevAccum(device)
        m = dict(val_loss=loss, val_recon_loss=recon_l, val_kld_loss=kld_l, val_cycle_loss=cycle_l,
                 val_canonical_loss=canonical_weight * can_l,
                 val_psnr=_psnr_dev(rotated_recon, x), val_ssim=_ssim_dev(rotated_recon, x),
                 val_latent_mean_abs=mu.abs().mean(), val_latent_std=torch.exp(0.5 * logvar).mean(),
                 val_rotation_std=torch.std(theta))
        canonical_input = rotate_to_canonical(x, theta)
        m["val_canonical_psnr"] = _psnr_dev(canonical_recon, canonical_input)
        m["val_canonical_ssim"] = _ssim_dev(canonical_recon, canonical_input)
        acc.add(**m)
    metric_logger.update(**acc.averages(1))


# ------------------------------------------------------------------------------------------------------------
# Logging / analysis helpers the reference's scripts import next to the training loops (train.py:680-936).
# Off the hot path: they run a forward pass through the same kernels and hand CPU tensors to TensorBoard.
# ------------------------------------------------------------------------------------------------------------
def _split_outputs(outputs):
    if len(outputs) == 3:
        return outputs[0], None, None, outputs[1]
    if len(outputs) == 5:
        return outputs[0], outputs[1], outputs[2], outputs[3]
    raise ValueError(f"Unexpected model output length: {len(outputs)}")


def log_reconstructions_tensorboard(model, images, writer, global_step: int, device, tag: str = "recon",
                                    normalize: bool = True, nrow: int | None = None) -> None:
    """[original | reconstruction | abs diff] image grids, plus the canonical-frame triple for an rVAE
    (reference train.py:791-853)"""
    from torchvision.utils import make_grid
    model.eval()
    with torch.no_grad():
        x = _materialise(images)
        if isinstance(x, (list, tuple)):
            x = x[0]
        x_dev = x.to(device)
        rotated_recon, canonical_recon, theta, _ = _split_outputs(model(x_dev))
        x_cpu, r_cpu = x_dev.detach().cpu(), rotated_recon.detach().cpu()
        n = nrow or x_cpu.size(0)
        grid = make_grid(torch.cat([x_cpu, r_cpu, (x_cpu - r_cpu).abs()], dim=0), nrow=n, normalize=normalize)
        writer.add_image(f"{tag}/original_recon_diff", grid, global_step)
        stn = getattr(getattr(model, "encoder", None), "rotation_stn", None)
        if canonical_recon is not None and theta is not None and stn is not None:
            c_in = rotate_to_canonical(x_dev, theta, stn).detach().cpu()
            c_rec = canonical_recon.detach().cpu()
            grid = make_grid(torch.cat([c_in, c_rec, (c_in - c_rec).abs()], dim=0), nrow=n, normalize=normalize)
            writer.add_image(f"{tag}/canonical_original_recon_diff", grid, global_step)


def log_scalar_metrics_tensorboard(writer, metrics: dict, global_step: int, prefix: str = "") -> None:
    """reference train.py:928-936"""
    for name, value in metrics.items():
        writer.add_scalar(f"{prefix}{name}", value, global_step)


def evaluate_rotation_invariance(model, images: torch.Tensor, angles=(0, 45, 90, 135, 180, 225, 270, 315),
                                 device=torch.device("cuda"), max_batches: int | None = None) -> dict:
    """Latent variance / reconstruction error / angle error of the model over rotated copies of a few images
    (reference train.py:680-788).  The rotations (TF.rotate, bilinear, fill 0) run in the `rotate_crop` kernel.
    Deviation, on purpose: the reference reads the predicted angle as atan2(theta[0,1], theta[0,0]), which raises
    IndexError for the real RVAE's theta [B,1]; a one-column theta is taken as the angle itself.
    `device` defaults to cuda (the reference's default is cpu; this package has no CPU path)."""
    model.eval()
    angles = [float(a) for a in angles]
    if images.dim() != 4:
        raise ValueError("images must have shape [B, C, H, W]")
    lat_var, rmse, psnr, ssim, ang_err = [], [], [], [], []
    S = images.shape[-1]
    with torch.no_grad():
        for i, img in enumerate(images):
            if max_batches is not None and i >= max_batches:
                break
            x = img.to(device).float().reshape(1, 1, S, S).expand(len(angles), 1, S, S).contiguous()
            deg = torch.tensor(angles, dtype=torch.float64, device=x.device)
            rotated = ops.rotate_crop(x, S, deg)
            outs = model(rotated)
            if len(outs) != 5:
                raise ValueError(f"Unexpected model output length: {len(outs)}")
            rotated_recon, _, theta, mu, _ = outs
            back = ops.rotate_crop(rotated_recon.contiguous(), S, -deg)
            lat_var.append(torch.var(mu, dim=0).mean().item())
            for r in back:
                r = r.unsqueeze(0)
                rmse.append(torch.sqrt(torch.mean((r - x[:1]) ** 2)).item())
                psnr.append(compute_psnr(r, x[:1]))
                ssim.append(compute_ssim(r, x[:1]))
            if theta is not None:
                pred = (torch.atan2(theta[:, 1], theta[:, 0]) if theta.shape[1] >= 2 else theta[:, 0]) * (180.0 / np.pi)
                d = (pred.cpu().double() - torch.tensor(angles, dtype=torch.float64)).abs()
                ang_err.append(float(torch.minimum(d, 360 - d).mean()))
    mean = lambda v: float(np.mean(v)) if v else 0.0
    return {"rotation_latent_variance": mean(lat_var), "rotation_recon_rmse": mean(rmse),
            "rotation_recon_psnr": mean(psnr), "rotation_recon_ssim": mean(ssim), "rotation_angle_error": mean(ang_err)}


def compute_atom_position_accuracy(original: torch.Tensor, reconstruction: torch.Tensor, lattice_spacing: float,
                                   threshold_ratio: float = 0.35) -> dict:
    """peak positions of a patch vs those of its reconstruction (reference train.py:856-925); host-side analysis"""
    from scipy.spatial.distance import cdist
    from .data import peak_local_max

    def plane(t):
        if t.dim() == 3:
            t = t[0] if t.size(0) == 1 else t.mean(dim=0)
        return t.detach().cpu().numpy()

    a, b = plane(original), plane(reconstruction)
    if lattice_spacing <= 0:
        raise ValueError("lattice_spacing must be positive")
    dmin = max(int(lattice_spacing * threshold_ratio), 1)
    pa, pb = peak_local_max(a, min_distance=dmin), peak_local_max(b, min_distance=dmin)
    if pa.size == 0 or pb.size == 0:
        return {"atom_detection_rate": 0.0, "atom_position_accuracy": 0.0, "atom_mean_position_error": float("inf"),
                "n_original_atoms": int(pa.shape[0]) if pa.size else 0,
                "n_reconstructed_atoms": int(pb.shape[0]) if pb.size else 0}
    nearest = cdist(pa, pb).min(axis=1)
    return {"atom_detection_rate": float(pb.shape[0] / pa.shape[0]),
            "atom_position_accuracy": float((nearest < lattice_spacing * threshold_ratio).sum() / pa.shape[0]),
            "atom_mean_position_error": float(nearest.mean()),
            "n_original_atoms": int(pa.shape[0]), "n_reconstructed_atoms": int(pb.shape[0])}

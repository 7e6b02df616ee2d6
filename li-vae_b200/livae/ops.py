"""torch.autograd seams over the C-ABI kernels (one Function per kernel family).

Activations between layers are NHWC float32 ([B,H,W,C], contiguous); 1-channel images are the
same memory as the reference's NCHW tensors.  Parameters keep torch's own layouts and dtypes
(SURVEY.md section 8b), so state_dicts and optimisers are interchangeable with the reference.
No op here has a CPU or ATen fallback.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, CONV, CONVT, ConvDesc, call, require_cuda

_scratch = {}


def _get_scratch(dev, key, nfloats):
    k = (dev.index, key)
    t = _scratch.get(k)
    if t is None or t.numel() < nfloats:
        t = torch.zeros(nfloats, dtype=torch.float32, device=dev)
        _scratch[k] = t
    return t


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------
# convolution / linear layers
# ------------------------------------------------------------------------------------------
class _ConvFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, kind, kh, kw, stride, pad, act, pool):
        x = _c(x); w = _c(w)
        require_cuda(x, w, b)
        B, Hin, Win, Cin = x.shape
        if kind == CONV:
            Cout = w.shape[0]
            assert w.numel() == Cout * Cin * kh * kw, "conv weight shape mismatch"
        else:
            assert w.shape[0] == Cin, "conv_transpose weight shape mismatch"
            Cout = w.shape[1]
        d = ConvDesc(kind, B, Hin, Win, Cin, Cout, kh, kw, stride, pad, act, int(pool))
        ho, wo = C.c_int(), C.c_int()
        L.lib().livae_conv_out_shape(C.byref(d), C.byref(ho), C.byref(wo))
        y = torch.empty((B, ho.value, wo.value, Cout), dtype=torch.float32, device=x.device)
        idx = ws = None
        if pool:
            idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
            ws = torch.empty(L.lib().livae_conv_fwd_ws_bytes(C.byref(d)) // 4, dtype=torch.float32,
                             device=x.device)
        call("livae_conv_fwd", C.byref(d), x, w, b, y, idx, ws)
        ctx.desc = d
        ctx.save_for_backward(x, w, y, idx)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, y, idx = ctx.saved_tensors
        gy = _c(gy)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(w) if need_w else None
        gb = torch.empty(y.shape[-1], dtype=torch.float32, device=x.device) if need_b else None
        if need_x or need_w or need_b:
            call("livae_conv_bwd", C.byref(ctx.desc), x, w, y, gy, idx, gw, gb, gx)
        return gx, gw, gb, None, None, None, None, None, None, None


def conv2d(x, w, b, kh, kw, stride, pad, act=ACT_NONE, pool=False):
    """nn.Conv2d (+ReLU/Sigmoid, +MaxPool2d(2,2)) on NHWC input (model.py:204-209, 290-296, 359-371)"""
    return _ConvFn.apply(x, w, b, CONV, kh, kw, stride, pad, act, pool)


def conv_transpose2d(x, w, b, kh, kw, stride, pad, act=ACT_NONE):
    """nn.ConvTranspose2d (+ReLU/Sigmoid) on NHWC input (model.py:90-96)"""
    return _ConvFn.apply(x, w, b, CONVT, kh, kw, stride, pad, act, False)


def linear_nhwc(x, w, b, act=ACT_NONE):
    """nn.Linear applied to the reference's NCHW flatten of an NHWC feature map x [B,H,W,C]
    (model.py:210-213, 321-324): the conv whose kernel covers the whole map."""
    B, H, W, _ = x.shape
    return _ConvFn.apply(x, w, b, CONV, H, W, 1, 0, act, False).view(B, -1)


class _UpsamplePadFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        require_cuda(x)
        B, H, W, Cc = x.shape
        out = torch.empty((B, 2 * H + 2, 2 * W + 2, Cc), dtype=torch.float32, device=x.device)
        call("livae_upsample_pad_fwd", x, B, H, W, Cc, out)
        ctx.shape = (B, H, W, Cc)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, H, W, Cc = ctx.shape
        g = _c(g)
        gx = torch.empty((B, H, W, Cc), dtype=torch.float32, device=g.device)
        call("livae_upsample_pad_bwd", g, B, H, W, Cc, None, gx)
        return gx


def upsample_pad(x):
    """nn.Upsample(x2, bilinear, align_corners=False) + nn.ReflectionPad2d(1) (model.py:357-358)"""
    return _UpsamplePadFn.apply(x)


class _DecFcFn(Function):
    @staticmethod
    def forward(ctx, z, w, b, Cc, q):
        z = _c(z); w = _c(w)
        require_cuda(z, w, b)
        B, Ld = z.shape
        assert w.shape == (Cc * q * q, Ld)
        out = torch.empty((B, q, q, Cc), dtype=torch.float32, device=z.device)
        call("livae_decfc_fwd", z, w, b, B, Ld, Cc, q * q, out)
        ctx.save_for_backward(z, w, out)
        ctx.dims = (B, Ld, Cc, q * q)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        z, w, y = ctx.saved_tensors
        B, Ld, Cc, HW = ctx.dims
        gy = _c(gy)
        gw = torch.empty_like(w)
        gb = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        call("livae_decfc_bwd", z, w, y, gy, B, Ld, Cc, HW, gw, gb, gz)
        return gz, gw, gb, None, None


def decoder_fc(z, w, b, channels, q):
    """relu(Linear(z)).view(B, channels, q, q) as an NHWC tensor (model.py:353, 383-384)"""
    return _DecFcFn.apply(z, w, b, channels, q)


# ------------------------------------------------------------------------------------------
# rotation: head, angle -> (cos, sin), fused affine_grid + grid_sample
# ------------------------------------------------------------------------------------------
class _StnHeadFn(Function):
    @staticmethod
    def forward(ctx, vec):
        vec = _c(vec)
        require_cuda(vec)
        B = vec.shape[0]
        cs = torch.empty((B, 2), dtype=torch.float32, device=vec.device)
        theta = torch.empty((B, 1), dtype=torch.float32, device=vec.device)
        call("livae_stn_head_fwd", vec, B, cs, theta)
        ctx.save_for_backward(vec)
        return cs, theta

    @staticmethod
    @once_differentiable
    def backward(ctx, gcs, gtheta):
        (vec,) = ctx.saved_tensors
        gvec = torch.empty_like(vec)
        call("livae_stn_head_bwd", vec, _c(gcs) if gcs is not None else None,
             _c(gtheta) if gtheta is not None else None, vec.shape[0], gvec)
        return gvec


def stn_head(vec):
    """F.normalize(vec, eps=1e-6) -> (cos, sin) [B,2]; theta = atan2(sin, cos) [B,1] (model.py:245-261)"""
    return _StnHeadFn.apply(vec)


class _AngleToCsFn(Function):
    @staticmethod
    def forward(ctx, theta):
        theta = _c(theta)
        require_cuda(theta)
        B = theta.numel()
        cs = torch.empty((B, 2), dtype=torch.float32, device=theta.device)
        call("livae_angle_to_cs", theta, B, cs)
        ctx.save_for_backward(theta)
        return cs

    @staticmethod
    @once_differentiable
    def backward(ctx, gcs):
        (theta,) = ctx.saved_tensors
        g = torch.empty_like(theta)
        call("livae_angle_to_cs_bwd", theta, _c(gcs), theta.numel(), g)
        return g


def angle_to_cs(theta):
    """(cos theta, sin theta) [B,2] -- the two free entries of get_rotation_matrix (model.py:220-235)"""
    return _AngleToCsFn.apply(theta)


class _RotSampleFn(Function):
    @staticmethod
    def forward(ctx, img, cs, sgn):
        img = _c(img); cs = _c(cs)
        require_cuda(img, cs)
        B, Cc, H, W = img.shape
        out = torch.empty_like(img)
        call("livae_rot_sample_fwd", img, cs, float(sgn), B, Cc, H, W, out)
        ctx.save_for_backward(img, cs)
        ctx.sgn = float(sgn)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        img, cs = ctx.saved_tensors
        B, Cc, H, W = img.shape
        gout = _c(gout)
        gimg = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        gcs = torch.empty_like(cs) if ctx.needs_input_grad[1] else None
        if gimg is not None or gcs is not None:
            call("livae_rot_sample_bwd", img, cs, ctx.sgn, gout, B, Cc, H, W, gimg, gcs)
        return gimg, gcs, None


def rot_sample(img, cs, sgn=1.0):
    """F.grid_sample(img, F.affine_grid([[c,-sgn*s,0],[sgn*s,c,0]], img.size(), align_corners=False),
    padding_mode='reflection', align_corners=False) for NCHW img (model.py:250-258, 465-470;
    train.py:675-677); sgn=-1 applies the inverse rotation."""
    return _RotSampleFn.apply(img, cs, sgn)


# ------------------------------------------------------------------------------------------
# reparameterisation and losses
# ------------------------------------------------------------------------------------------
class _ReparamFn(Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu = _c(mu); logvar = _c(logvar); eps = _c(eps)
        require_cuda(mu, logvar, eps)
        z = torch.empty_like(mu)
        call("livae_reparam_fwd", mu, logvar, eps, mu.numel(), z)
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, gz):
        logvar, eps = ctx.saved_tensors
        gz = _c(gz)
        gmu = torch.empty_like(gz)
        glv = torch.empty_like(gz)
        call("livae_reparam_bwd", gz, logvar, eps, gz.numel(), gmu, glv)
        return gmu, glv, None


def reparam(mu, logvar, eps):
    """z = mu + eps * exp(0.5 * logvar) (model.py:436-439) with eps drawn by the caller"""
    return _ReparamFn.apply(mu, logvar, eps)


class _ElboSumsFn(Function):
    @staticmethod
    def forward(ctx, recon, x, mu, logvar):
        recon = _c(recon); x = _c(x)
        n_lat = 0
        if mu is not None:
            mu = _c(mu); logvar = _c(logvar)        # heads of foreign models may be strided slices of one tensor
            n_lat = mu.numel()
        require_cuda(recon, x, mu, logvar)
        assert recon.numel() == x.numel()
        sums = torch.empty(2, dtype=torch.float32, device=x.device)
        scratch = _get_scratch(x.device, "elbo", L.lib().livae_elbo_scratch_floats())
        call("livae_elbo_fwd", recon, x, recon.numel(), mu, logvar, n_lat, sums, scratch)
        ctx.save_for_backward(recon, x, mu, logvar)
        return sums

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        recon, x, mu, logvar = ctx.saved_tensors
        g = _c(g)
        d_recon = torch.empty_like(recon) if ctx.needs_input_grad[0] else None
        d_x = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        n_lat = 0 if mu is None else mu.numel()
        d_mu = torch.empty_like(mu) if n_lat else None
        d_lv = torch.empty_like(logvar) if n_lat else None
        call("livae_elbo_bwd", recon, x, recon.numel(), mu, logvar, n_lat, g, d_recon, d_x, d_mu, d_lv)
        return d_recon, d_x, d_mu, d_lv


def elbo_sums(recon, x, mu=None, logvar=None):
    """-> tensor[2] = (sum((recon-x)^2), sum(-0.5*(1+logvar-mu^2-exp(logvar)))) in one launch
    (loss.py:116-119, 165-169; train.py:391-393).  With mu=None only the first entry is meaningful."""
    return _ElboSumsFn.apply(recon, x, mu, logvar)


class _CycleFn(Function):
    @staticmethod
    def forward(ctx, theta, theta_rot, angle):
        theta = _c(theta); theta_rot = _c(theta_rot); angle = _c(angle)
        require_cuda(theta, theta_rot, angle)
        B = theta.numel()
        assert theta_rot.numel() == B and angle.numel() == B
        loss = torch.empty((), dtype=torch.float32, device=theta.device)
        call("livae_cycle_fwd", theta, theta_rot, angle, B, loss)
        ctx.save_for_backward(theta, theta_rot, angle)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        theta, theta_rot, angle = ctx.saved_tensors
        g = _c(g)
        d_t = torch.empty_like(theta) if ctx.needs_input_grad[0] else None
        d_r = torch.empty_like(theta_rot) if ctx.needs_input_grad[1] else None
        if d_t is not None or d_r is not None:
            call("livae_cycle_bwd", theta, theta_rot, angle, g, theta.numel(), d_t, d_r)
        return d_t, d_r, None


def cycle_loss(theta, theta_rot, angle):
    """mean(1 - cos(theta_rot - theta + angle)) (loss.py:52-94)"""
    return _CycleFn.apply(theta, theta_rot, angle)


# ------------------------------------------------------------------------------------------
# data side
# ------------------------------------------------------------------------------------------
def patch_gather(images, sites, P, out=None):
    """images [n_img,H,W] float32/float64 (device), sites int32 [N,3] (img, cy, cx) ->
    float32 [N,1,P,P] == float32(img)[cy-P/2:cy+P/2, cx-P/2:cx+P/2] (data.py:211-250, transform=None)"""
    if not images.is_cuda or not sites.is_cuda:
        raise RuntimeError("livae.patch_gather: device tensors required; there is no CPU path")
    assert images.dim() == 3 and images.is_contiguous()
    assert sites.dtype == torch.int32 and sites.dim() == 2 and sites.shape[1] == 3 and sites.is_contiguous()
    n_img, H, W = images.shape
    N = sites.shape[0]
    if out is None:
        out = torch.empty((N, 1, P, P), dtype=torch.float32, device=images.device)
    if images.dtype == torch.float32:
        call("livae_patch_gather_f32", images, n_img, H, W, sites, N, P, out)
    elif images.dtype == torch.float64:
        call("livae_patch_gather_f64", images, n_img, H, W, sites, N, P, out)
    else:
        raise RuntimeError(f"patch_gather: unsupported image dtype {images.dtype}")
    return out


def patch_gather_subpixel(images, img_idx, yx, P, out=None):
    """images [n_img,H,W] float32/float64 (device), img_idx int32 [N], yx float64 [N,2] float sites (cy, cx) ->
    float32 [N,1,P,P]: the sub-pixel crop of AdaptiveLatticeDataset.__getitem__ (data.py:478-551, transform=None)
    before its min-max (apply patch_minmax_ for data.py:553-558)"""
    if not images.is_cuda or not yx.is_cuda or not img_idx.is_cuda:
        raise RuntimeError("livae.patch_gather_subpixel: device tensors required; there is no CPU path")
    assert images.dim() == 3 and images.is_contiguous()
    assert img_idx.dtype == torch.int32 and img_idx.is_contiguous()
    assert yx.dtype == torch.float64 and yx.dim() == 2 and yx.shape[1] == 2 and yx.is_contiguous()
    n_img, H, W = images.shape
    N = yx.shape[0]
    assert img_idx.numel() == N
    if out is None:
        out = torch.empty((N, 1, P, P), dtype=torch.float32, device=images.device)
    if images.dtype == torch.float32:
        call("livae_patch_gather_subpixel_f32", images, n_img, H, W, img_idx, yx, N, P, out)
    elif images.dtype == torch.float64:
        call("livae_patch_gather_subpixel_f64", images, n_img, H, W, img_idx, yx, N, P, out)
    else:
        raise RuntimeError(f"patch_gather_subpixel: unsupported image dtype {images.dtype}")
    return out


def patch_gather_roi(images, img_idx, yx, S, roi, out=None):
    """the [N,1,S,S] `patch_big` of Adaptive/PairedAdaptiveLatticeDataset.__getitem__ (data.py:496-546): as
    patch_gather_subpixel but reading only the reference's integer ROI window of `roi` pixels around round(site)"""
    if not images.is_cuda or not yx.is_cuda or not img_idx.is_cuda:
        raise RuntimeError("livae.patch_gather_roi: device tensors required; there is no CPU path")
    assert images.dim() == 3 and images.is_contiguous()
    assert img_idx.dtype == torch.int32 and img_idx.is_contiguous()
    assert yx.dtype == torch.float64 and yx.dim() == 2 and yx.shape[1] == 2 and yx.is_contiguous()
    n_img, H, W = images.shape
    N = yx.shape[0]
    assert img_idx.numel() == N
    if out is None:
        out = torch.empty((N, 1, S, S), dtype=torch.float32, device=images.device)
    if images.dtype == torch.float32:
        call("livae_patch_gather_roi_f32", images, n_img, H, W, img_idx, yx, N, S, roi, out)
    elif images.dtype == torch.float64:
        call("livae_patch_gather_roi_f64", images, n_img, H, W, img_idx, yx, N, S, roi, out)
    else:
        raise RuntimeError(f"patch_gather_roi: unsupported image dtype {images.dtype}")
    return out


def augment(patches, scale, flags, shift, out=None):
    """default_transform(rotation=False) with given draws (data.py:78-116): patches [N,1,S,S] fp32, scale fp32 [N],
    flags int32 [N] (bit0 hflip, bit1 vflip), shift int32 [N,2] (shift_y, shift_x) -> [N,1,S,S]"""
    require_cuda(patches, scale)
    N, S = patches.shape[0], patches.shape[-1]
    assert patches.numel() == N * S * S and scale.numel() == N
    assert flags.dtype == torch.int32 and flags.numel() == N and flags.is_cuda and flags.is_contiguous()
    assert shift.dtype == torch.int32 and shift.numel() == 2 * N and shift.is_cuda and shift.is_contiguous()
    if out is None:
        out = torch.empty_like(patches)
    call("livae_augment", patches, N, S, scale, flags, shift, out)
    return out


def rotate_crop(patches, P, angle_deg=None, normalise=False, out=None):
    """TF.rotate(angle_deg, bilinear, fill=0) (or no rotation when angle_deg is None) -> TF.center_crop(P) ->
    optional per-patch min-max (data.py:698-730).  patches [N,1,S,S] fp32, angle_deg float64 [N] -> [N,1,P,P]"""
    require_cuda(patches)
    N, S = patches.shape[0], patches.shape[-1]
    assert patches.numel() == N * S * S
    if angle_deg is not None:
        assert angle_deg.dtype == torch.float64 and angle_deg.numel() == N and angle_deg.is_cuda
    if out is None:
        out = torch.empty((N, 1, P, P), dtype=torch.float32, device=patches.device)
    call("livae_rotate_crop", patches, N, S, P, angle_deg, 0 if angle_deg is None else 1, int(normalise), out)
    return out


def patch_minmax_(patches):
    """in-place per-patch min-max normalisation to [0,1] (data.py:553-558)"""
    require_cuda(patches)
    N = patches.shape[0]
    P = patches.shape[-1]
    assert patches.numel() == N * P * P
    call("livae_patch_minmax", patches, N, P)
    return patches


# ------------------------------------------------------------------------------------------
# optimiser side
# ------------------------------------------------------------------------------------------
def l2norm_clip_(flat_grads, max_norm, apply=True):
    """-> tensor[2] (norm, clip coefficient); scales flat_grads in place when apply (train.py:396)"""
    require_cuda(flat_grads)
    out = torch.empty(2, dtype=torch.float32, device=flat_grads.device)
    scratch = _get_scratch(flat_grads.device, "l2", L.lib().livae_l2norm_scratch_floats())
    call("livae_l2norm_clip", flat_grads, flat_grads.numel(), float(max_norm), out, scratch, int(apply))
    return out


def adamw_(p, g, m, v, step_dev, lr, betas, eps, weight_decay, decoupled=True, gscale=None, inc_step=True):
    require_cuda(p, g, m, v, step_dev)
    call("livae_adamw", p, g, m, v, p.numel(), float(lr), float(betas[0]), float(betas[1]), float(eps),
         float(weight_decay), int(decoupled), step_dev, gscale, int(inc_step))


# ------------------------------------------------------------------------------------------
# engine 1: tcgen05 / TMA convolution (bf16 operands, fp32 accumulate)
# ------------------------------------------------------------------------------------------
def cast(t, dtype):
    """fp32 <-> bf16 conversion kernel"""
    t = _c(t)
    if t.dtype == dtype:
        return t
    codes = {torch.float32: L.F32, torch.bfloat16: L.BF16}
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    call("livae_cast", t, codes[t.dtype], out, codes[dtype], t.numel())
    return out


# ---- packed-weight cache ---------------------------------------------------------------------------------------
# The bf16 re-layouts of a parameter depend on the parameter alone, which changes once per optimiser step, while a
# training step asks for the same pack several times (model(x) and model.encoder(x_rot) share every encoder weight).
# An entry is reused only for the SAME tensor object (weak reference), at the same storage address and autograd
# version, within the same optimiser epoch: `WEIGHT_EPOCH` is bumped by every torch optimiser step (global post-step
# hook) and by FlatAdamW.step (which updates the flat buffer from a raw kernel, invisible to the version counter).
# In-place edits through `.data` are invisible to all of these: call `invalidate_weight_packs()` after such an edit.
_PACKS: dict = {}
WEIGHT_EPOCH = [0]


def invalidate_weight_packs() -> None:
    WEIGHT_EPOCH[0] += 1
    if len(_PACKS) > 256:
        _PACKS.clear()


from torch.optim.optimizer import register_optimizer_step_post_hook as _post_step_hook  # noqa: E402

_post_step_hook(lambda *a, **k: invalidate_weight_packs())


def _packed(w, key, make):
    k = (id(w),) + key
    ent = _PACKS.get(k)
    if (ent is not None and ent[0]() is w and ent[1] == w._version and ent[2] == WEIGHT_EPOCH[0]
            and ent[3] == w.data_ptr()):
        return ent[4]
    out = make()
    try:
        ref = weakref.ref(w, lambda _r, k=k, d=_PACKS: d.pop(k, None))
    except TypeError:
        return out
    _PACKS[k] = (ref, w._version, WEIGHT_EPOCH[0], w.data_ptr(), out)
    return out


def tc_pack_weights(w, Cs, Cb, kh, kw, mode, Cs_pad=None):
    """cached (see above) bf16 re-layout of a torch-layout weight for the tensor-core engine"""
    return _packed(w, ("pack", Cs, Cb, kh, kw, mode, Cs_pad), lambda: _tc_pack_weights(w, Cs, Cb, kh, kw, mode, Cs_pad))


def _tc_pack_weights(w, Cs, Cb, kh, kw, mode, Cs_pad=None):
    """torch-layout fp32 weight [Cs,Cb,kh,kw] -> bf16 packed for the tensor-core engine:
    mode 0 [tap][Cs][Cb] (forward), 1 [flipped tap][Cb][Cs], 2 [tap][Cb][Cs] (data gradient),
    3 [Cs_pad][tap][Cb] (Linear forward, K in (h,w,c) order), 4 [tap][Cb][Cs_pad] (Linear data gradient)"""
    w = _c(w)
    assert w.numel() == Cs * Cb * kh * kw and w.dtype == torch.float32 and w.is_cuda
    Cp = Cs if Cs_pad is None else Cs_pad
    assert Cp >= Cs and (Cp == Cs or mode >= 3)
    taps = kh * kw
    shape = {0: (taps, Cp, Cb), 1: (taps, Cb, Cp), 2: (taps, Cb, Cp), 3: (Cp, taps, Cb), 4: (taps, Cb, Cp)}[mode]
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    call("livae_tc_pack_weights", w, Cp, Cb, kh, kw, mode, Cs, out)
    return out


def tc_conv_supported(B, Hin, Win, Cin, Cout, kh, kw, stride, pad):
    d = L.TcConvDesc(B, Hin, Win, Cin, Cout, kh, kw, stride, pad, 0, 0)
    return bool(L.lib().livae_tc_conv_supported(C.byref(d)))


def tc_conv(x, wpacked, bias, kh, kw, stride, pad, act=ACT_NONE, out_f32=False, relu_mask=None):
    """raw tensor-core convolution (no autograd): x bf16 [B,H,W,Cin], wpacked bf16 [taps,Cout,Cin]"""
    assert x.dtype == torch.bfloat16 and wpacked.dtype == torch.bfloat16 and x.is_cuda and x.is_contiguous()
    B, Hin, Win, Cin = x.shape
    taps, Cout, Cin2 = wpacked.shape
    assert Cin2 == Cin and taps == kh * kw
    d = L.TcConvDesc(B, Hin, Win, Cin, Cout, kh, kw, stride, pad, act, int(out_f32))
    Ho = (Hin + 2 * pad - kh) // stride + 1
    Wo = (Win + 2 * pad - kw) // stride + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    if relu_mask is not None:
        assert relu_mask.dtype == torch.bfloat16 and relu_mask.shape == y.shape and relu_mask.is_contiguous()
    call("livae_tc_conv", C.byref(d), x, wpacked, bias, y, relu_mask)
    return y


def tc_conv_dgrad(gy, wpacked2, bias, Hin, Win, kh, kw, stride, pad, act=ACT_NONE, out_f32=False, relu_mask=None):
    """raw tensor-core data gradient of conv(Cin->Cout) / ConvTranspose2d forward: gy bf16 [B,Ho,Wo,Cout],
    wpacked2 bf16 [taps,Cin,Cout] (mode-2 packing) -> gx [B,Hin,Win,Cin]"""
    assert gy.dtype == torch.bfloat16 and wpacked2.dtype == torch.bfloat16 and gy.is_cuda and gy.is_contiguous()
    B, Ho, Wo, Cout = gy.shape
    taps, Cin, Cout2 = wpacked2.shape
    assert Cout2 == Cout and taps == kh * kw
    d = L.TcConvDesc(B, Hin, Win, Cin, Cout, kh, kw, stride, pad, act, int(out_f32))
    gx = torch.empty((B, Hin, Win, Cin), dtype=torch.float32 if out_f32 else torch.bfloat16, device=gy.device)
    if relu_mask is not None:
        assert relu_mask.dtype == torch.bfloat16 and relu_mask.shape == gx.shape and relu_mask.is_contiguous()
    call("livae_tc_conv_dgrad", C.byref(d), gy, wpacked2, bias, gx, relu_mask)
    return gx


def tc_conv_wgrad(x, gy, kh, kw, stride, pad, want_bias=True):
    """raw tensor-core weight gradient: x bf16 [B,Hin,Win,Cin], gy bf16 [B,Ho,Wo,Cout] (pre-activation
    gradient) -> gw fp32 [Cout,Cin,kh,kw], gb fp32 [Cout] or None"""
    assert x.dtype == torch.bfloat16 and gy.dtype == torch.bfloat16 and x.is_contiguous() and gy.is_contiguous()
    B, Hin, Win, Cin = x.shape
    Cout = gy.shape[-1]
    d = L.TcConvDesc(B, Hin, Win, Cin, Cout, kh, kw, stride, pad, 0, 0)
    gw = torch.empty((Cout, Cin, kh, kw), dtype=torch.float32, device=x.device)
    gb = torch.empty(Cout, dtype=torch.float32, device=x.device) if want_bias else None
    ws = torch.empty(L.lib().livae_tc_wgrad_ws_bytes(C.byref(d)) // 4, dtype=torch.float32, device=x.device)
    call("livae_tc_conv_wgrad", C.byref(d), x, gy, gw, gb, ws)
    return gw, gb


# ---- conv 5x5 p2 + ReLU + MaxPool2 in space-to-depth form (csrc/conv_s2d.cu; STN conv2, model.py:207-209)
def conv5pool_supported(B, H, W, Ci, Co):
    return bool(L.lib().livae_tc_conv5pool_supported(B, H, W, Ci, Co))


def conv5pool_pack(w, mode):
    return _packed(w, ("c5p", mode), lambda: _conv5pool_pack(w, mode))


def _conv5pool_pack(w, mode):
    Co, Ci = w.shape[0], w.shape[1]
    shape = (9, 4 * Co, 4 * Ci) if mode == 0 else (9, 4 * Ci, 4 * Co)
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    call("livae_tc_conv5pool_pack", _c(w), Co, Ci, mode, out)
    return out


def conv5pool_fwd(x, w, bias):
    """x bf16 [B,H,W,Ci], w fp32 [Co,Ci,5,5] -> (pooled bf16 [B,H/2,W/2,Co], idx uint8)"""
    B, H, W, Ci = x.shape
    Co = w.shape[0]
    y = torch.empty((B, H // 2, W // 2, Co), dtype=torch.bfloat16, device=x.device)
    idx = torch.empty((B, H // 2, W // 2, Co), dtype=torch.uint8, device=x.device)
    call("livae_tc_conv5pool_fwd", x, conv5pool_pack(w, 0), bias, B, H, W, Ci, Co, y, idx)
    return y, idx


def conv5pool_bwd(x, w, g_pooled, idx, want_dgrad=True):
    """x bf16 [B,H,W,Ci] (the layer input = ReLU output of the layer below, used as its mask), g_pooled bf16
    PRE-activation gradient [B,H/2,W/2,Co] -> (gw fp32 [Co,Ci,5,5], gb fp32 [Co], gx bf16 [B,H,W,Ci] masked by x > 0)"""
    B, H, W, Ci = x.shape
    Co = w.shape[0]
    g4 = torch.empty((B, H // 2, W // 2, 4 * Co), dtype=torch.bfloat16, device=x.device)
    call("livae_unpool_s2d_bf16", g_pooled, idx, B, H // 2, W // 2, Co, g4)
    gw = torch.empty_like(w)
    gb = torch.empty(Co, dtype=torch.float32, device=x.device)
    ws = torch.empty(L.lib().livae_tc_conv5pool_wgrad_ws_bytes(Ci, Co) // 4, dtype=torch.float32, device=x.device)
    # bias gradient from the POOLED gradient (a quarter of the elements of its routed 4-phase form)
    call("livae_colsum_bf16", g_pooled, B * (H // 2) * (W // 2), Co, gb)
    call("livae_tc_conv5pool_wgrad", x, g4, B, H, W, Ci, Co, gw, None, ws)
    gx = None
    if want_dgrad:
        gx = torch.empty_like(x)
        call("livae_tc_conv5pool_dgrad", g4, conv5pool_pack(w, 1), x, B, H, W, Ci, Co, gx)
    return gw, gb, gx


def dgrad_s2blk_supported(Hin, Win, Cin, Cout):
    return bool(L.lib().livae_tc_dgrad_s2blk_supported(Hin, Win, Cin, Cout))


def dgrad_s2blk(gy, w, Hin, Win, relu_mask=None):
    """data gradient of conv(Cin->Cout, 4x4, stride 2, pad 1): gy bf16 [B,Hin/2,Win/2,Cout], w fp32 [Cout,Cin,4,4]
    -> gx bf16 [B,Hin,Win,Cin] (times relu_mask > 0); one launch in block form (csrc/conv_s2d.cu)"""
    B = gy.shape[0]
    Cout, Cin = w.shape[0], w.shape[1]

    def make():
        t = torch.empty((9, 4 * Cin, Cout), dtype=torch.bfloat16, device=gy.device)
        call("livae_tc_dgrad_s2blk_pack", _c(w), Cout, Cin, t)
        return t

    wblk = _packed(w, ("s2blk",), make)
    gx = torch.empty((B, Hin, Win, Cin), dtype=torch.bfloat16, device=gy.device)
    call("livae_tc_dgrad_s2blk", gy, wblk, relu_mask, B, Hin, Win, Cin, Cout, gx)
    return gx


# ---- decoder blocks d1-d3 phase-folded onto the low-resolution input (csrc/upfold.cu; model.py:356-368)
def upfold_supported(B, h, w, Cin, Cout):
    return bool(L.lib().livae_upfold_supported(B, h, w, Cin, Cout))


def _upfold_pack(w):
    Cout, Cin = w.shape[0], w.shape[1]
    wf = torch.empty((9, 4 * Cout, Cin), dtype=torch.bfloat16, device=w.device)
    wd = torch.empty((9, Cin, 4 * Cout), dtype=torch.bfloat16, device=w.device)
    call("livae_upfold_pack", _c(w), Cout, Cin, wf, wd)
    return wf, wd


def upfold_pack(w):
    """(forward, data-gradient) folded weight packs of a [Cout,Cin,3,3] weight, cached per optimiser epoch"""
    return _packed(w, ("upfold",), lambda: _upfold_pack(w))


def _strip_packs(w, mode):
    """mode-0 / mode-2 packs of the layer's own weight (top / bottom strips) and of its (ky,kx)-transpose (the left /
    right strips are stored transposed)"""
    Cout, Cin = w.shape[0], w.shape[1]
    wt = _packed(w, ("upfold_wT",), lambda: w.detach().transpose(2, 3).contiguous())
    return (tc_pack_weights(w, Cout, Cin, 3, 3, mode),
            _packed(w, ("upfold_wT_pack", mode), lambda: _tc_pack_weights(wt, Cout, Cin, 3, 3, mode)))


def upfold_fwd(x, w, bias):
    """x bf16 [B,h,w,Cin] -> y bf16 [B,2h,2w,Cout] = ReLU(Conv3x3(ReflectionPad(Upsample2(x))) + bias); also returns the
    border strips (needed again by the weight gradient).  Every batch of strips [2B,4,L,Cin] is handed to the
    convolution kernels as ONE tall image [1,8B,L,Cin] (pad 0): the rows between two strips' outputs are slack."""
    assert x.dtype == torch.bfloat16 and x.is_cuda and x.is_contiguous()
    B, h, ww, Cin = x.shape
    Cout = w.shape[0]
    dev = x.device
    s_tb = torch.empty((2 * B, 4, 2 * ww + 2, Cin), dtype=torch.bfloat16, device=dev)
    s_lr = torch.empty((2 * B, 4, 2 * h + 2, Cin), dtype=torch.bfloat16, device=dev)
    call("livae_upfold_strips", x, B, h, ww, Cin, s_tb, s_lr)
    p_tb, p_lr = _strip_packs(w, 0)
    corr_tb = tc_conv(s_tb.view(1, 8 * B, 2 * ww + 2, Cin), p_tb, None, 3, 3, 1, 0, ACT_NONE, out_f32=True)
    corr_lr = tc_conv(s_lr.view(1, 8 * B, 2 * h + 2, Cin), p_lr, None, 3, 3, 1, 0, ACT_NONE, out_f32=True)
    y = torch.empty((B, 2 * h, 2 * ww, Cout), dtype=torch.bfloat16, device=dev)
    call("livae_upfold_fwd", x, upfold_pack(w)[0], bias, B, h, ww, Cin, Cout, y)
    call("livae_upfold_ring", corr_tb, corr_lr, B, h, ww, Cout, y)
    return y, s_tb, s_lr


def upfold_bwd(x, w, gz, s_tb, s_lr, want_dgrad=True):
    """x bf16 [B,h,w,Cin] (the layer input, also the ReLU mask of the layer below), gz bf16 [B,2h,2w,Cout]
    (pre-activation gradient) -> (gw fp32 [Cout,Cin,3,3], gx bf16 [B,h,w,Cin] masked by x > 0 or None)"""
    assert gz.dtype == torch.bfloat16 and gz.is_contiguous()
    B, h, ww, Cin = x.shape
    Cout = w.shape[0]
    dev = x.device
    g_tb = torch.empty((2 * B, 4, 2 * ww, Cout), dtype=torch.bfloat16, device=dev)
    g_lr = torch.empty((2 * B, 4, 2 * h, Cout), dtype=torch.bfloat16, device=dev)
    call("livae_upfold_gather", gz, B, h, ww, Cout, g_tb, g_lr)
    # tall images: x side [1,8B,L,Cin], gradient side [1,8B-2,L-2,Cout] (its slack rows are zero)
    t_tb = g_tb.view(8 * B, 2 * ww, Cout)[:8 * B - 2].unsqueeze(0)
    t_lr = g_lr.view(8 * B, 2 * h, Cout)[:8 * B - 2].unsqueeze(0)
    gw_tb, _ = tc_conv_wgrad(s_tb.view(1, 8 * B, 2 * ww + 2, Cin), t_tb, 3, 3, 1, 0, want_bias=False)
    gw_lr, _ = tc_conv_wgrad(s_lr.view(1, 8 * B, 2 * h + 2, Cin), t_lr, 3, 3, 1, 0, want_bias=False)
    gw = torch.empty_like(w)
    ws = torch.empty(L.lib().livae_upfold_wgrad_ws_bytes(Cin, Cout) // 4, dtype=torch.float32, device=dev)
    call("livae_upfold_wgrad", x, gz, gw_tb, gw_lr, B, h, ww, Cin, Cout, gw, ws)
    gx = None
    if want_dgrad:
        gx = torch.empty_like(x)
        call("livae_upfold_dgrad", gz, upfold_pack(w)[1], x, B, h, ww, Cin, Cout, gx)
        p_tb, p_lr = _strip_packs(w, 2)
        gs_tb = tc_conv_dgrad(t_tb, p_tb, None, 8 * B, 2 * ww + 2, 3, 3, 1, 0, out_f32=True)
        gs_lr = tc_conv_dgrad(t_lr, p_lr, None, 8 * B, 2 * h + 2, 3, 3, 1, 0, out_f32=True)
        call("livae_upfold_patch", gs_tb, gs_lr, x, B, h, ww, Cin, gx)
    return gw, gx

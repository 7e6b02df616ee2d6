"""ctypes binding of liblivae_sm100.so (C ABI declared in include/livae_b200.h).

There is NO fallback: if the shared library is missing, cannot be loaded, or the device is not
sm_100, every op raises.  Tensors are passed as raw device pointers; every kernel is enqueued on
torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblivae_sm100.so")

_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float


class ConvDesc(C.Structure):
    """mirror of livae_conv_desc"""
    _fields_ = [(n, C.c_int) for n in ("kind", "B", "Hin", "Win", "Cin", "Cout", "kh", "kw", "stride",
                                       "pad", "act", "pool")]


class TcConvDesc(C.Structure):
    """mirror of livae_tc_conv_desc"""
    _fields_ = [(n, C.c_int) for n in ("B", "Hin", "Win", "Cin", "Cout", "kh", "kw", "stride", "pad", "act",
                                       "out_f32")]


CONV, CONVT = 0, 2
F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2

# name -> argument type codes: p pointer, i int, l int64, f float, s stream, d ConvDesc*
_SIGS = {
    "livae_patch_gather_f32": "piiipiips",
    "livae_patch_gather_f64": "piiipiips",
    "livae_patch_minmax": "piis",
    "livae_patch_gather_subpixel_f32": "piiippiips",
    "livae_patch_gather_subpixel_f64": "piiippiips",
    "livae_patch_gather_roi_f32": "piiippiiips",
    "livae_patch_gather_roi_f64": "piiippiiips",
    "livae_augment": "piipppps",
    "livae_rotate_crop": "piiipiips",
    "livae_rot_sample_fwd": "ppfiiiips",
    "livae_rot_sample_bwd": "ppfpiiiipps",
    "livae_stn_head_fwd": "pipps",
    "livae_stn_head_bwd": "pppips",
    "livae_stn_tail_fwd": "pppiippps",
    "livae_stn_tail_bwd": "pppppiippps",
    "livae_angle_to_cs": "pips",
    "livae_angle_to_cs_bwd": "ppips",
    "livae_reparam_fwd": "pppips",
    "livae_reparam_bwd": "pppipps",
    "livae_elbo_fwd": "pplppipps",
    "livae_elbo_bwd": "pplppi" + "ppppp" + "s",
    "livae_cycle_fwd": "pppips",
    "livae_cycle_bwd": "ppppipps",
    "livae_axpby_dev": "pppplps",
    "livae_conv_fwd": "d" + "p" * 6 + "s",
    "livae_conv_bwd": "d" + "p" * 8 + "s",
    "livae_tc_pack_weights": "piiiiiips",
    "livae_permute_linear_grad": "piiips",
    "livae_linear_dgrad": "pppiiips",
    "livae_thin_conv1c_fwd": "ipppiiipps",
    "livae_thin_conv1c_wgrad": "ipppiiipps",
    "livae_thin_conv1c_dgrad": "ppiiips",
    "livae_thin_convc1_fwd": "pppiiiips",
    "livae_thin_convc1_wgrad": "ppiiipps",
    "livae_upconv_c1_bwd": "pppiiipppps",
    "livae_upconv_c1_fwd": "pppiiiips",
    "livae_sigmoid_bwd": "ppplps",
    "livae_relu_mask_cast_bf16": "pplps",
    "livae_maxpool_bf16": "piiiipps",
    "livae_unpool_bf16": "ppiiiips",
    "livae_upsample_pad_fwd_bf16": "piiiips",
    "livae_upsample_pad_bwd_bf16": "piiiipps",
    "livae_upsample_pad_bwd_bias_bf16": "piiiippps",
    "livae_colsum_bf16": "plips",
    "livae_ssim_box": "ppliiiffpps",
    "livae_thin_convt_c1_fwd": "pppiiiiips",
    "livae_thin_convt_c1_dgrad": "pppiiiips",
    "livae_thin_convt_c1_wgrad": "ppiiiipps",
    "livae_decfc_fwd_bf16": "pppiiiips",
    "livae_decfc_bwd_bf16": "pppiiiippps",
    "livae_tc_conv": "t" + "p" * 5 + "s",
    "livae_tc_conv_dgrad": "t" + "p" * 5 + "s",
    "livae_tc_conv_wgrad": "t" + "p" * 5 + "s",
    "livae_cast": "pipils",
    "livae_tc_dgrad_s2blk_pack": "piips",
    "livae_tc_dgrad_s2blk": "pppiiiiips",
    "livae_tc_conv5pool_pack": "piiips",
    "livae_tc_conv5pool_fwd": "pppiiiiipps",
    "livae_unpool_s2d_bf16": "ppiiiips",
    "livae_tc_conv5pool_dgrad": "pppiiiiips",
    "livae_tc_conv5pool_wgrad": "ppiiiiippps",
    "livae_upfold_pack": "piipps",
    "livae_upfold_strips": "piiiipps",
    "livae_upfold_fwd": "pppiiiiips",
    "livae_upfold_ring": "ppiiiips",
    "livae_upfold_gather": "piiiipps",
    "livae_upfold_dgrad": "pppiiiiips",
    "livae_upfold_patch": "pppiiiips",
    "livae_upfold_wgrad": "ppppiiiiipps",
    "livae_upsample_pad_fwd": "piiiips",
    "livae_upsample_pad_bwd": "piiiipps",
    "livae_decfc_fwd": "pppiiiips",
    "livae_decfc_bwd": "ppppiiiippps",
    "livae_l2norm_clip": "plfppis",
    "livae_adamw": "pppplfffffippis",
}
_CODE = {"p": _P, "i": _I, "l": _L, "f": _F, "s": _P, "d": C.POINTER(ConvDesc), "t": C.POINTER(TcConvDesc)}

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python li-vae_b200/build.py` "
            "(livae has no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    L.livae_last_error.restype = C.c_char_p
    L.livae_abi_version.restype = C.c_int
    L.livae_device_ok.restype = C.c_int
    for n in ("livae_elbo_scratch_floats", "livae_l2norm_scratch_floats", "livae_launch_count"):
        getattr(L, n).restype = C.c_int64
        getattr(L, n).argtypes = []
    L.livae_conv_fwd_ws_bytes.restype = C.c_int64
    L.livae_conv_fwd_ws_bytes.argtypes = [C.POINTER(ConvDesc)]
    L.livae_tc_wgrad_ws_bytes.restype = C.c_int64
    L.livae_tc_wgrad_ws_bytes.argtypes = [C.POINTER(TcConvDesc)]
    L.livae_tc_set_halo_mode.restype = None
    L.livae_tc_set_halo_mode.argtypes = [C.c_int]
    L.livae_tc_set_wgrad_halo.restype = None
    L.livae_tc_set_wgrad_halo.argtypes = [C.c_int]
    L.livae_set_probe.restype = None
    L.livae_set_probe.argtypes = [C.c_void_p]
    L.livae_thin_set_tc.restype = None
    L.livae_thin_set_tc.argtypes = [C.c_int]
    L.livae_tc_dgrad_s2blk_supported.restype = C.c_int
    L.livae_tc_dgrad_s2blk_supported.argtypes = [C.c_int] * 4
    L.livae_tc_conv5pool_supported.restype = C.c_int
    L.livae_tc_conv5pool_supported.argtypes = [C.c_int] * 5
    L.livae_tc_conv5pool_wgrad_ws_bytes.restype = C.c_int64
    L.livae_tc_conv5pool_wgrad_ws_bytes.argtypes = [C.c_int, C.c_int]
    L.livae_tc_halo_geometry.restype = C.c_int
    L.livae_tc_halo_geometry.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 4
    L.livae_upfold_supported.restype = C.c_int
    L.livae_upfold_supported.argtypes = [C.c_int] * 5
    L.livae_upfold_wgrad_ws_bytes.restype = C.c_int64
    L.livae_upfold_wgrad_ws_bytes.argtypes = [C.c_int, C.c_int]
    L.livae_tc_conv_supported.restype = C.c_int
    L.livae_tc_conv_supported.argtypes = [C.POINTER(TcConvDesc)]
    L.livae_ssim_box_ws_floats.restype = C.c_int64
    L.livae_ssim_box_ws_floats.argtypes = [C.c_int64, C.c_int]
    L.livae_set_scratch.restype = C.c_int
    L.livae_set_scratch.argtypes = [C.c_void_p, C.c_int64]
    L.livae_conv_out_shape.restype = None
    L.livae_conv_out_shape.argtypes = [C.POINTER(ConvDesc), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for name, sig in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [_CODE[c] for c in sig]
    if os.environ.get("LIVAE_HALO16"):      # A/B switch: 16-column halo boxes only (csrc/conv_tc.cu launch_conv_tc_halo)
        L.livae_tc_set_halo_mode(2)
    _lib = L
    return L


def exported_symbols():
    """every entry point include/livae_b200.h declares (used by the CPU symbol test)"""
    return sorted(list(_SIGS) + ["livae_last_error", "livae_abi_version", "livae_device_ok",
                                 "livae_elbo_scratch_floats", "livae_l2norm_scratch_floats",
                                 "livae_launch_count", "livae_tc_conv_supported", "livae_tc_set_halo_mode", "livae_tc_set_wgrad_halo", "livae_set_scratch", "livae_thin_set_tc", "livae_set_probe", "livae_tc_conv5pool_supported", "livae_tc_halo_geometry", "livae_upfold_supported", "livae_upfold_wgrad_ws_bytes", "livae_tc_dgrad_s2blk_supported", "livae_tc_conv5pool_wgrad_ws_bytes", "livae_tc_wgrad_ws_bytes", "livae_conv_fwd_ws_bytes", "livae_conv_out_shape", "livae_ssim_box_ws_floats"])


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def stream(device=None):
    return torch.cuda.current_stream(device).cuda_stream


# Per-device scratch the library parks per-CTA partial sums in, so that reductions split over CTAs are finished in
# a fixed order (bit-reproducible forward pass and weight gradients; see csrc/common.cuh scratch_floats).  Owned
# here (a torch allocation, kept alive for the life of the process), registered with livae_set_scratch on first use.
SCRATCH_BYTES = 96 << 20
_scratch = {}


def _ensure_scratch(dev):
    if dev.index in _scratch:
        return
    with torch.cuda.device(dev):
        buf = torch.empty(SCRATCH_BYTES, dtype=torch.uint8, device=dev)
        if lib().livae_set_scratch(buf.data_ptr(), SCRATCH_BYTES) != 0:
            raise RuntimeError("livae_set_scratch failed: " + lib().livae_last_error().decode(errors="replace"))
    _scratch[dev.index] = buf


# When set to a list, every call is bracketed by CUDA events on the launching stream and
# (name, args, start_event, end_event) is appended: bench.py uses it to time each kernel family.
PROFILE = None


def call(name, *args):
    """invoke an int-returning entry point; tensors -> device pointers; raises on error"""
    L = lib()
    conv = []
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor):
            conv.append(a.data_ptr())
            if dev is None:
                dev = a.device
        else:
            conv.append(a)
    if dev is not None and dev.index not in _scratch:
        _ensure_scratch(dev)
    if dev is not None and dev.index != torch.cuda.current_device():
        # tensors on another GPU than the thread's current one: launch in THEIR context, on their stream
        with torch.cuda.device(dev):
            return call(name, *args)
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, name)(*conv, stream())
        e1.record()
        PROFILE.append((name, args, e0, e1))
    else:
        rc = getattr(L, name)(*conv, stream())
    if rc != 0:
        msg = L.livae_last_error().decode(errors="replace")
        raise RuntimeError(f"{name} failed (rc={rc}): {msg}")


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("livae: tensors must live on a CUDA (sm_100) device; there is no CPU path")
        if t.dtype != torch.float32:
            raise RuntimeError(f"livae: expected float32, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("livae: tensors must be contiguous")

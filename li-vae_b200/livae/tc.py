"""Tensor-core path of the rVAE: the encoder (STN + conv stack + heads) and the decoder as TWO
autograd Functions whose forward/backward are hand-scheduled sequences of C-ABI launches
(reference model.py:237-262, 305-326, 375-388).

Storage between layers is bf16 NHWC; parameters, their gradients, the 1-channel images, the
latent heads and all loss arithmetic stay fp32.  The dense layers (STN conv2, encoder c2-c4,
decoder d1-d3) and the three large Linear layers run on tcgen05 (csrc/conv_tc.cu,
csrc/wgrad_tc.cu); the 1-channel ends run on the SIMT kernels in csrc/thin.cu.  Convention inside
these Functions: a tensor handed from one layer's backward to the next is the PRE-activation
gradient (the ReLU mask of a tensor is applied by whichever kernel produces its gradient).
Parity, measured at the benchmark's shapes (tests/test_gpu_parity_c3.py): ELBO 1e-6 and reconstructions 2e-4
against the fp32 oracle (bars 1e-3 / 1e-2); decoder convolution gradients at 1e-2; gradients below the latent bottleneck at
0.25 - 0.95 of the bf16-operand floor of the network (which is above 1e-2 there for ANY 16-bit-operand evaluation,
the reference's own autocast modes included).
"""
from __future__ import annotations

import ctypes as C

import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import ops
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, call

BF = torch.bfloat16


def supported(patch_size: int, latent_dim: int, in_channels: int) -> bool:
    return in_channels == 1 and patch_size % 16 == 0 and patch_size >= 32 and 2 * latent_dim <= 256


def _pad_heads(n):
    """column count of the [fc_mu; fc_logvar] GEMM: the tensor-core kernels take 16, 32 or a multiple of 64 channels
    (conv_tc.cu channels_ok, wgrad_tc.cu) -- e.g. latent_dim 20 -> 40 columns -> 64"""
    return 16 if n <= 16 else 32 if n <= 32 else (n + 63) // 64 * 64


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def _heads_cat(muw, mub, lvw, lvb, Npad):
    """[fc_mu; fc_logvar] as ONE weight [2L, K] and one zero-padded bias [Npad], cached per optimiser epoch like the
    packed weights (ops._packed): the concatenation, the bias fill and both bf16 packs of the result used to be redone
    by every encoder pass (two per step)"""
    Ld = muw.shape[0]

    def make():
        wcat = torch.cat([muw.detach(), lvw.detach()], 0)
        bcat = torch.zeros(Npad, dtype=torch.float32, device=muw.device)
        bcat[:Ld] = mub.detach(); bcat[Ld:2 * Ld] = lvb.detach()
        return wcat, bcat

    others = tuple(v for t in (lvw, mub, lvb) for v in (id(t), t._version, t.data_ptr()))
    return ops._packed(muw, ("heads", Npad) + others, make)


def _linear_fwd(x2d, w, Cb, kh, kw, bias_pad, Npad, act, out_f32=True):
    """nn.Linear over an NHWC-flattened map as a 1x1 tensor-core convolution; x2d bf16 [B, kh*kw*Cb]"""
    B, K = x2d.shape
    N = w.shape[0]
    wp = ops.tc_pack_weights(w, N, Cb, kh, kw, 3, Cs_pad=Npad).view(1, Npad, K)
    return ops.tc_conv(x2d.view(B, 1, 1, K), wp, bias_pad, 1, 1, 1, 0, act, out_f32=out_f32).view(B, Npad)


def _linear_bwd(x2d, w, Cb, kh, kw, g_bf, Npad, relu_mask):
    """-> (gw [N, Cb*kh*kw] in torch's (c,h,w) order, gb [N], gx bf16 [B, K] (masked by relu_mask > 0))"""
    B, K = x2d.shape
    N = w.shape[0]
    gw_hwc, gb = ops.tc_conv_wgrad(x2d.view(B, 1, 1, K), g_bf.view(B, 1, 1, Npad), 1, 1, 1, 0)
    gw = _empty((N, K), torch.float32, w.device)
    call("livae_permute_linear_grad", gw_hwc, N, Cb, kh * kw, gw)
    wp = ops.tc_pack_weights(w, N, Cb, kh, kw, 4, Cs_pad=Npad).view(1, K, Npad)
    if Npad <= 64 and K % 8 == 0:
        # J = Npad <= 64: writing B x K and reading the mask is the whole cost -> coalesced skinny kernel (csrc/skinny.cu)
        gx = _empty((B, K), BF, w.device)
        call("livae_linear_dgrad", g_bf, wp, relu_mask, B, K, Npad, gx)
    else:
        gx = ops.tc_conv(g_bf.view(B, 1, 1, Npad), wp, None, 1, 1, 1, 0, ACT_NONE, out_f32=False,
                         relu_mask=relu_mask.view(B, 1, 1, K) if relu_mask is not None else None).view(B, K)
    return gw, gb[:N].contiguous(), gx


def _dgrad_s2(gy, w, Cout, Cin, hin, mask):
    """data gradient of an encoder 4x4 stride-2 convolution, masked by the ReLU of the layer below"""
    # block form (one launch, N = 4*Cin, 4 of 9 taps per phase non-zero) wins while N <= 128; at Cin = 64 the four
    # per-phase launches (N = 64, no zero taps) are 1.3x faster (tools/micro.py s2blk_64_128_32)
    if Cin <= 32 and ops.dgrad_s2blk_supported(hin, hin, Cin, Cout):
        return ops.dgrad_s2blk(gy, w, hin, hin, relu_mask=mask)
    return ops.tc_conv_dgrad(gy, ops.tc_pack_weights(w, Cout, Cin, 4, 4, 2), None, hin, hin, 4, 4, 2, 1, relu_mask=mask)


class EncoderTc(Function):
    """(x, 20 parameters) -> (mu, logvar, theta); reference Encoder.forward, model.py:305-326"""

    @staticmethod
    def forward(ctx, x, w0, b0, w3, b3, w7, b7, w9, b9, c0w, c0b, c2w, c2b, c4w, c4b, c6w, c6b, muw, mub, lvw, lvb,
                theta_only=False):
        """theta_only: stop after the STN localisation (returns None for mu, logvar and the rotated batch) -- for call
        sites that discard everything but theta (train.py:376-377, pretrain_stn.py:106-107)"""
        ops.require_cuda(x)
        x = x.contiguous()
        B, _, P, _ = x.shape
        dev = x.device
        Ld = muw.shape[0]
        h, q4, q16 = P // 2, P // 4, P // 16
        # --- STN localisation (model.py:203-214)
        a1 = _empty((B, h, h, 16), BF, dev); idx1 = _empty((B, h, h, 16), torch.uint8, dev)
        call("livae_thin_conv1c_fwd", 0, x, w0, b0, B, P, P, a1, idx1)
        if ops.conv5pool_supported(B, h, h, 16, 32):     # space-to-depth form, pooling in the epilogue
            a2, idx2 = ops.conv5pool_fwd(a1, w3, b3)
        else:
            a2f = ops.tc_conv(a1, ops.tc_pack_weights(w3, 32, 16, 5, 5, 0), b3, 5, 5, 1, 2, ACT_RELU)
            a2 = _empty((B, q4, q4, 32), BF, dev); idx2 = _empty((B, q4, q4, 32), torch.uint8, dev)
            call("livae_maxpool_bf16", a2f, B, h, h, 32, a2, idx2)
            del a2f
        f1 = _linear_fwd(a2.view(B, -1), w7, 32, q4, q4, b7, 32, ACT_RELU)           # fp32 [B,32]
        vec = _empty((B, 2), torch.float32, dev)
        cs = _empty((B, 2), torch.float32, dev); theta = _empty((B, 1), torch.float32, dev)
        call("livae_stn_tail_fwd", f1, w9, b9, B, 32, vec, cs, theta)         # fc2 + normalize + atan2
        if theta_only:
            e = _empty((0,), BF, dev)
            ctx.save_for_backward(x, w0, w3, w7, w9, c0w, c2w, c4w, c6w, w9, a1, idx1, a2, idx2, f1, vec, cs, e, e, e, e, e)
            ctx.dims = (B, P, Ld, 0)
            ctx.set_materialize_grads(False)
            return None, None, theta, None
        x_rot = torch.empty_like(x)
        call("livae_rot_sample_fwd", x, cs, 1.0, B, 1, P, P, x_rot)
        # --- encoder conv stack (model.py:289-298)
        h1 = _empty((B, h, h, 32), BF, dev)
        call("livae_thin_conv1c_fwd", 1, x_rot, c0w, c0b, B, P, P, h1, None)
        h2 = ops.tc_conv(h1, ops.tc_pack_weights(c2w, 64, 32, 4, 4, 0), c2b, 4, 4, 2, 1, ACT_RELU)
        h3 = ops.tc_conv(h2, ops.tc_pack_weights(c4w, 128, 64, 4, 4, 0), c4b, 4, 4, 2, 1, ACT_RELU)
        h4 = ops.tc_conv(h3, ops.tc_pack_weights(c6w, 256, 128, 4, 4, 0), c6b, 4, 4, 2, 1, ACT_RELU)
        # --- heads (model.py:302-303, 321-324): one GEMM for [fc_mu; fc_logvar]
        Npad = _pad_heads(2 * Ld)
        wcat, bcat = _heads_cat(muw, mub, lvw, lvb, Npad)
        mulv = _linear_fwd(h4.view(B, -1), wcat, 256, q16, q16, bcat, Npad, ACT_NONE)
        mu = mulv[:, :Ld].contiguous(); logvar = mulv[:, Ld:2 * Ld].contiguous()
        ctx.save_for_backward(x, w0, w3, w7, w9, c0w, c2w, c4w, c6w, wcat, a1, idx1, a2, idx2, f1, vec, cs, x_rot,
                              h1, h2, h3, h4)
        ctx.dims = (B, P, Ld, Npad)
        ctx.set_materialize_grads(False)
        # x_rot is returned as a fourth, differentiable output: it IS rotate_to_canonical(x, theta)
        # (train.py:675-677 resamples the same x with the same rotation), so the trainer can take it from
        # Encoder.take_canonical instead of launching a second rot_sample forward + backward
        return mu, logvar, theta, x_rot

    @staticmethod
    @once_differentiable
    def backward(ctx, g_mu, g_lv, g_theta, g_xrot_out=None):
        (x, w0, w3, w7, w9, c0w, c2w, c4w, c6w, wcat, a1, idx1, a2, idx2, f1, vec, cs, x_rot,
         h1, h2, h3, h4) = ctx.saved_tensors
        B, P, Ld, Npad = ctx.dims
        dev = x.device
        h, q4, q16 = P // 2, P // 4, P // 16
        if g_mu is None and g_lv is None:
            # only theta is consumed downstream (train.py:376-377, pretrain_stn.py:106-107): the
            # gradient reaches the STN localisation only
            gcs = None
            if g_xrot_out is not None:
                gcs = _empty((B, 2), torch.float32, dev)
                call("livae_rot_sample_bwd", x, cs, 1.0, g_xrot_out.contiguous(), B, 1, P, P, None, gcs)
            if g_theta is None and gcs is None:
                return (None,) * 22
            return EncoderTc._backward_stn(ctx, gcs, g_theta)
        # --- heads
        g16 = torch.zeros((B, Npad), dtype=torch.float32, device=dev)
        if g_mu is not None:
            g16[:, :Ld] = g_mu
        if g_lv is not None:
            g16[:, Ld:2 * Ld] = g_lv
        g16b = ops.cast(g16, BF)
        gwcat, gbcat, gh4 = _linear_bwd(h4.view(B, -1), wcat, 256, q16, q16, g16b, Npad, h4.view(B, -1))
        gh4 = gh4.view(B, q16, q16, 256)
        # --- encoder convs c4, c3, c2 (tensor cores)
        gw6, gb6 = ops.tc_conv_wgrad(h3, gh4, 4, 4, 2, 1)
        gh3 = _dgrad_s2(gh4, c6w, 256, 128, P // 8, h3)
        gw4, gb4 = ops.tc_conv_wgrad(h2, gh3, 4, 4, 2, 1)
        gh2 = _dgrad_s2(gh3, c4w, 128, 64, q4, h2)
        gw2, gb2 = ops.tc_conv_wgrad(h1, gh2, 4, 4, 2, 1)
        gh1 = _dgrad_s2(gh2, c2w, 64, 32, h, h1)
        # --- encoder c1 (thin) and the rotation
        gc0w = torch.empty_like(c0w); gc0b = _empty((32,), torch.float32, dev)
        call("livae_thin_conv1c_wgrad", 1, x_rot, gh1, None, B, P, P, gc0w, gc0b)
        g_xrot = torch.empty_like(x)
        call("livae_thin_conv1c_dgrad", gh1, c0w, B, P, P, g_xrot)
        if g_xrot_out is not None:      # the canonical-MSE gradient arrives on the same tensor (linear in g)
            call("livae_axpby_dev", g_xrot, None, g_xrot_out.contiguous(), None, g_xrot.numel(), g_xrot)
        gcs = _empty((B, 2), torch.float32, dev)
        call("livae_rot_sample_bwd", x, cs, 1.0, g_xrot, B, 1, P, P, None, gcs)
        stn = EncoderTc._backward_stn(ctx, gcs, g_theta)
        gmuw, glvw = gwcat[:Ld].contiguous(), gwcat[Ld:2 * Ld].contiguous()
        gmub, glvb = gbcat[:Ld].contiguous(), gbcat[Ld:2 * Ld].contiguous()
        return stn[:9] + (gc0w, gc0b, gw2, gb2, gw4, gb4, gw6, gb6, gmuw.view(Ld, -1), gmub, glvw.view(Ld, -1), glvb, None)

    @staticmethod
    def _backward_stn(ctx, gcs, g_theta):
        """gradients of the 8 STN parameters from d/d(cos,sin) and d/dtheta (model.py:203-214, 245-261)"""
        (x, w0, w3, w7, w9, c0w, c2w, c4w, c6w, wcat, a1, idx1, a2, idx2, f1, vec, cs, x_rot,
         h1, h2, h3, h4) = ctx.saved_tensors
        B, P, Ld, Npad = ctx.dims
        dev = x.device
        h, q4 = P // 2, P // 4
        # --- STN head + fc2 (one fused kernel), fc1, conv2 (tensor cores), conv1 (thin)
        gw9 = torch.empty_like(w9); gb9 = _empty((2,), torch.float32, dev)
        gf1b = _empty((B, 32), BF, dev)
        call("livae_stn_tail_bwd", f1, w9, vec, gcs, g_theta.contiguous() if g_theta is not None else None, B, 32,
             gw9, gb9, gf1b)
        gw7, gb7, ga2 = _linear_bwd(a2.view(B, -1), w7, 32, q4, q4, gf1b, 32, a2.view(B, -1))
        if ops.conv5pool_supported(B, h, h, 16, 32):
            gw3, gb3, ga1 = ops.conv5pool_bwd(a1, w3, ga2.contiguous(), idx2)
        else:
            g2full = _empty((B, h, h, 32), BF, dev)
            call("livae_unpool_bf16", ga2, idx2, B, h, h, 32, g2full)
            gw3, gb3 = ops.tc_conv_wgrad(a1, g2full, 5, 5, 1, 2)
            ga1 = ops.tc_conv_dgrad(g2full, ops.tc_pack_weights(w3, 32, 16, 5, 5, 2), None, h, h, 5, 5, 1, 2,
                                    relu_mask=a1)
        gw0 = torch.empty_like(w0); gb0 = _empty((16,), torch.float32, dev)
        call("livae_thin_conv1c_wgrad", 0, x, ga1, idx1, B, P, P, gw0, gb0)
        return (None, gw0, gb0, gw3, gb3, gw7.view_as(w7), gb7, gw9, gb9) + (None,) * 13


# Decoder block d3 (64 -> 32 channels, the widest map) runs phase-folded onto its low-resolution input (csrc/upfold.cu).
# LIVAE_UPFOLD=0: always the materialised Upsample -> ReflectionPad path; 2: fold every block the kernels accept (d2 as
# well -- measured slower there: its 16 x 16 input fills the 8 x 14 halo tiles to 57 %, see profiles/r02_notes.md).
UPFOLD = int(os.environ.get("LIVAE_UPFOLD", "1"))


class DecoderTc(Function):
    """(z, 10 parameters) -> recon [B,1,P,P]; reference Decoder.forward, model.py:375-388"""

    @staticmethod
    def forward(ctx, z, fcw, fcb, d1w, d1b, d2w, d2b, d3w, d3b, d4w, d4b):
        ops.require_cuda(z)
        z = z.contiguous()
        B, Ld = z.shape
        dev = z.device
        q = int(round((fcw.shape[0] // 256) ** 0.5))
        y0 = _empty((B, q, q, 256), BF, dev)
        call("livae_decfc_fwd_bf16", z, fcw, fcb, B, Ld, 256, q * q, y0)
        ys, aux = [y0], []
        cur, hw = y0, q
        folded = []
        for w, b, cin, cout in ((d1w, d1b, 256, 128), (d2w, d2b, 128, 64), (d3w, d3b, 64, 32)):
            if UPFOLD and (cout == 32 or UPFOLD > 1) and ops.upfold_supported(B, hw, hw, cin, cout):
                # phase-folded block (csrc/upfold.cu): no up-sampled tensor; the border strips are kept for the weight gradient
                cur, s_tb, s_lr = ops.upfold_fwd(cur, w, b)
                aux += [s_tb, s_lr]; folded.append(True)
            else:
                u = _empty((B, 2 * hw + 2, 2 * hw + 2, cin), BF, dev)
                call("livae_upsample_pad_fwd_bf16", cur, B, hw, hw, cin, u)
                cur = ops.tc_conv(u, ops.tc_pack_weights(w, cout, cin, 3, 3, 0), b, 3, 3, 1, 0, ACT_RELU)
                aux += [u, u]; folded.append(False)
            ys.append(cur)
            hw *= 2
        P = 2 * hw
        recon = _empty((B, 1, P, P), torch.float32, dev)
        # d4 (upsample -> pad -> 32->1 conv -> sigmoid) straight from the low-resolution map (csrc/upconv_c1.cu)
        call("livae_upconv_c1_fwd", cur, d4w, d4b, B, hw, hw, ACT_SIGMOID, recon)
        ctx.save_for_backward(z, fcw, d1w, d2w, d3w, d4w, recon, *ys, *aux)
        ctx.dims = (B, Ld, q, P)
        ctx.folded = tuple(folded)
        return recon

    @staticmethod
    @once_differentiable
    def backward(ctx, g_recon):
        saved = ctx.saved_tensors
        z, fcw, d1w, d2w, d3w, d4w, recon = saved[:7]
        ys, aux = saved[7:11], saved[11:17]
        B, Ld, q, P = ctx.dims
        folded = ctx.folded
        dev = z.device
        gpre4 = torch.empty_like(recon)
        call("livae_sigmoid_bwd", recon, g_recon.contiguous(), None, recon.numel(), gpre4)
        gd4w = torch.empty_like(d4w); gd4b = _empty((1,), torch.float32, dev)
        grads = []
        hw = P // 2
        gu = gx = None          # gradient handed down by the layer above: w.r.t. its up-sampled input / its low-res input
        for i, (w, cin, cout) in zip((3, 2, 1), ((d3w, 64, 32), (d2w, 128, 64), (d1w, 256, 128))):
            y = ys[i]
            gb = _empty((cout,), torch.float32, dev)
            if i == 3:
                # the whole backward of d4 (weight, bias, data gradient + upsample/pad adjoint + this layer's bias
                # gradient) in one kernel over the LOW-resolution tensors (csrc/upconv_c1.cu)
                gy = torch.empty_like(y)                          # pre-activation gradient of conv i
                call("livae_upconv_c1_bwd", gpre4, d4w, y, B, hw, hw, gy, gb, gd4w, gd4b)
            elif gx is not None:                                  # the folded block above already applied this layer's ReLU mask
                gy = gx
                call("livae_colsum_bf16", gy, B * hw * hw, cout, gb)
            else:
                gy = torch.empty_like(y)
                call("livae_upsample_pad_bwd_bias_bf16", gu, B, hw, hw, cout, y, gy, gb)
            if folded[i - 1]:
                gw, gx = ops.upfold_bwd(ys[i - 1], w, gy, aux[2 * (i - 1)], aux[2 * (i - 1) + 1])
                gu = None
            else:
                u = aux[2 * (i - 1)]
                gw, _ = ops.tc_conv_wgrad(u, gy, 3, 3, 1, 0, want_bias=False)
                gu = ops.tc_conv_dgrad(gy, ops.tc_pack_weights(w, cout, cin, 3, 3, 2), None, hw + 2, hw + 2, 3, 3, 1, 0)
                gx = None
            grads.append((gw, gb))
            hw //= 2
        if gx is not None:
            gy0 = gx
        else:
            gy0 = torch.empty_like(ys[0])
            call("livae_upsample_pad_bwd_bf16", gu, B, q, q, 256, ys[0], gy0)

        gfcw = torch.empty_like(fcw); gfcb = _empty((fcw.shape[0],), torch.float32, dev)
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        call("livae_decfc_bwd_bf16", z, fcw, gy0, B, Ld, 256, q * q, gfcw, gfcb, gz)
        (g3w, g3b), (g2w, g2b), (g1w, g1b) = grads
        return gz, gfcw, gfcb, g1w, g1b, g2w, g2b, g3w, g3b, gd4w, gd4b


# ------------------------------------------------------------------------------------------------
# plain VAE (reference model.py:9-182; scripts/train_vae.py): the same conv stack without the STN, and a
# ConvTranspose2d decoder.  ConvTranspose2d(k4,s2,p1) forward IS the stride-2 data-gradient kernel (its weight
# [Cin,Cout,kh,kw] is the weight [Cout',Cin',kh,kw] of the convolution it is the adjoint of); its data gradient is
# that convolution's forward and its weight gradient that convolution's weight gradient with input and output
# gradient swapped.  The 32 -> 1 layer (N = 1) runs on the fp32 engine.
# ------------------------------------------------------------------------------------------------
class VAEEncoderTc(Function):
    """(x, 12 parameters) -> (mu, logvar); reference VAEEncoder.forward, model.py:45-61"""

    @staticmethod
    def forward(ctx, x, c0w, c0b, c2w, c2b, c4w, c4b, c6w, c6b, muw, mub, lvw, lvb):
        ops.require_cuda(x)
        x = x.contiguous()
        B, _, P, _ = x.shape
        dev = x.device
        Ld = muw.shape[0]
        h, q16 = P // 2, P // 16
        h1 = _empty((B, h, h, 32), BF, dev)
        call("livae_thin_conv1c_fwd", 1, x, c0w, c0b, B, P, P, h1, None)
        h2 = ops.tc_conv(h1, ops.tc_pack_weights(c2w, 64, 32, 4, 4, 0), c2b, 4, 4, 2, 1, ACT_RELU)
        h3 = ops.tc_conv(h2, ops.tc_pack_weights(c4w, 128, 64, 4, 4, 0), c4b, 4, 4, 2, 1, ACT_RELU)
        h4 = ops.tc_conv(h3, ops.tc_pack_weights(c6w, 256, 128, 4, 4, 0), c6b, 4, 4, 2, 1, ACT_RELU)
        Npad = _pad_heads(2 * Ld)
        wcat, bcat = _heads_cat(muw, mub, lvw, lvb, Npad)
        mulv = _linear_fwd(h4.view(B, -1), wcat, 256, q16, q16, bcat, Npad, ACT_NONE)
        ctx.save_for_backward(x, c2w, c4w, c6w, wcat, h1, h2, h3, h4)
        ctx.dims = (B, P, Ld, Npad)
        ctx.set_materialize_grads(False)
        return mulv[:, :Ld].contiguous(), mulv[:, Ld:2 * Ld].contiguous()

    @staticmethod
    @once_differentiable
    def backward(ctx, g_mu, g_lv):
        x, c2w, c4w, c6w, wcat, h1, h2, h3, h4 = ctx.saved_tensors
        B, P, Ld, Npad = ctx.dims
        dev = x.device
        h, q4, q16 = P // 2, P // 4, P // 16
        if g_mu is None and g_lv is None:
            return (None,) * 13
        g16 = torch.zeros((B, Npad), dtype=torch.float32, device=dev)
        if g_mu is not None:
            g16[:, :Ld] = g_mu
        if g_lv is not None:
            g16[:, Ld:2 * Ld] = g_lv
        gwcat, gbcat, gh4 = _linear_bwd(h4.view(B, -1), wcat, 256, q16, q16, ops.cast(g16, BF), Npad, h4.view(B, -1))
        gh4 = gh4.view(B, q16, q16, 256)
        gw6, gb6 = ops.tc_conv_wgrad(h3, gh4, 4, 4, 2, 1)
        gh3 = _dgrad_s2(gh4, c6w, 256, 128, P // 8, h3)
        gw4, gb4 = ops.tc_conv_wgrad(h2, gh3, 4, 4, 2, 1)
        gh2 = _dgrad_s2(gh3, c4w, 128, 64, q4, h2)
        gw2, gb2 = ops.tc_conv_wgrad(h1, gh2, 4, 4, 2, 1)
        gh1 = _dgrad_s2(gh2, c2w, 64, 32, h, h1)
        gc0w = _empty((32, 1, 4, 4), torch.float32, dev); gc0b = _empty((32,), torch.float32, dev)
        call("livae_thin_conv1c_wgrad", 1, x, gh1, None, B, P, P, gc0w, gc0b)
        gmuw, glvw = gwcat[:Ld].contiguous(), gwcat[Ld:2 * Ld].contiguous()
        gmub, glvb = gbcat[:Ld].contiguous(), gbcat[Ld:2 * Ld].contiguous()
        return (None, gc0w, gc0b, gw2, gb2, gw4, gb4, gw6, gb6, gmuw.view(Ld, -1), gmub, glvw.view(Ld, -1), glvb)


class VAEDecoderTc(Function):
    """(z, 10 parameters) -> recon [B,1,P,P]; reference VAEDecoder.forward, model.py:100-113"""

    @staticmethod
    def forward(ctx, z, fcw, fcb, t1w, t1b, t2w, t2b, t3w, t3b, t4w, t4b):
        ops.require_cuda(z)
        z = z.contiguous()
        B, Ld = z.shape
        dev = z.device
        q = int(round((fcw.shape[0] // 256) ** 0.5))
        y0 = _empty((B, q, q, 256), BF, dev)
        call("livae_decfc_fwd_bf16", z, fcw, fcb, B, Ld, 256, q * q, y0)
        ys = [y0]
        cur, hw = y0, q
        for w, b, cin, cout in ((t1w, t1b, 256, 128), (t2w, t2b, 128, 64), (t3w, t3b, 64, 32)):
            # ConvTranspose2d weight [cin, cout, 4, 4] == weight of the adjoint convolution cout -> cin
            cur = ops.tc_conv_dgrad(cur, ops.tc_pack_weights(w, cin, cout, 4, 4, 2), b, 2 * hw, 2 * hw, 4, 4, 2, 1, ACT_RELU)
            ys.append(cur)
            hw *= 2
        P = 2 * hw
        recon = _empty((B, 1, P, P), torch.float32, dev)
        call("livae_thin_convt_c1_fwd", cur, t4w, t4b, B, hw, hw, 32, ACT_SIGMOID, recon)     # N = 1: thin kernel
        ctx.save_for_backward(z, fcw, t1w, t2w, t3w, t4w, recon, *ys)
        ctx.dims = (B, Ld, q, P)
        return recon

    @staticmethod
    @once_differentiable
    def backward(ctx, g_recon):
        saved = ctx.saved_tensors
        z, fcw, t1w, t2w, t3w, t4w, recon = saved[:7]
        ys = saved[7:11]
        B, Ld, q, P = ctx.dims
        dev = z.device
        hw = P // 2
        gpre4 = torch.empty_like(recon)
        call("livae_sigmoid_bwd", recon, g_recon.contiguous(), None, recon.numel(), gpre4)
        gt4w = torch.empty_like(t4w); gt4b = _empty((1,), torch.float32, dev)
        call("livae_thin_convt_c1_wgrad", ys[3], gpre4, B, hw, hw, 32, gt4w, gt4b)
        g = _empty((B, hw, hw, 32), BF, dev)                     # pre-activation gradient of t3's output
        call("livae_thin_convt_c1_dgrad", gpre4, t4w, ys[3], B, hw, hw, 32, g)
        grads = []
        for i, (w, cin, cout) in zip((3, 2, 1), ((t3w, 64, 32), (t2w, 128, 64), (t1w, 256, 128))):
            y_in = ys[i - 1]                                      # [B, hw/2, hw/2, cin], post-ReLU input of layer i
            gb = _empty((cout,), torch.float32, dev)
            call("livae_colsum_bf16", g, B * hw * hw, cout, gb)
            gw, _ = ops.tc_conv_wgrad(g, y_in, 4, 4, 2, 1, want_bias=False)          # [cin, cout, 4, 4]
            g = ops.tc_conv(g, ops.tc_pack_weights(w, cin, cout, 4, 4, 0), None, 4, 4, 2, 1, ACT_NONE, relu_mask=y_in)
            grads.append((gw, gb))
            hw //= 2
        gfcw = torch.empty_like(fcw); gfcb = _empty((fcw.shape[0],), torch.float32, dev)
        gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        call("livae_decfc_bwd_bf16", z, fcw, g, B, Ld, 256, q * q, gfcw, gfcb, gz)
        (g3w, g3b), (g2w, g2b), (g1w, g1b) = grads
        return gz, gfcw, gfcb, g1w, g1b, g2w, g2b, g3w, g3b, gt4w, gt4b

/*
 * livae_b200.h -- C ABI of liblivae_sm100.so, the B200 (sm_100a) kernels under the
 * LI-VAE rVAE/VAE training-step hot path.
 *
 * The reference (jerrydzhang/LI-VAE) is pure Python/PyTorch and exposes NO FFI or
 * operator/plugin layer; the drop-in boundary towards users is the
 * Python API in li-vae_b200/livae/.  This header is the thin C-ABI underneath those
 * classes: plain device pointers and sizes, no torch types.  Each entry point cites
 * the reference code (path under /root/reference, file:line) whose arithmetic it
 * replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all activation tensors are NHWC ("pixels x channels"), contiguous;
 *     1-channel images are therefore identical to the reference's NCHW layout;
 *   - `dt` selects the storage type of activations/packed weights:
 *       LIVAE_F32 (0) float, LIVAE_F16 (1) __half, LIVAE_BF16 (2) __nv_bfloat16;
 *     accumulation is always fp32, parameters and their gradients are always fp32
 *     in torch's own layouts (Conv2d [Cout,Cin,kh,kw], ConvTranspose2d
 *     [Cin,Cout,kh,kw], Linear [out,in] with NCHW flatten order);
 *   - return 0 on success, <0 argument error, >0 cudaError_t; message via
 *     livae_last_error() (thread local).  Nothing is allocated, retained or freed
 *     by the library; workspaces are passed in.  All work is enqueued on `stream`.
 */
#ifndef LIVAE_B200_H
#define LIVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* livae_stream_t; /* cudaStream_t */

enum { LIVAE_F32 = 0, LIVAE_F16 = 1, LIVAE_BF16 = 2 };
enum { LIVAE_ACT_NONE = 0, LIVAE_ACT_RELU = 1, LIVAE_ACT_SIGMOID = 2 };

/* conv "kinds": which torch weight layout is read and which side is the strided one */
enum {
  LIVAE_CONV = 0,        /* nn.Conv2d, zero padding                      (model.py:204,207,290-296) */
  LIVAE_CONVT = 2        /* nn.ConvTranspose2d(k4,s2,p1)                 (model.py:90-96) */
};

int livae_abi_version(void);
const char* livae_last_error(void);
/* kernels launched by this library since load (bench.py reports the delta as gpu_launches) */
int64_t livae_launch_count(void);
/* 1 if the running device is sm_100 (B200); kernels refuse to launch otherwise */
int livae_device_ok(void);
/* Register (ptr != NULL) or withdraw (ptr == NULL) a caller-owned device scratch buffer for the CURRENT device,
 * 256-byte aligned.  The library never allocates: kernels that split a reduction over CTAs (split-K Linear layers,
 * tensor-core weight gradients, column sums) park per-CTA partial sums there and add them in a fixed order, which
 * makes the forward pass and those gradients bit-reproducible (the reference's CPU path is).  Without a
 * buffer (or with too small a one: 96 MB covers every layer of the rVAE at batch 2048) they fall back to
 * red.global.add.  The pointer is retained until replaced; use it from one stream at a time. */
int livae_set_scratch(void* ptr, int64_t bytes);

/* ---- a1: peak-centred integer patch gather ------------------------------------------
 * replaces PatchDataset.__getitem__ with transform=None (data.py:211-250): whole-image
 * translate + center_crop == float32(img)[cy-P/2:cy+P/2, cx-P/2:cx+P/2], bit exact.
 * images: [n_img,H,W] (float or double, cast to float like data.py:226);
 * sites: int32 [N,3] = (img_idx, cy, cx); out: float [N,P,P].  Out-of-image pixels -> 0. */
int livae_patch_gather_f32(const float* images, int n_img, int H, int W,
                           const int32_t* sites, int N, int P, float* out, livae_stream_t stream);
int livae_patch_gather_f64(const double* images, int n_img, int H, int W,
                           const int32_t* sites, int N, int P, float* out, livae_stream_t stream);
/* a2 tail: per-patch min-max normalisation to [0,1] (data.py:553-558), in place */
/* a2: sub-pixel patch gather == AdaptiveLatticeDataset.__getitem__ with transform=None before its min-max
   (data.py:478-551): bilinear resampling of float32(img), zero outside the image, centred on the float site.
   img_idx: int32 [N]; yx: float64 [N,2] (cy, cx); P even; out: fp32 [N,1,P,P].  Follow with livae_patch_minmax
   (data.py:553-558).  Within 2e-5 of the reference (whose torchvision affine grid is fp32). */
int livae_patch_gather_subpixel_f32(const float* images, int n_img, int H, int W, const int32_t* img_idx,
                                    const double* yx, int N, int P, float* out, livae_stream_t stream);
int livae_patch_gather_subpixel_f64(const double* images, int n_img, int H, int W, const int32_t* img_idx,
                                    const double* yx, int N, int P, float* out, livae_stream_t stream);
int livae_patch_minmax(float* patches, int N, int P, livae_stream_t stream);
/* a2 with the ROI window: the [S,S] `patch_big` (S = P + 2*padding) of Adaptive/PairedAdaptiveLatticeDataset
   .__getitem__ (data.py:496-546, 640-691).  As livae_patch_gather_subpixel, but only pixels inside the reference's
   integer window of `roi` = P + max(16, 2*padding) pixels around round-half-even(site) are read (zero beyond
   it, as after TF.pad + TF.affine(fill=None)).  roi even, >= S. */
int livae_patch_gather_roi_f32(const float* images, int n_img, int H, int W, const int32_t* img_idx,
                               const double* yx, int N, int S, int roi, float* out, livae_stream_t stream);
int livae_patch_gather_roi_f64(const double* images, int n_img, int H, int W, const int32_t* img_idx,
                               const double* yx, int N, int S, int roi, float* out, livae_stream_t stream);
/* a3: default_transform(rotation=False) on [N,S,S] patches (data.py:78-116): scale by scale[n] about the centre
   (TF.affine, bilinear, zeros outside; skipped when flags[n] bit2 is set), flags[n] bit0 = hflip, bit1 = vflip, then torch.roll by
   shift[n] = (shift_y, shift_x).  The draws are the caller's (Python `random` in the reference). out != in. */
int livae_augment(const float* in, int N, int S, const float* scale, const int32_t* flags, const int32_t* shift,
                  float* out, livae_stream_t stream);
/* TF.rotate(angle_deg[n], bilinear, expand=False, fill=0) (mode 1; data.py:97-103, 698-704) or identity
   (mode 0, angle_deg may be NULL) of [N,S,S] patches, TF.center_crop to [N,P,P] (data.py:710-713) and, if
   normalise, the per-patch min-max of data.py:716-730 (constant patch -> zeros).  S - P even. out != in. */
int livae_rotate_crop(const float* in, int N, int S, int P, const double* angle_deg, int mode, int normalise,
                      float* out, livae_stream_t stream);

/* ---- a6: fused rotate + bilinear sample (affine_grid + grid_sample) ------------------
 * replaces F.affine_grid + F.grid_sample(bilinear, reflection, align_corners=False) at
 * model.py:254-258, model.py:467-470 and train.py:675-677.  cs: [B,2] = (cos, sin);
 * the matrix used is [[c,-sgn*s,0],[sgn*s,c,0]] so sgn=-1 gives the inverse rotation.
 * img/out: float [B,C,H,W].  No [B,H,W,2] grid is materialised. */
int livae_rot_sample_fwd(const float* img, const float* cs, float sgn, int B, int C, int H, int W,
                         float* out, livae_stream_t stream);
/* gimg may be NULL (x never requires grad).  gcs: [B,2] = dL/d(cos), dL/d(sin) (written,
 * not accumulated; already multiplied by sgn for the sin component). */
int livae_rot_sample_bwd(const float* img, const float* cs, float sgn, const float* gout,
                         int B, int C, int H, int W, float* gimg, float* gcs, livae_stream_t stream);

/* ---- a4 tail: rotation head ------------------------------------------------------------
 * F.normalize(vec, eps=1e-6) -> (cos, sin); theta = atan2(sin, cos)   (model.py:245-261) */
int livae_stn_head_fwd(const float* vec, int B, float* cs, float* theta, livae_stream_t stream);
/* gvec = J^T (gcs + gtheta * (-s, c));  gcs / gtheta may be NULL */
int livae_stn_head_bwd(const float* vec, const float* gcs, const float* gtheta, int B,
                       float* gvec, livae_stream_t stream);
/* STN tail fused: Linear(32 -> 2) (model.py:213) + livae_stn_head_{fwd,bwd}.  f1: fp32 [B,K] post-ReLU fc1 output,
   w9 [2,K], b9 [2], K == 32.  fwd writes vec [B,2], cs [B,2], theta [B] (may be NULL).  bwd: gcs / gtheta may be
   NULL; writes gw9 [2,K], gb9 [2] and gf1 = (gvec w9) * (f1 > 0) as bf16 [B,K]. */
int livae_stn_tail_fwd(const float* f1, const float* w9, const float* b9, int B, int K, float* vec, float* cs,
                       float* theta, livae_stream_t stream);
int livae_stn_tail_bwd(const float* f1, const float* w9, const float* vec, const float* gcs, const float* gtheta,
                       int B, int K, float* gw9, float* gb9, void* gf1_bf16, livae_stream_t stream);
/* (cos, sin) of an angle tensor, for get_rotation_matrix(theta) (model.py:220-235) */
int livae_angle_to_cs(const float* theta, int B, float* cs, livae_stream_t stream);
/* gtheta = -sin*gc + cos*gs */
int livae_angle_to_cs_bwd(const float* theta, const float* gcs, int B, float* gtheta, livae_stream_t stream);

/* ---- a8: reparameterisation (model.py:426-440; eps drawn by torch, passed in) ---------- */
int livae_reparam_fwd(const float* mu, const float* logvar, const float* eps, int n, float* z,
                      livae_stream_t stream);
/* gmu = gz, glv = gz*eps*0.5*exp(0.5*lv) */
int livae_reparam_bwd(const float* gz, const float* logvar, const float* eps, int n, float* gmu,
                      float* glv, livae_stream_t stream);

/* ---- a12/a13/a14: fused ELBO reductions ---------------------------------------------------
 * RVAELoss (loss.py:162-169): recon = sum((r-x)^2)/B, kld = mean_b(-0.5*sum_l(1+lv-mu^2-e^lv))
 * VAELoss  (loss.py:116-119): recon = mean((r-x)^2),  kld = -0.5*mean(1+lv-mu^2-e^lv)
 * canonical MSE (train.py:391-393): mean((recon - canonical_input)^2)   (n_lat = 0)
 * One launch: sums[0] = sum((r-x)^2), sums[1] = sum(-0.5*(1+lv-mu^2-e^lv)) (raw; the host side
 * applies 1/B, 1/n, beta to the two device scalars).  Deterministic two-stage reduction.
 * scratch: livae_elbo_scratch_floats() floats, zeroed once by the caller (self-resetting). */
int64_t livae_elbo_scratch_floats(void);
int livae_elbo_fwd(const float* recon, const float* x, int64_t n_pix, const float* mu,
                   const float* logvar, int n_lat, float* sums, float* scratch, livae_stream_t stream);
/* g: device pointer to the two upstream gradients (dL/dsums[0], dL/dsums[1]).
 * d_recon = g[0]*2(r-x); d_x = -d_recon; d_mu = g[1]*mu; d_lv = g[1]*0.5(e^lv-1).
 * d_recon / d_x may be NULL; d_mu, d_lv required when n_lat > 0. */
int livae_elbo_bwd(const float* recon, const float* x, int64_t n_pix, const float* mu,
                   const float* logvar, int n_lat, const float* g, float* d_recon, float* d_x,
                   float* d_mu, float* d_lv, livae_stream_t stream);
/* cycle_consistency_loss (loss.py:52-94): loss[0] = mean(1-cos(th_r - th + angle)) */
int livae_cycle_fwd(const float* theta, const float* theta_rot, const float* angle, int B,
                    float* loss, livae_stream_t stream);
/* d_theta = -g[0]*sin(d)/B, d_theta_rot = +g[0]*sin(d)/B (either may be NULL) */
int livae_cycle_bwd(const float* theta, const float* theta_rot, const float* angle, const float* g,
                    int B, float* d_theta, float* d_theta_rot, livae_stream_t stream);
/* out = a*x + b*y elementwise (y may be NULL); a_dev/b_dev are device pointers to one float
 * each (NULL = 1), so scaling by a device scalar needs no host sync. */
int livae_axpby_dev(const float* x, const float* a_dev, const float* y, const float* b_dev,
                    int64_t n, float* out, livae_stream_t stream);

/* ---- a4/a7/a9/a11: convolution / linear layers, engine 0 (exact fp32) -----------------------
 * One descriptor covers nn.Conv2d (+ReLU/Sigmoid, +MaxPool2d(2,2)) and
 * nn.ConvTranspose2d(k4,s2,p1) (model.py:204-209, 290-296, 359-371, 90-96).  A Linear over a
 * flattened NHWC feature map is the LIVAE_CONV whose kernel covers the whole map (kh=Hin,
 * kw=Win, pad 0): torch's Linear weight [N, C*H*W] in NCHW flatten order IS that conv's
 * [N,C,H,W] weight (model.py:210-213, 321-324).
 * x: [B,Hin,Win,Cin] NHWC fp32; y: [B,Ho,Wo,Cout] (Ho,Wo AFTER pooling when pool=1);
 * w, bias: fp32 in torch's own layouts (Conv2d [Cout,Cin,kh,kw]; ConvTranspose2d [Cin,Cout,kh,kw]). */
typedef struct {
  int kind;                 /* LIVAE_CONV / LIVAE_CONVT */
  int B, Hin, Win, Cin;
  int Cout, kh, kw, stride, pad;
  int act;                  /* LIVAE_ACT_* applied after bias */
  int pool;                 /* 1: MaxPool2d(2,2) after the activation (LIVAE_CONV only) */
} livae_conv_desc;

void livae_conv_out_shape(const livae_conv_desc* d, int* Ho, int* Wo);
int64_t livae_conv_fwd_ws_bytes(const livae_conv_desc* d);   /* non-zero only when pool=1 */
/* y = [pool](act(conv(x, w) + bias)); pool_idx: uint8 argmax position, same shape as y */
int livae_conv_fwd(const livae_conv_desc* d, const float* x, const float* w, const float* bias,
                   float* y, uint8_t* pool_idx, void* ws, livae_stream_t stream);
/* gy: gradient w.r.t. y (post-activation, post-pool).  The activation derivative and the pool
 * routing are applied while loading gy (y, pool_idx are the forward outputs).  gw, gb, gx are
 * written, not accumulated; each may be NULL. */
int livae_conv_bwd(const livae_conv_desc* d, const float* x, const float* w, const float* y,
                   const float* gy, const uint8_t* pool_idx, float* gw, float* gb, float* gx,
                   livae_stream_t stream);

/* ---- a4/a7/a9: convolution layers, engine 1 (tcgen05 + TMA, bf16 operands, fp32 accumulate) ---
 * The dense mid layers (encoder c2-c4, decoder d1-d3, STN conv2; model.py:207, 292-296, 359-367)
 * as a tap-decomposed implicit GEMM: per 128-pixel output tile and per filter tap one TMA box
 * load of the shifted NHWC input (zero padding = TMA OOB fill) feeds tcgen05.mma; accumulators
 * live in TMEM.  Parity class 1e-2 (bf16 GEMM inputs).
 * x: bf16 [B,Hin,Win,Cin]; wpacked: bf16 [kh*kw][Cout][Cin] from livae_tc_pack_weights;
 * y: bf16 (or fp32 when out_f32) [B,Ho,Wo,Cout]; relu_mask (optional): bf16, same shape as y,
 * y *= (relu_mask > 0) -- used when y is the gradient of a post-ReLU tensor. */
typedef struct {
  int B, Hin, Win, Cin;
  int Cout, kh, kw, stride, pad;
  int act;
  int out_f32;
} livae_tc_conv_desc;
int livae_tc_conv_supported(const livae_tc_conv_desc* d);
/* w: fp32 torch layout [Cs][Cb][kh][kw].  mode 0 -> [tap][Cs][Cb] (forward of Conv2d with
 * Cout=Cs, Cin=Cb); mode 1 -> [flipped tap][Cb][Cs] (its stride-1 data gradient run as a
 * forward convolution over gy with pad' = k-1-pad); mode 2 -> [tap][Cb][Cs] (livae_tc_conv_dgrad). */
/* modes 3 / 4: nn.Linear over an NHWC-flattened [kh,kw,Cb] map (model.py:210-213, 321-324), seen as a
 * 1x1 convolution with K = kh*kw*Cb in (h,w,c) order: mode 3 -> [Cs][tap][Cb] (forward), mode 4 ->
 * [tap][Cb][Cs] (data gradient).  Cs may be padded: rows >= Cs_real are written as zeros. */
int livae_tc_pack_weights(const float* w, int Cs, int Cb, int kh, int kw, int mode, int Cs_real,
                          void* out_bf16, livae_stream_t stream);
int livae_tc_conv(const livae_tc_conv_desc* d, const void* x, const void* wpacked, const float* bias,
                  void* y, const void* relu_mask, livae_stream_t stream);
/* Data gradient of the convolution d (= nn.ConvTranspose2d forward with x := gy, model.py:90-96):
 * gx[B,Hin,Win,Cin] = act(sum gy[b,(iy+pad-ky)/s,(ix+pad-kx)/s,co] * w[co,ci,ky,kx] + bias) [* (relu_mask>0)].
 * gy: bf16 [B,Ho,Wo,Cout]; wpacked: mode-2 packing; stride 2 runs as four output-parity phases. */
int livae_tc_conv_dgrad(const livae_tc_conv_desc* d, const void* gy, const void* wpacked,
                        const float* bias, void* gx, const void* relu_mask, livae_stream_t stream);
/* Weight (+ bias) gradient of the convolution d on tensor cores.  x: bf16 [B,Hin,Win,Cin];
 * gy: bf16 [B,Ho,Wo,Cout], already the PRE-activation gradient; gw: fp32 torch layout
 * [Cout][Cin][kh][kw] (written); gb: fp32 [Cout] (written, may be NULL);
 * ws: livae_tc_wgrad_ws_bytes(d) bytes of scratch. */
int64_t livae_tc_wgrad_ws_bytes(const livae_tc_conv_desc* d);
int livae_tc_conv_wgrad(const livae_tc_conv_desc* d, const void* x, const void* gy, float* gw, float* gb,
                        void* ws, livae_stream_t stream);
/* ---- conv 5x5 p2 + ReLU + MaxPool2 (STN conv2, model.py:207-209) in space-to-depth form (csrc/conv_s2d.cu):
 * a 3x3 convolution over 2x2 pixel blocks with 4*Ci -> 4*Co channels; the pooling is the epilogue.
 *   pack   mode 0: bf16 [9][4*Co][4*Ci] (forward), mode 1: bf16 [9][4*Ci][4*Co] (data gradient), w fp32 [Co][Ci][5][5]
 *   fwd    x bf16 [B,H,W,Ci] -> y bf16 [B,H/2,W/2,Co] + idx (argmax position 0..3, torch scan order)
 *   unpool_s2d: pooled PRE-activation gradient + idx -> g_s2d bf16 [B,H/2,W/2,4*Co]
 *   dgrad  g_s2d -> gx bf16 [B,H,W,Ci], masked by relu_mask (same layout as gx) > 0
 *   wgrad  gw fp32 [Co][Ci][5][5], gb fp32 [Co] (may be NULL); ws: livae_tc_conv5pool_wgrad_ws_bytes */
int livae_tc_conv5pool_supported(int B, int H, int W, int Ci, int Co);
int livae_tc_conv5pool_pack(const float* w, int Co, int Ci, int mode, void* out_bf16, livae_stream_t stream);
int livae_tc_conv5pool_fwd(const void* x, const void* wpacked0, const float* bias, int B, int H, int W, int Ci, int Co,
                           void* y, uint8_t* idx, livae_stream_t stream);
int livae_unpool_s2d_bf16(const void* g_pooled, const uint8_t* idx, int B, int Hp, int Wp, int Co, void* g_s2d,
                          livae_stream_t stream);
int livae_tc_conv5pool_dgrad(const void* g_s2d, const void* wpacked1, const void* relu_mask, int B, int H, int W, int Ci,
                             int Co, void* gx, livae_stream_t stream);
int64_t livae_tc_conv5pool_wgrad_ws_bytes(int Ci, int Co);
int livae_tc_conv5pool_wgrad(const void* x, const void* g_s2d, int B, int H, int W, int Ci, int Co, float* gw, float* gb,
                             void* ws, livae_stream_t stream);

/* Decoder blocks d1-d3, Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv2d(Cin->Cout, 3x3) -> ReLU (reference
 * model.py:356-368), phase-folded onto the LOW-resolution input (csrc/upfold.cu): one convolution Cin -> 4*Cout over
 * h x w pixels; the up-sampled tensor and its gradient are never formed.  x bf16 [B,h,w,Cin], y / gz bf16 [B,2h,2w,Cout].
 *   pack    w fp32 [Cout][Cin][3][3] -> wf bf16 [9][4*Cout][Cin] (forward), wd bf16 [9][Cin][4*Cout] (data gradient)
 *   strips  the border terms of the layer input: s_tb bf16 [2B,4,2w+2,Cin], s_lr bf16 [2B,4,2h+2,Cin] (transposed), two live
 *           and two zero rows each, so that a whole batch is ONE tall image [1,8B,.,Cin] for livae_tc_conv; its pad-0
 *           3x3 convolution with w (s_tb) / w transposed in (ky,kx) (s_lr) into fp32 is corr_tb [2B,4,2w,Cout] /
 *           corr_lr [2B,4,2h,Cout] (rows 0,1 of every group of 4 are used)
 *   fwd     y = ReLU(folded conv(x) + bias); the two outermost rows / columns are left as conv + bias
 *   ring    those rows / columns: y = ReLU(y + corr)
 *   gather  g_tb bf16 [2B,4,2w,Cout] / g_lr bf16 [2B,4,2h,Cout] = the two outermost rows / columns of gz + two zero rows
 *   dgrad   gx bf16 [B,h,w,Cin] = (folded data gradient of gz) * (x_mask > 0), x_mask = x (the ReLU output below)
 *   patch   gx += (x_mask > 0) * adjoint(strips)(gs_tb, gs_lr); gs_* fp32 = livae_tc_conv_dgrad of g_tb / g_lr
 *   wgrad   gw fp32 [Cout][Cin][3][3] = unfolded folded weight gradient + gw_tb + transpose(gw_lr) (either may be NULL) */
int livae_upfold_supported(int B, int h, int w, int Cin, int Cout);
int livae_upfold_pack(const float* w, int Cout, int Cin, void* wf, void* wd, livae_stream_t stream);
int livae_upfold_strips(const void* x, int B, int h, int w, int Cin, void* s_tb, void* s_lr, livae_stream_t stream);
int livae_upfold_fwd(const void* x, const void* wf, const float* bias, int B, int h, int w, int Cin, int Cout, void* y,
                     livae_stream_t stream);
int livae_upfold_ring(const float* corr_tb, const float* corr_lr, int B, int h, int w, int Cout, void* y, livae_stream_t stream);
int livae_upfold_gather(const void* gz, int B, int h, int w, int Cout, void* g_tb, void* g_lr, livae_stream_t stream);
int livae_upfold_dgrad(const void* gz, const void* wd, const void* x_mask, int B, int h, int w, int Cin, int Cout, void* gx,
                       livae_stream_t stream);
int livae_upfold_patch(const float* gs_tb, const float* gs_lr, const void* x_mask, int B, int h, int w, int Cin, void* gx,
                       livae_stream_t stream);
int64_t livae_upfold_wgrad_ws_bytes(int Cin, int Cout);
int livae_upfold_wgrad(const void* x, const void* gz, const float* gw_tb, const float* gw_lr, int B, int h, int w, int Cin,
                       int Cout, float* gw, void* ws, livae_stream_t stream);
/* data gradient of a 4x4 stride-2 pad-1 convolution as one 3x3 convolution over the gy grid that writes whole
 * 2x2 blocks of gx (csrc/conv_s2d.cu); wblk = livae_tc_dgrad_s2blk_pack(w fp32 [Cout][Cin][4][4]) bf16 [9][4*Cin][Cout] */
int livae_tc_dgrad_s2blk_supported(int Hin, int Win, int Cin, int Cout);
int livae_tc_dgrad_s2blk_pack(const float* w, int Cout, int Cin, void* out_bf16, livae_stream_t stream);
int livae_tc_dgrad_s2blk(const void* gy, const void* wblk, const void* relu_mask, int B, int Hin, int Win, int Cin,
                         int Cout, void* gx, livae_stream_t stream);
/* Linear weight gradient: tensor-core layout [N][(h,w,c)] -> torch [N][(c,h,w)] */
int livae_permute_linear_grad(const float* src_hwc, int N, int C, int HW, float* dst_chw, livae_stream_t stream);
/* Data gradient of a wide nn.Linear (STN fc1, model.py:211; latent heads, model.py:302-303) whose output is only J =
 * 16 / 32 / 64 wide: gx bf16 [B][K] = (mask > 0) * (g bf16 [B][J] @ w^T), w = bf16 [K][J] (livae_tc_pack_weights mode
 * 4), mask = the bf16 [B][K] activation whose ReLU is being differentiated (may be NULL).  HBM-bound: 128-byte row
 * segments in and out (csrc/skinny.cu). */
int livae_linear_dgrad(const void* g, const void* w_kj, const void* mask, int B, int K, int J, void* gx,
                       livae_stream_t stream);

/* ---- thin 1-channel layers around the tensor-core convolutions (SIMT, bf16 on the wide side) ----
 * kind 0: STN conv1 1->16 5x5 p2 +ReLU +MaxPool (model.py:204-206): out bf16 [B,H/2,W/2,16] + pool_idx
 * kind 1: encoder c1 1->32 4x4 s2 p1 +ReLU (model.py:290):          out bf16 [B,H/2,W/2,32]
 * kind 2: data gradient of decoder d4 (32->1 3x3 p0, model.py:371): img = fp32 pre-activation
 *         gradient [B,H,W], out bf16 [B,H+2,W+2,32]
 * Kind 2 and the two livae_thin_convc1_* entry points below are the MATERIALISING form of decoder d4 (they take /
 * produce the up-sampled, reflect-padded tensor).  livae's model path uses livae_upconv_c1_fwd / _bwd instead (no
 * up-sampled tensor); these stay exported for callers that already hold such a tensor and as the cross-check in
 * tests/test_gpu_thin.py. */
int livae_thin_conv1c_fwd(int kind, const float* img, const float* w, const float* bias, int B, int H,
                          int W, void* out_bf16, uint8_t* pool_idx, livae_stream_t stream);
/* weight/bias gradients of kinds 0, 1; g: bf16 PRE-activation gradient (kind 0: pooled, routed by pool_idx) */
int livae_thin_conv1c_wgrad(int kind, const float* img, const void* g_bf16, const uint8_t* pool_idx, int B,
                            int H, int W, float* gw, float* gb, livae_stream_t stream);
/* encoder c1 data gradient: g bf16 [B,H/2,W/2,32] -> gimg fp32 [B,H,W] */
int livae_thin_conv1c_dgrad(const void* g_bf16, const float* w, int B, int H, int W, float* gimg,
                            livae_stream_t stream);
/* decoder d4 forward (32->1 3x3 p0 + activation): x bf16 [B,H,W,32] -> out fp32 [B,H-2,W-2] */
int livae_thin_convc1_fwd(const void* x_bf16, const float* w, const float* bias, int B, int H, int W,
                          int act, float* out, livae_stream_t stream);
/* decoder d4 weight/bias gradient; g fp32 [B,H-2,W-2] pre-activation gradient */
int livae_thin_convc1_wgrad(const void* x_bf16, const float* g, int B, int H, int W, float* gw, float* gb,
                            livae_stream_t stream);
/* out = (g1 + g2) * y * (1 - y); g2 may be NULL */
int livae_sigmoid_bwd(const float* y, const float* g1, const float* g2, int64_t n, float* out,
                      livae_stream_t stream);
/* out_bf16 = g * (y > 0) (y may be NULL: plain cast) */
int livae_relu_mask_cast_bf16(const float* g, const float* y, int64_t n, void* out_bf16, livae_stream_t stream);
int livae_maxpool_bf16(const void* full, int B, int H, int W, int C, void* pooled, uint8_t* idx,
                       livae_stream_t stream);
int livae_unpool_bf16(const void* g_pooled, const uint8_t* idx, int B, int H, int W, int C, void* g_full,
                      livae_stream_t stream);
/* bf16 variants of upsample_pad / decoder fc for the tensor-core path; gradients handed in are
 * PRE-activation gradients, so decfc_bwd_bf16 does not re-apply the ReLU mask */
int livae_upsample_pad_fwd_bf16(const void* x, int B, int H, int W, int C, void* out, livae_stream_t stream);
int livae_upsample_pad_bwd_bf16(const void* g, int B, int H, int W, int C, const void* relu_mask_y, void* gx,
                                livae_stream_t stream);
/* the same + gb[C] (fp32, written) = sum of gx over (b,i,j): the bias gradient of the convolution whose
   pre-activation gradient gx is (autograd of Conv2d bias, model.py:359-367); needs C/8 a power of two, W*C/8 <= 256 */
int livae_upsample_pad_bwd_bias_bf16(const void* g, int B, int H, int W, int C, const void* relu_mask_y, void* gx,
                                     float* gb, livae_stream_t stream);
/* Decoder d4 forward (Upsample x2 bilinear -> ReflectionPad2d(1) -> Conv3x3(32 -> 1) + activation, model.py:369-372)
   from the LOW-resolution input, == livae_upsample_pad_fwd_bf16 + livae_thin_convc1_fwd without the up-sampled
   tensor (and without its bf16 rounding).  x: bf16 [B,H,W,32]; w: fp32 [1,32,3,3]; bias fp32 [1]; out fp32 [B,2H,2W]. */
int livae_upconv_c1_fwd(const void* x_bf16, const float* w, const float* bias, int B, int H, int W, int act,
                        float* out, livae_stream_t stream);
/* The whole backward of decoder d4 (Upsample x2 -> ReflectionPad2d(1) -> Conv3x3(32 -> 1), model.py:369-372) in one
   kernel, replacing livae_thin_convc1_wgrad + livae_thin_conv1c_fwd(kind 2) + livae_upsample_pad_bwd_bias_bf16 and
   both up-sampled [B,2H+2,2W+2,32] tensors.  gpre: fp32 [B,2H,2W] pre-activation gradient of the conv output;
   w: fp32 [1,32,3,3]; x: bf16 [B,H,W,32] post-ReLU LOW-resolution input of the layer.  Written: gx bf16 [B,H,W,32]
   = d/dx times (x > 0); gb_low fp32 [32] (may be NULL) = sum of gx over (b,i,j) (bias gradient of the layer
   below); gw fp32 [1,32,3,3]; gb fp32 [1] (may be NULL) = sum of gpre.  H, W >= 4. */
int livae_upconv_c1_bwd(const float* gpre, const float* w, const void* x_bf16, int B, int H, int W, void* gx_bf16,
                        float* gb_low, float* gw, float* gb, livae_stream_t stream);
/* Last layer of the plain VAE decoder on the tensor-core path: ConvTranspose2d(C -> 1, k4, s2, p1) + activation
   (model.py:95-96, 111-113) and its backward.  x: bf16 [B,H,W,C] (C = 32); w: fp32 [C,1,4,4]; out / g: fp32 [B,2H,2W]
   (g = PRE-activation gradient); gx: bf16 [B,H,W,C] times (relu_mask > 0); gw [C,1,4,4], gb [1] written. */
int livae_thin_convt_c1_fwd(const void* x, const float* w, const float* bias, int B, int H, int W, int C, int act,
                            float* out, livae_stream_t stream);
int livae_thin_convt_c1_dgrad(const float* g, const float* w, const void* relu_mask, int B, int H, int W, int C,
                              void* gx, livae_stream_t stream);
int livae_thin_convt_c1_wgrad(const void* x, const float* g, int B, int H, int W, int C, float* gw, float* gb,
                              livae_stream_t stream);
/* Per-step metric block (train.py:606-667, compute_ssim): mean of the box-filter SSIM map of two fp32 image batches
   [planes = B*C, H, W] -- the five avg_pool2d(win, stride 1, pad win/2, zero padding counted) passes, the map and its
   mean in one pass over the two images.  ws: livae_ssim_box_ws_floats(planes, H) floats; out: 1 float. */
int64_t livae_ssim_box_ws_floats(int64_t planes, int H);
int livae_ssim_box(const float* a, const float* b, int64_t planes, int H, int W, int win, float c1, float c2,
                   float* ws, float* out, livae_stream_t stream);
/* gb[C] (fp32, written) = column sums of the bf16 matrix g[R,C] (bias gradient from a pre-activation gradient) */
int livae_colsum_bf16(const void* g, int64_t R, int C, float* gb, livae_stream_t stream);
int livae_decfc_fwd_bf16(const float* z, const float* w, const float* bias, int B, int L, int C, int HW,
                         void* out, livae_stream_t stream);
int livae_decfc_bwd_bf16(const float* z, const float* w, const void* gy, int B, int L, int C, int HW,
                         float* gw, float* gb, float* gz, livae_stream_t stream);

/* tuning / test hook: 0 = fetch one TMA box per filter tap; 1 (default) = fetch one haloed box per tap
 * group and address each tap as a row shift of it */
void livae_tc_set_halo_mode(int mode);
/* Host-side tile geometry of the halo convolution kernel for an Hq x Wq output grid whose taps shift by up to max_sx
 * columns inside a tap group (kw - 1 at stride 1): box width (16..20), output rows and valid columns per 128-position
 * tile, tiles per map.  No device work. */
int livae_tc_halo_geometry(int Hq, int Wq, int max_sx, int* bw, int* th, int* tw, int* tiles);
void livae_tc_set_wgrad_halo(int mode);   /* the same switch for the weight-gradient kernel */
/* 1 (default): the thin 1-channel layers run on tcgen05 (thread-built im2col / col2im operands,
 * csrc/thin_tc.cu) where the shape is eligible; 0: SIMT kernels only */
void livae_thin_set_tc(int mode);
/* pipeline tracing: device buffer of 4 x 1024 int64 that CTA 0 of a traced kernel fills with
 * (slot << 56 | clock64) records per warp role (tools/probe.py prints the timeline); NULL = off */
void livae_set_probe(void* dev_ptr);
/* dtype conversion between LIVAE_F32 and LIVAE_BF16, n elements */
int livae_cast(const void* src, int dt_src, void* dst, int dt_dst, int64_t n, livae_stream_t stream);

/* Upsample(x2, bilinear, align_corners=False) + ReflectionPad2d(1) (model.py:357-358 etc.):
 * x [B,H,W,C] -> out [B,2H+2,2W+2,C]; the decoder's 3x3 p0 conv then runs on `out`. */
int livae_upsample_pad_fwd(const float* x, int B, int H, int W, int C, float* out, livae_stream_t stream);
/* adjoint (gather form, no atomics).  relu_mask_y (may be NULL): post-ReLU tensor that x was;
 * when given, gx is multiplied by (relu_mask_y > 0), i.e. it is already the pre-activation
 * gradient of the layer that produced x. */
int livae_upsample_pad_bwd(const float* g, int B, int H, int W, int C, const float* relu_mask_y,
                           float* gx, livae_stream_t stream);

/* Decoder.fc / VAEDecoder.fc (model.py:353, 383-384; 84, 108-110):
 * out[b,h,w,c] = relu(z[b,:] . w[(c,h,w),:] + bias[(c,h,w)]), w: torch Linear [C*HW, L]. */
int livae_decfc_fwd(const float* z, const float* w, const float* bias, int B, int L, int C, int HW,
                    float* out, livae_stream_t stream);
/* gy: gradient w.r.t. out (post-ReLU; masked with y>0 on load).  gw [C*HW,L], gb [C*HW], gz [B,L]. */
int livae_decfc_bwd(const float* z, const float* w, const float* y, const float* gy, int B, int L,
                    int C, int HW, float* gw, float* gb, float* gz, livae_stream_t stream);

/* ---- optimiser side (train.py:396-405): global L2 norm of a flat fp32 gradient buffer;
 * out[0] = norm, out[1] = min(1, max_norm/(norm+1e-6)) (torch clip_grad_norm_); apply != 0
 * scales grads in place.  scratch: livae_l2norm_scratch_floats() floats. */
int64_t livae_l2norm_scratch_floats(void);
int livae_l2norm_clip(float* grads, int64_t n, float max_norm, float* out_norm_coef, float* scratch,
                      int apply, livae_stream_t stream);
/* fused Adam/AdamW on flat fp32 buffers (torch.optim semantics; scripts/train_rvae.py:157-159,
 * scripts/train_vae.py:142).  decoupled=1: AdamW.  step_dev: device float step counter, read as
 * t = step+1 and incremented when inc_step != 0; gscale_dev: optional device multiplier applied
 * to the gradient (the clip coefficient). */
int livae_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                float beta2, float eps, float weight_decay, int decoupled, float* step_dev,
                const float* gscale_dev, int inc_step, livae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIVAE_B200_H */

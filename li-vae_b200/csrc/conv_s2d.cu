// STN conv2 (16 -> 32, 5x5 p2, +ReLU +MaxPool2, reference model.py:207-209) in SPACE-TO-DEPTH form.
//
// With 16 input and 32 output channels a 128 x N UMMA is starved: K = 16 per tap and N = 32 leave the
// tensor pipe waiting on its 4 KB A-operand read for every 64 K-MACs (measured: ~40 cycles per MMA, 25
// MMAs per 128 pixels, the same again for the data gradient and twice that for the MN-major weight
// gradient).  Folding each 2x2 pixel block into the channel axis turns the layer into a 3x3 convolution
// over the 32x32 block grid with 64 -> 128 channels (weights zero where a (block offset, phase) pair falls
// outside the 5x5 window: 69 % dense), i.e. 36 MMAs of N = 128 per 512 pixels -- 2.3x fewer tensor-pipe
// cycles per pixel forward, 3-4x fewer for the gradients -- and the 2x2 max-pool becomes a max over four
// column groups of the accumulator row, done in the epilogue: the un-pooled activation never exists.
// The re-blocking costs nothing: a strided TMA tensor map reads the plain NHWC tensor block-wise (HaloOpts).
//
//   forward : x [B,H,W,16] --s2d--> [B,H/2,W/2,64] * W'[9][128][64] -> pooled y [B,H/2,W/2,32] + argmax
//   dgrad   : g (pooled, routed by argmax) --unpool_s2d--> [B,H/2,W/2,128] * W''[9][64][128] -> gx [B,H,W,16]
//   wgrad   : gw'[9][64][128] = sum x_s2d^T g_s2d, folded back to torch's [32][16][5][5]
#include "tc_common.cuh"

namespace livae {
namespace tc {

// mode 0 (forward):  out[tap][n = (dy,dx,co)][k = (ey,ex,ci)] = w[co][ci][ky][kx], ky = 2*DY + ey - dy + 2
// mode 1 (dgrad):    out[tap][n = (ey,ex,ci)][k = (dy,dx,co)] = w[co][ci][ky][kx], ky = ey - dy - 2*DY + 2
// tap = (DY+1)*3 + (DX+1), DY/DX = block offset of the operand read relative to the block written;
// zero where (ky, kx) falls outside the 5x5 filter.
__global__ void pack_s2d_kernel(const float* __restrict__ w, int Co, int Ci, int mode, __nv_bfloat16* __restrict__ out) {
  const int N = mode == 0 ? 4 * Co : 4 * Ci, K = mode == 0 ? 4 * Ci : 4 * Co;
  const int total = 9 * N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % K; int t = i / K; const int n = t % N; const int tap = t / N;
    const int DY = tap / 3 - 1, DX = tap % 3 - 1;
    const int cn = mode == 0 ? Co : Ci, ck = mode == 0 ? Ci : Co;
    const int pn = n / cn, chn = n % cn, pk = k / ck, chk = k % ck;
    const int ny = pn >> 1, nx = pn & 1, ky_ = pk >> 1, kx_ = pk & 1;
    int ky, kx, co, ci;
    if (mode == 0) { ky = 2 * DY + ky_ - ny + 2; kx = 2 * DX + kx_ - nx + 2; co = chn; ci = chk; }
    else { ky = ny - ky_ - 2 * DY + 2; kx = nx - kx_ - 2 * DX + 2; ci = chn; co = chk; }
    float v = 0.f;
    if (ky >= 0 && ky < 5 && kx >= 0 && kx < 5) v = w[((co * Ci + ci) * 5 + ky) * 5 + kx];
    out[i] = __float2bfloat16_rn(v);
  }
}

// out[b,Y,X,(phase,co)] = idx[b,Y,X,co] == phase ? g[b,Y,X,co] : 0      (8 channels per thread)
__global__ void __launch_bounds__(256) unpool_s2d_kernel(const uint4* __restrict__ g, const uint2* __restrict__ idx,
                                                         int64_t npix, int C8, uint4* __restrict__ out) {
  const int64_t n = npix * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C8; const int c = (int)(i - pix * C8);
    const uint4 gv = __ldg(g + i);
    const uint2 iv = __ldg(idx + i);
    const uint8_t* id = reinterpret_cast<const uint8_t*>(&iv);
    const uint16_t* gs = reinterpret_cast<const uint16_t*>(&gv);
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      uint4 o; uint16_t* os = reinterpret_cast<uint16_t*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) os[j] = id[j] == ph ? gs[j] : (uint16_t)0;
      out[(pix * 4 + ph) * C8 + c] = o;
    }
  }
}

// gw[co][ci][ky][kx] = sum over the four output phases (dy,dx) of acc[tap][(ey,ex,ci)][(dy,dx,co)] with
// 2*DY + ey = dy + ky - 2 (same for x); gb[co] = sum over phases of gb4[(dy,dx,co)]
__global__ void fold_s2d_wgrad_kernel(const float* __restrict__ acc, const float* __restrict__ gb4, int Co, int Ci,
                                      float* __restrict__ gw, float* __restrict__ gb) {
  const int total = Co * Ci * 25;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kx = i % 5; int t = i / 5; const int ky = t % 5; t /= 5; const int ci = t % Ci; const int co = t / Ci;
    float s = 0.f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int ry = dy + ky - 2, rx = dx + kx - 2;              // input pixel relative to the block origin
        const int DY = ry >= 0 ? ry >> 1 : -((1 - ry) >> 1), DX = rx >= 0 ? rx >> 1 : -((1 - rx) >> 1);   // floor(r / 2)
        const int ey = ry - 2 * DY, ex = rx - 2 * DX;
        const int tap = (DY + 1) * 3 + (DX + 1);
        s += acc[((int64_t)tap * 4 * Ci + (ey * 2 + ex) * Ci + ci) * 4 * Co + (dy * 2 + dx) * Co + co];
      }
    gw[i] = s;
  }
  if (gb && blockIdx.x == 0)
    for (int co = threadIdx.x; co < Co; co += blockDim.x) gb[co] = gb4[co] + gb4[Co + co] + gb4[2 * Co + co] + gb4[3 * Co + co];
}

// Data gradient of a 4x4 stride-2 pad-1 convolution (encoder c2-c4, model.py:292-296) as ONE 3x3
// convolution over the gy grid that produces whole 2x2 blocks of gx:
//   out[tap = (DY+1)*3 + DX+1][n = (ey,ex,ci)][k = co] = w[co][ci][ky][kx],  ky = ey + 1 - 2*DY, kx = ex + 1 - 2*DX
// (zero outside the 4x4 filter: 4 of 9 taps per output phase).  Compared with the four per-parity launches
// this reads gy once, writes gx in full 128-byte lines and has N = 4*Cin instead of Cin.
__global__ void pack_s2blk_kernel(const float* __restrict__ w, int Co, int Ci, __nv_bfloat16* __restrict__ out) {
  const int N = 4 * Ci, total = 9 * N * Co;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % Co; int t = i / Co; const int n = t % N; const int tap = t / N;
    const int DY = tap / 3 - 1, DX = tap % 3 - 1;
    const int ph = n / Ci, ci = n % Ci;
    const int ky = (ph >> 1) + 1 - 2 * DY, kx = (ph & 1) + 1 - 2 * DX;
    float v = 0.f;
    if (ky >= 0 && ky < 4 && kx >= 0 && kx < 4) v = w[((co * Ci + ci) * 4 + ky) * 4 + kx];
    out[i] = __float2bfloat16_rn(v);
  }
}

static void block_taps(int* tdy, int* tdx, int* tw) {
  for (int t = 0; t < 9; ++t) { tdy[t] = t / 3 - 1; tdx[t] = t % 3 - 1; tw[t] = t; }
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;

static bool s2d_shape_ok(int B, int H, int W, int Ci, int Co) {
  return B > 0 && Ci == 16 && (Co == 16 || Co == 32 || Co == 64) && H >= 16 && W >= 16 && (H % 2) == 0 && (W % 2) == 0;
}

extern "C" int livae_tc_conv5pool_supported(int B, int H, int W, int Ci, int Co) { return s2d_shape_ok(B, H, W, Ci, Co) ? 1 : 0; }

// mode 0: forward weights bf16 [9][4*Co][4*Ci]; mode 1: data-gradient weights bf16 [9][4*Ci][4*Co]
extern "C" int livae_tc_conv5pool_pack(const float* w, int Co, int Ci, int mode, void* out_bf16, livae_stream_t stream) {
  LIVAE_CHECK_ARG(w && out_bf16 && Co > 0 && Ci > 0 && (mode == 0 || mode == 1), "tc_conv5pool_pack: bad args");
  if (int e = require_sm100()) return e;
  const int n = 9 * 16 * Co * Ci;
  pack_s2d_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Co, Ci, mode, (__nv_bfloat16*)out_bf16);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// y[B,H/2,W/2,Co] (bf16) = maxpool2(relu(conv5x5_p2(x[B,H,W,Ci]) + bias)), idx = argmax position 0..3
extern "C" int livae_tc_conv5pool_fwd(const void* x, const void* wpacked0, const float* bias, int B, int H, int W,
                                      int Ci, int Co, void* y, uint8_t* idx, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(s2d_shape_ok(B, H, W, Ci, Co), "tc_conv5pool_fwd: shape not supported");
  LIVAE_CHECK_ARG(x && wpacked0 && y && idx, "tc_conv5pool_fwd: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)wpacked0 | (uintptr_t)y | (uintptr_t)idx) & 15) == 0, "tc_conv5pool_fwd: alignment");
  if (int e = require_sm100()) return e;
  int tdy[9], tdx[9], tw[9];
  block_taps(tdy, tdx, tw);
  const int Hb = H / 2, Wb = W / 2;
  int rc = launch_conv_tc_halo(x, B, Hb, Wb, 4 * Ci, wpacked0, 9, 4 * Co, Hb, Wb, Hb, Wb, 1, 0, 0, 1, 9, tdy, tdx, tw, y, 0,
                               bias, LIVAE_ACT_RELU, nullptr, (cudaStream_t)stream, HaloOpts{1, 1, idx, 0});
  if (rc == 1) { set_error("tc_conv5pool_fwd: shape rejected by the halo kernel"); return -1; }
  return rc;
}

// g_s2d[B,H/2,W/2,4*Co] = pooled gradient routed to its argmax phase
extern "C" int livae_unpool_s2d_bf16(const void* g_pooled, const uint8_t* idx, int B, int Hp, int Wp, int Co, void* g_s2d,
                                     livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g_pooled && idx && g_s2d && (Co & 7) == 0, "unpool_s2d_bf16: bad args");
  LIVAE_CHECK_ARG((((uintptr_t)g_pooled | (uintptr_t)g_s2d) & 15) == 0 && ((uintptr_t)idx & 7) == 0, "unpool_s2d_bf16: alignment");
  if (int e = require_sm100()) return e;
  const int64_t npix = (int64_t)B * Hp * Wp;
  int64_t blocks = (npix * (Co / 8) + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  unpool_s2d_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)g_pooled, (const uint2*)idx, npix, Co / 8,
                                                                 (uint4*)g_s2d);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// gx[B,H,W,Ci] (bf16) = conv5x5 data gradient of g_s2d, times (relu_mask[B,H,W,Ci] > 0)
extern "C" int livae_tc_conv5pool_dgrad(const void* g_s2d, const void* wpacked1, const void* relu_mask, int B, int H, int W,
                                        int Ci, int Co, void* gx, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(s2d_shape_ok(B, H, W, Ci, Co) && 4 * Co >= 64 && (4 * Co) % 64 == 0, "tc_conv5pool_dgrad: shape not supported");
  LIVAE_CHECK_ARG(g_s2d && wpacked1 && gx, "tc_conv5pool_dgrad: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)g_s2d | (uintptr_t)wpacked1 | (uintptr_t)gx | (uintptr_t)relu_mask) & 15) == 0, "tc_conv5pool_dgrad: alignment");
  if (int e = require_sm100()) return e;
  int tdy[9], tdx[9], tw[9];
  block_taps(tdy, tdx, tw);
  const int Hb = H / 2, Wb = W / 2;
  int rc = launch_conv_tc_halo(g_s2d, B, Hb, Wb, 4 * Co, wpacked1, 9, 4 * Ci, Hb, Wb, Hb, Wb, 1, 0, 0, 1, 9, tdy, tdx, tw, gx, 0,
                               nullptr, LIVAE_ACT_NONE, relu_mask, (cudaStream_t)stream, HaloOpts{0, 2, nullptr, 0});
  if (rc == 1) { set_error("tc_conv5pool_dgrad: shape rejected by the halo kernel"); return -1; }
  return rc;
}

extern "C" int64_t livae_tc_conv5pool_wgrad_ws_bytes(int Ci, int Co) { return (int64_t)(9 * 16 * Ci * Co + 4 * Co) * 4; }

// gw fp32 [Co][Ci][5][5], gb fp32 [Co] (may be NULL) from x [B,H,W,Ci] and g_s2d [B,H/2,W/2,4*Co]
extern "C" int livae_tc_conv5pool_wgrad(const void* x, const void* g_s2d, int B, int H, int W, int Ci, int Co, float* gw,
                                        float* gb, void* ws, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(s2d_shape_ok(B, H, W, Ci, Co) && (4 * Co == 32 || (4 * Co) % 64 == 0), "tc_conv5pool_wgrad: shape not supported");
  LIVAE_CHECK_ARG(x && g_s2d && gw && ws, "tc_conv5pool_wgrad: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)g_s2d | (uintptr_t)ws) & 15) == 0, "tc_conv5pool_wgrad: alignment");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int Hb = H / 2, Wb = W / 2;
  float* acc = (float*)ws;
  float* gb4 = acc + (size_t)9 * 16 * Ci * Co;
  cudaError_t ce = cudaMemsetAsync(ws, 0, (size_t)livae_tc_conv5pool_wgrad_ws_bytes(Ci, Co), st);
  if (ce != cudaSuccess) { set_error("tc_conv5pool_wgrad memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  livae_tc_conv_desc d;
  d.B = B; d.Hin = Hb; d.Win = Wb; d.Cin = 4 * Ci; d.Cout = 4 * Co; d.kh = 3; d.kw = 3; d.stride = 1; d.pad = 1; d.act = 0; d.out_f32 = 0;
  int rc = launch_wgrad_halo(&d, x, g_s2d, acc, Hb, Wb, st, 1);
  if (rc == 1) { set_error("tc_conv5pool_wgrad: shape rejected by the halo kernel"); return -1; }
  if (rc != 0) return rc;
  if (gb) colsum_bf16(g_s2d, (int64_t)B * Hb * Wb, 4 * Co, gb4, st);
  const int n = Co * Ci * 25;
  fold_s2d_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(acc, gb4, Co, Ci, gw, gb);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// ---- 4x4 stride-2 pad-1 data gradient in block form (see pack_s2blk_kernel) ----
extern "C" int livae_tc_dgrad_s2blk_supported(int Hin, int Win, int Cin, int Cout) {
  return ((Hin | Win) & 1) == 0 && Hin >= 16 && Win >= 16 && (Cin == 16 || Cin == 32) &&     // N = 4*Cin <= 128
         (Cout == 32 || (Cout % 64) == 0) ? 1 : 0;
}
// w fp32 [Cout][Cin][4][4] -> bf16 [9][4*Cin][Cout]
extern "C" int livae_tc_dgrad_s2blk_pack(const float* w, int Cout, int Cin, void* out_bf16, livae_stream_t stream) {
  LIVAE_CHECK_ARG(w && out_bf16 && Cout > 0 && Cin > 0, "tc_dgrad_s2blk_pack: bad args");
  if (int e = require_sm100()) return e;
  const int n = 9 * 4 * Cin * Cout;
  pack_s2blk_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, (__nv_bfloat16*)out_bf16);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
// gx bf16 [B,Hin,Win,Cin] = data gradient of conv(Cin->Cout, 4x4, s2, p1) at gy bf16 [B,Hin/2,Win/2,Cout], times (relu_mask > 0)
extern "C" int livae_tc_dgrad_s2blk(const void* gy, const void* wblk, const void* relu_mask, int B, int Hin, int Win, int Cin,
                                    int Cout, void* gx, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(livae_tc_dgrad_s2blk_supported(Hin, Win, Cin, Cout), "tc_dgrad_s2blk: shape not supported");
  LIVAE_CHECK_ARG(gy && wblk && gx, "tc_dgrad_s2blk: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)gy | (uintptr_t)wblk | (uintptr_t)gx | (uintptr_t)relu_mask) & 15) == 0, "tc_dgrad_s2blk: alignment");
  if (int e = require_sm100()) return e;
  int tdy[9], tdx[9], tw[9];
  block_taps(tdy, tdx, tw);
  const int Ho = Hin / 2, Wo = Win / 2;
  int rc = launch_conv_tc_halo(gy, B, Ho, Wo, Cout, wblk, 9, 4 * Cin, Ho, Wo, Ho, Wo, 1, 0, 0, 1, 9, tdy, tdx, tw, gx, 0,
                               nullptr, LIVAE_ACT_NONE, relu_mask, (cudaStream_t)stream, HaloOpts{0, 2, nullptr, 0});
  if (rc == 1) { set_error("tc_dgrad_s2blk: shape rejected by the halo kernel"); return -1; }
  return rc;
}

// Geometry shared by both convolution engines.  Everything is expressed in terms of ONE strided
// cross-correlation between a "big" NHWC tensor [B,Hb,Wb,Cb] and a "small" NHWC tensor
// [B,Hs,Ws,Cs] with weights w[Cs][Cb][kh][kw] (torch Conv2d layout; torch ConvTranspose2d's
// [Cin,Cout,kh,kw] is the same array with Cin=Cs, Cout=Cb):
//   P1 (big -> small):  small[b,oy,ox,cs] = sum_{ky,kx,cb} big[b, oy*s-p+ky, ox*s-p+kx, cb] * w[cs,cb,ky,kx]
//   P2 (small -> big):  big[b,iy,ix,cb]   = sum_{ky,kx,cs} small[b,(iy+p-ky)/s,(ix+p-kx)/s,cs] * w[cs,cb,ky,kx]
//   P3 (weights):       gw[cs,cb,ky,kx]   = sum_{b,oy,ox}  big[b, oy*s-p+ky, ox*s-p+kx, cb] * small[b,oy,ox,cs]
// nn.Conv2d:          fwd = P1, dgrad = P2, wgrad = P3(big = x,  small = gy)
// nn.ConvTranspose2d: fwd = P2, dgrad = P1, wgrad = P3(big = gy, small = x)
#pragma once
#include "common.cuh"

namespace livae {

struct ConvGeom {
  int B, Hb, Wb, Cb;   // big side
  int Hs, Ws, Cs;      // small side (before any pooling)
  int kh, kw, stride, pad;
};

// A tensor operand, optionally seen through the derivative of the activation (and the 2x2
// max-pool routing) of the layer that produced it: value = g * act'(y) [* (argmax == pos)].
struct TensorRef {
  const float* p;         // data (plain operand) or incoming gradient g
  const float* yact;      // post-activation output of the layer (nullptr => plain)
  const uint8_t* pidx;    // max-pool argmax position 0..3 (nullptr => no pool)
  int act;                // LIVAE_ACT_*
};

__device__ __forceinline__ float apply_dact(float g, float y, int act) {
  if (act == LIVAE_ACT_RELU) return y > 0.f ? g : 0.f;
  if (act == LIVAE_ACT_SIGMOID) return g * y * (1.f - y);
  return g;
}

// element (b, y, x, c) of an [B,H,W,C] operand
__device__ __forceinline__ float ref_load(const TensorRef& t, int b, int y, int x, int c, int H, int W,
                                          int C) {
  if (t.pidx) {  // pooled storage [B,H/2,W/2,C]
    int64_t i = (((int64_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * C + c;
    if (t.pidx[i] != (((y & 1) << 1) | (x & 1))) return 0.f;
    return apply_dact(t.p[i], t.yact[i], t.act);
  }
  int64_t i = (((int64_t)b * H + y) * W + x) * C + c;
  float g = t.p[i];
  return t.yact ? apply_dact(g, t.yact[i], t.act) : g;
}

}  // namespace livae

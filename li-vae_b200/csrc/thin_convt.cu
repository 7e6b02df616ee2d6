// Last layer of the plain VAE decoder (reference model.py:95-96, 111-113): ConvTranspose2d(C -> 1, k4, s2, p1) +
// Sigmoid, and its backward, for the tensor-core path (bf16 NHWC input [B,H,W,C], fp32 image [B,2H,2W] out).
// N = 1 output channel: no tensor-core shape; the layer is bound by reading the C-channel input once
// (C*2 bytes per input pixel against 512 MACs), so these are plain CUDA-core kernels with 16-byte loads.
//
//   out[b, oy, ox] = act(bias + sum_{ci, ky, kx} x[b, iy, ix, ci] * w[ci, 0, ky, kx]),  oy = 2*iy - 1 + ky
// Per 2x2 output block (by, bx): output row 2*by + py reads input rows by (ky = py + 1) and by - 1 + 2*py
// (ky = 3 - 3*py), same in x: a block needs the 3x3 input neighbourhood and each of its 4 pixels uses 2x2 of it.
#include "common.cuh"

namespace livae {

__device__ __forceinline__ void bf8_unpack(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

// ---- forward: one thread per 2x2 output block -------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) convt_c1_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int B, int H, int W, int act,
                                                           float* __restrict__ out) {
  __shared__ float sw[C * 16];                 // [ci][ky][kx]
  for (int i = threadIdx.x; i < C * 16; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  constexpr int C8 = C / 8;
  const float b0 = bias ? bias[0] : 0.f;
  const int64_t nblk = (int64_t)B * H * W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nblk; t += (int64_t)gridDim.x * blockDim.x) {
    const int bx = (int)(t % W); const int64_t r = t / W; const int by = (int)(r % H); const int64_t b = r / H;
    float acc[2][2] = {{b0, b0}, {b0, b0}};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int iy = by + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ix = bx + dx;
        if (ix < 0 || ix >= W) continue;
        const uint4* px = x + ((b * H + iy) * W + ix) * C8;
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) {
          float v[8];
          bf8_unpack(__ldg(px + c8), v);
          // input row by + dy feeds output row 2*by + py with ky = py + 1 - 2*dy (when 0 <= ky < 4)
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            const int ky = py + 1 - 2 * dy;
            if (ky < 0 || ky > 3) continue;
#pragma unroll
            for (int pxx = 0; pxx < 2; ++pxx) {
              const int kx = pxx + 1 - 2 * dx;
              if (kx < 0 || kx > 3) continue;
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[py][pxx] = fmaf(v[e], sw[(c8 * 8 + e) * 16 + ky * 4 + kx], acc[py][pxx]);
            }
          }
        }
      }
    }
    float* o = out + (b * 2 * H + 2 * by) * (2 * W) + 2 * bx;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      float a0 = acc[py][0], a1 = acc[py][1];
      if (act == LIVAE_ACT_SIGMOID) { a0 = 1.f / (1.f + __expf(-a0)); a1 = 1.f / (1.f + __expf(-a1)); }
      else if (act == LIVAE_ACT_RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
      *reinterpret_cast<float2*>(o + py * 2 * W) = make_float2(a0, a1);
    }
  }
}

// ---- data gradient: one thread per (input pixel, 8 channels) ------------------------------------------
//   gx[b, iy, ix, ci] = (mask > 0) * sum_{ky, kx} g[b, 2*iy - 1 + ky, 2*ix - 1 + kx] * w[ci, ky, kx]
template <int C>
__global__ void __launch_bounds__(256) convt_c1_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                             const uint4* __restrict__ mask, int B, int H, int W,
                                                             uint4* __restrict__ gx) {
  __shared__ float sw[C * 16];
  for (int i = threadIdx.x; i < C * 16; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  constexpr int C8 = C / 8;
  const int64_t n = (int64_t)B * H * W * C8;
  const int Ho = 2 * H, Wo = 2 * W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(t % C8); int64_t r = t / C8;
    const int ix = (int)(r % W); r /= W; const int iy = (int)(r % H); const int64_t b = r / H;
    float win[16];
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int oy = 2 * iy - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ox = 2 * ix - 1 + kx;
        win[ky * 4 + kx] = (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) ? __ldg(g + (b * Ho + oy) * Wo + ox) : 0.f;
      }
    }
    float m[8], o[8];
    if (mask) bf8_unpack(__ldg(mask + t), m);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float* wc = sw + (c8 * 8 + e) * 16;
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) a = fmaf(win[k], wc[k], a);
      o[e] = (mask && !(m[e] > 0.f)) ? 0.f : a;
    }
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
    gx[t] = q;
  }
}

// ---- weight + bias gradient ----------------------------------------------------------------------------
//   gw[ci, ky, kx] = sum_{b, iy, ix} x[b, iy, ix, ci] * g[b, 2*iy - 1 + ky, 2*ix - 1 + kx];  gb = sum g
// Thread (ci, ky) keeps the four kx accumulators over the CTA's strip of input ROWS: per pixel a warp reads the
// pixel's C channels as one coalesced line and four gradient values as broadcasts; the row loop has no index
// arithmetic, so the compiler keeps several pixels' loads in flight.  CTA = C * 4 threads.
template <int C>
__global__ void __launch_bounds__(C * 4) convt_c1_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ g,
                                                               int B, int H, int W, int rows_per_cta,
                                                               float* __restrict__ gw, float* __restrict__ gb) {
  const int ci = threadIdx.x % C, ky = threadIdx.x / C;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t nrows = (int64_t)B * H;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < nrows ? r0 + rows_per_cta : nrows;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, sb = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t b = r / H; const int iy = (int)(r - b * H);
    const int oy = 2 * iy - 1 + ky;
    const __nv_bfloat16* xr = x + r * W * C + ci;
    if (oy >= 0 && oy < Ho) {
      const float* gr = g + (b * Ho + oy) * Wo;
      {   // ix = 0: output column -1 does not exist
        const float xv = __bfloat162float(xr[0]);
        a[1] = fmaf(xv, __ldg(gr), a[1]); a[2] = fmaf(xv, __ldg(gr + 1), a[2]); a[3] = fmaf(xv, __ldg(gr + 2), a[3]);
      }
#pragma unroll 4
      for (int ix = 1; ix < W - 1; ++ix) {
        const float xv = __bfloat162float(xr[ix * C]);
        const float* gp = gr + 2 * ix - 1;
        a[0] = fmaf(xv, __ldg(gp), a[0]); a[1] = fmaf(xv, __ldg(gp + 1), a[1]);
        a[2] = fmaf(xv, __ldg(gp + 2), a[2]); a[3] = fmaf(xv, __ldg(gp + 3), a[3]);
      }
      if (W > 1) {   // ix = W-1: output column 2W does not exist
        const float xv = __bfloat162float(xr[(W - 1) * C]);
        const float* gp = gr + 2 * W - 3;
        a[0] = fmaf(xv, __ldg(gp), a[0]); a[1] = fmaf(xv, __ldg(gp + 1), a[1]); a[2] = fmaf(xv, __ldg(gp + 2), a[2]);
      }
    }
    // bias gradient: the two output rows of this input row, summed by the threads of ky = 0 (C lanes stride the row)
    if (gb && ky == 0)
      for (int k = ci; k < 2 * Wo; k += C) sb += __ldg(g + (b * Ho + 2 * iy) * Wo + k);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(gw + ci * 16 + ky * 4 + k, a[k]);
  if (gb && ky == 0) {
    sb = warp_sum(sb);
    if ((threadIdx.x & 31) == 0) atomicAdd(gb, sb);
  }
}

}  // namespace livae

using namespace livae;

// x: bf16 [B,H,W,C] (C = 32), w: fp32 [C,1,4,4] (torch ConvTranspose2d layout), out: fp32 [B,2H,2W]
extern "C" int livae_thin_convt_c1_fwd(const void* x, const float* w, const float* bias, int B, int H, int W, int C, int act,
                                       float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0, "thin_convt_c1_fwd: bad sizes");
  LIVAE_CHECK_ARG(C == 32, "thin_convt_c1_fwd: C must be 32 (got %d)", C);
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x && w && out && (((uintptr_t)x | (uintptr_t)out) & 15) == 0, "thin_convt_c1_fwd: null or misaligned pointer");
  if (int e = require_sm100()) return e;
  const int64_t nblk = (int64_t)B * H * W;
  int64_t blocks = (nblk + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  convt_c1_fwd_kernel<32><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, w, bias, B, H, W, act, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// g: fp32 [B,2H,2W] PRE-activation gradient; gx: bf16 [B,H,W,C] = data gradient times (relu_mask > 0) (mask optional)
extern "C" int livae_thin_convt_c1_dgrad(const float* g, const float* w, const void* relu_mask, int B, int H, int W, int C,
                                         void* gx, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0, "thin_convt_c1_dgrad: bad sizes");
  LIVAE_CHECK_ARG(C == 32, "thin_convt_c1_dgrad: C must be 32 (got %d)", C);
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g && w && gx && (((uintptr_t)gx | (uintptr_t)relu_mask) & 15) == 0, "thin_convt_c1_dgrad: null or misaligned pointer");
  if (int e = require_sm100()) return e;
  const int64_t n = (int64_t)B * H * W * (C / 8);
  int64_t blocks = (n + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  convt_c1_dgrad_kernel<32><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, w, (const uint4*)relu_mask, B, H, W, (uint4*)gx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// gw: fp32 [C,1,4,4], gb: fp32 [1] (optional); both written
extern "C" int livae_thin_convt_c1_wgrad(const void* x, const float* g, int B, int H, int W, int C, float* gw, float* gb,
                                         livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 1, "thin_convt_c1_wgrad: bad sizes (W >= 2)");
  LIVAE_CHECK_ARG(C == 32, "thin_convt_c1_wgrad: C must be 32 (got %d)", C);
  LIVAE_CHECK_ARG(gw, "thin_convt_c1_wgrad: null gw");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemsetAsync(gw, 0, (size_t)C * 16 * sizeof(float), st);
  if (ce == cudaSuccess && gb) ce = cudaMemsetAsync(gb, 0, sizeof(float), st);
  if (ce != cudaSuccess) { set_error("thin_convt_c1_wgrad memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x && g, "thin_convt_c1_wgrad: null pointer");
  const int64_t nrows = (int64_t)B * H;
  int64_t ctas = kNumSMs * 16;
  if (ctas > nrows) ctas = nrows;
  const int64_t per = (nrows + ctas - 1) / ctas;
  ctas = (nrows + per - 1) / per;
  convt_c1_wgrad_kernel<32><<<(int)ctas, 128, 0, st>>>((const __nv_bfloat16*)x, g, B, H, W, (int)per, gw, gb);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// Thin layers around the tensor-core convolutions: the 1-channel ends of the networks
//   STN conv1  1->16 5x5 p2 +ReLU +MaxPool   (model.py:204-206)
//   encoder c1 1->32 4x4 s2 p1 +ReLU          (model.py:290)
//   decoder d4 32->1 3x3 p0 +Sigmoid          (model.py:371-372)
// Their GEMM shapes have N = 1 or K <= 25, useless for a 128xN MMA tile, and they are bound by the
// bf16 activation traffic on the C-channel side, so they are SIMT kernels: fp32 FMA, bf16 NHWC
// storage on the wide side, fp32 images on the 1-channel side.  Gradient conventions follow the
// tensor-core engine: tensors handed between layers are PRE-activation gradients.
#include "common.cuh"
#include "thin_tc.cuh"

namespace livae {

__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------------------------
// 1 -> C convolution forward (also: data gradient of a C -> 1 convolution, with flip = 1).
// img fp32 [B,H,W]; w fp32 [C][K*K] (torch [C,1,K,K] or [1,C,K,K]); out bf16 NHWC.
// POOL: ReLU + 2x2 max-pool fused, idx = argmax position (torch scan order).
template <int C, int K, int S, bool POOL>
__global__ void __launch_bounds__(256) conv1c_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int H, int W,
                                                         int Ho, int Wo, int pad, int act, int flip,
                                                         __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ idx) {
  __shared__ float sw[K * K][C];
  __shared__ float sb[C];
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) {
    int c = i % C, t = i / C;
    sw[t][c] = w[c * K * K + (flip ? K * K - 1 - t : t)];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const float* im = img + (int64_t)b * H * W;
  if (POOL) {
    const int Hp = Ho >> 1, Wp = Wo >> 1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Hp * Wp; q += gridDim.x * blockDim.x) {
      const int py = q / Wp, px = q - py * Wp;
      float best[C]; int bi[C];
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const int oy = 2 * py + (sub >> 1), ox = 2 * px + (sub & 1);
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = sb[c];
        for (int ky = 0; ky < K; ++ky) {
          const int iy = oy * S - pad + ky;
          if (iy < 0 || iy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const int ix = ox * S - pad + kx;
            if (ix < 0 || ix >= W) continue;
            const float v = __ldg(im + iy * W + ix);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] = fmaf(v, sw[ky * K + kx][c], acc[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float a = fmaxf(acc[c], 0.f);
          if (sub == 0 || a > best[c]) { best[c] = a; bi[c] = sub; }
        }
      }
      __nv_bfloat16* o = out + (((int64_t)b * Hp + py) * Wp + px) * C;
      uint8_t* oi = idx + (((int64_t)b * Hp + py) * Wp + px) * C;
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(best[c], best[c + 1]);
        oi[c] = (uint8_t)bi[c]; oi[c + 1] = (uint8_t)bi[c + 1];
      }
    }
  } else {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < Ho * Wo; q += gridDim.x * blockDim.x) {
      const int oy = q / Wo, ox = q - oy * Wo;
      float acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = sb[c];
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S - pad + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox * S - pad + kx;
          if (ix < 0 || ix >= W) continue;
          const float v = __ldg(im + iy * W + ix);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(v, sw[ky * K + kx][c], acc[c]);
        }
      }
      __nv_bfloat16* o = out + ((int64_t)b * Ho * Wo + q) * C;
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        float a0 = acc[c], a1 = acc[c + 1];
        if (act == LIVAE_ACT_RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(a0, a1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 1 -> C convolution weight (+bias) gradient: gw[c][tap] = sum img[b, oy*S-pad+ky, ox*S-pad+kx] * g[b,oy,ox,c]
// g bf16: [B,Ho,Wo,C] pre-activation gradient, or (POOL) the pooled gradient [B,Ho/2,Wo/2,C] routed by idx.
// CTA = a band of RB output rows, walking over images.  Per image the band's zero-padded image rows
// (fp32) and its gradient rows (bf16, pool routing resolved while staging) are staged in shared memory
// with coalesced loads; thread = (tap row ky, 4 channels, output row) runs along x with K*4 register
// accumulators (K + 1 shared loads per K*4 FMAs).  One shared reduction + one global atomic per
// weight per CTA at the very end.
template <int C, int K, int S, bool POOL>
__global__ void __launch_bounds__(256) conv1c_wgrad_kernel(const float* __restrict__ img,
                                                           const __nv_bfloat16* __restrict__ g,
                                                           const uint8_t* __restrict__ idx, int B, int H, int W, int Ho,
                                                           int Wo, int pad, float* __restrict__ gw,
                                                           float* __restrict__ gb) {
  constexpr int CG = C / 4;
  constexpr int COMBOS = K * CG;
  constexpr int RB = 256 / COMBOS;          // output rows per band
  constexpr int IR = (RB - 1) * S + K;      // staged image rows
  extern __shared__ __align__(16) uint8_t dsm[];
  const int IWp = (Wo - 1) * S + K;         // staged image row width (zero padded)
  float* simg = reinterpret_cast<float*>(dsm);                                  // [IR][IWp]
  __nv_bfloat16* sg = reinterpret_cast<__nv_bfloat16*>(simg + ((IR * IWp + 3) & ~3));   // [RB][Wo][C]
  __shared__ float sacc[K * K][C];
  __shared__ float sbias[C];
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) (&sacc[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sbias[i] = 0.f;
  const int t = threadIdx.x;
  const int ky = t % K, cg = (t / K) % CG, r = t / COMBOS;
  const int oy0 = blockIdx.x * RB;
  const int nrows = min(RB, Ho - oy0);
  const bool active = r < nrows;
  float acc[K][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int iy0 = oy0 * S - pad;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    __syncthreads();
    // stage image rows [iy0, iy0+IR) x [-pad, -pad+IWp), zero outside
    const float* im = img + (int64_t)b * H * W;
    for (int i = t; i < IR * IWp; i += blockDim.x) {
      const int rr = i / IWp, cc = i - rr * IWp;
      const int iy = iy0 + rr, ix = cc - pad;
      simg[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(im + iy * W + ix) : 0.f;
    }
    // stage gradient rows (4 channels = 8 bytes per item)
    for (int i = t; i < nrows * Wo * CG; i += blockDim.x) {
      const int c4 = i % CG; const int q = i / CG; const int ox = q % Wo; const int rr = q / Wo;
      const int oy = oy0 + rr;
      uint2 v;
      if (POOL) {
        const int64_t pi = ((((int64_t)b * (Ho >> 1) + (oy >> 1)) * (Wo >> 1)) + (ox >> 1)) * C + c4 * 4;
        const int pos = ((oy & 1) << 1) | (ox & 1);
        const uint32_t id4 = __ldg(reinterpret_cast<const uint32_t*>(idx + pi));
        v = __ldg(reinterpret_cast<const uint2*>(g + pi));
        if ((int)(id4 & 0xff) != pos) v.x &= 0xffff0000u;
        if ((int)((id4 >> 8) & 0xff) != pos) v.x &= 0x0000ffffu;
        if ((int)((id4 >> 16) & 0xff) != pos) v.y &= 0xffff0000u;
        if ((int)((id4 >> 24) & 0xff) != pos) v.y &= 0x0000ffffu;
      } else {
        v = __ldg(reinterpret_cast<const uint2*>(g + (((int64_t)b * Ho + oy) * Wo + ox) * C + c4 * 4));
      }
      *reinterpret_cast<uint2*>(sg + ((int64_t)(rr * Wo + ox)) * C + c4 * 4) = v;
    }
    __syncthreads();
    if (active) {
      const float* irow = simg + (r * S + ky) * IWp;
      const __nv_bfloat16* grow = sg + (int64_t)r * Wo * C + cg * 4;
#pragma unroll 2
      for (int ox = 0; ox < Wo; ++ox) {
        const uint2 gv = *reinterpret_cast<const uint2*>(grow + (int64_t)ox * C);
        const __nv_bfloat162* gp = reinterpret_cast<const __nv_bfloat162*>(&gv);
        const float2 g01 = __bfloat1622float2(gp[0]), g23 = __bfloat1622float2(gp[1]);
        if (ky == 0) { accb[0] += g01.x; accb[1] += g01.y; accb[2] += g23.x; accb[3] += g23.y; }
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float v = irow[ox * S + kx];
          acc[kx][0] = fmaf(v, g01.x, acc[kx][0]);
          acc[kx][1] = fmaf(v, g01.y, acc[kx][1]);
          acc[kx][2] = fmaf(v, g23.x, acc[kx][2]);
          acc[kx][3] = fmaf(v, g23.y, acc[kx][3]);
        }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int kx = 0; kx < K; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sacc[ky * K + kx][cg * 4 + j], acc[kx][j]);
    if (ky == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sbias[cg * 4 + j], accb[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) {
    int c = i % C, tp = i / C;
    atomicAdd(gw + c * K * K + tp, sacc[tp][c]);
  }
  if (gb)
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(gb + i, sbias[i]);
}

// ---------------------------------------------------------------------------------------------
// C -> 1 convolution forward, stride 1 (decoder d4): out[b,oy,ox] = act(sum_{tap,c} x[b,oy+ky,ox+kx,c] w[c][tap] + bias)
// x bf16 [B,H,W,C] (already padded: pad 0), w fp32 [1][C][K][K], out fp32 [B,Ho,Wo].
template <int C, int K>
__global__ void __launch_bounds__(256) convc1_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int H, int W, int Ho,
                                                         int Wo, int act, float* __restrict__ out) {
  __shared__ float sw[K * K][C];
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) { int c = i % C, t = i / C; sw[t][c] = w[c * K * K + t]; }
  __syncthreads();
  const float b0 = bias ? bias[0] : 0.f;
  const int64_t n = (int64_t)B * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo); int64_t q = i / Wo; const int oy = (int)(q % Ho); const int b = (int)(q / Ho);
    float acc = b0;
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const uint4* p = reinterpret_cast<const uint4*>(x + (((int64_t)b * H + oy + ky) * W + ox + kx) * C);
#pragma unroll
        for (int v = 0; v < C / 8; ++v) {
          const uint4 u = __ldg(p + v);
          const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc = fmaf(bf(e[j]), sw[ky * K + kx][v * 8 + j], acc);
        }
      }
    if (act == LIVAE_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    else if (act == LIVAE_ACT_RELU) acc = fmaxf(acc, 0.f);
    out[i] = acc;
  }
}

// C -> 1 convolution weight (+bias) gradient: gw[c][tap] = sum x[b,oy+ky,ox+kx,c] * g[b,oy,ox]  (g fp32, pre-activation)
// CTA = a band of RWB output rows, walking over images; the band's (RWB+K-1) input rows (bf16, all C
// channels) and gradient rows (fp32) are staged in shared memory with 16-byte coalesced loads.
// thread = (tap, channel pair, row-slice): one bf16x2 shared load feeds two FMAs.
template <int C, int K, int RWB>
__global__ void __launch_bounds__(K* K*(C / 2) * 2) convc1_wgrad_kernel(const __nv_bfloat16* __restrict__ x,
                                                                        const float* __restrict__ g, int B, int H,
                                                                        int W, int Ho, int Wo, float* __restrict__ gw,
                                                                        float* __restrict__ gb) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int NT = K * K * (C / 2) * 2;
  __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(dsm);                 // [RWB+K-1][W][C]
  float* sg = reinterpret_cast<float*>(sx + (size_t)(RWB + K - 1) * W * C);  // [RWB][Wo]
  const int t = threadIdx.x;
  const int half = t & 1;
  const int c2 = (t >> 1) % (C / 2);
  const int tap = (t >> 1) / (C / 2);
  const int ky = tap / K, kx = tap % K;
  const int oy0 = blockIdx.x * RWB;
  const int nrows = min(RWB, Ho - oy0);
  float a0 = 0.f, a1 = 0.f, accb = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(x + ((int64_t)b * H + oy0) * W * C);
    const int n16 = (nrows + K - 1) * W * C / 8;
    for (int i = t; i < n16; i += NT) reinterpret_cast<uint4*>(sx)[i] = __ldg(src + i);
    for (int i = t; i < nrows * Wo; i += NT) {
      const float v = g[((int64_t)b * Ho + oy0) * Wo + i];
      sg[i] = v;
      accb += v;
    }
    __syncthreads();
    for (int r = half; r < nrows; r += 2) {
      const __nv_bfloat16* xr = sx + ((size_t)(r + ky) * W + kx) * C + 2 * c2;
      const float* gr = sg + r * Wo;
#pragma unroll 4
      for (int ox = 0; ox < Wo; ++ox) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + (size_t)ox * C));
        const float gg = gr[ox];
        a0 = fmaf(v.x, gg, a0);
        a1 = fmaf(v.y, gg, a1);
      }
    }
  }
  atomicAdd(gw + (2 * c2) * K * K + tap, a0);
  atomicAdd(gw + (2 * c2 + 1) * K * K + tap, a1);
  if (gb) {
    accb = warp_sum(accb);
    if ((t & 31) == 0) atomicAdd(gb, accb);
  }
}

// Data gradient of a 1 -> C strided convolution (encoder c1): gimg[b,iy,ix] = sum_{ky,kx,c} g[b,oy,ox,c] w[c][ky][kx]
// with oy = (iy+pad-ky)/S.  g bf16 [B,Ho,Wo,C] pre-activation gradient; gimg fp32 [B,H,W].
template <int C, int K, int S>
__global__ void __launch_bounds__(256) conv1c_dgrad_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ w,
                                                           int B, int H, int W, int Ho, int Wo, int pad,
                                                           float* __restrict__ gimg) {
  __shared__ float sw[K * K][C];
  for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) { int c = i % C, t = i / C; sw[t][c] = w[c * K * K + t]; }
  __syncthreads();
  const int64_t n = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % W); int64_t q = i / W; const int iy = (int)(q % H); const int b = (int)(q / H);
    float acc = 0.f;
    for (int ky = 0; ky < K; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % S != 0) continue;
      const int oy = ty / S;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % S != 0) continue;
        const int ox = tx / S;
        if (ox >= Wo) continue;
        const uint4* p = reinterpret_cast<const uint4*>(g + (((int64_t)b * Ho + oy) * Wo + ox) * C);
#pragma unroll
        for (int v = 0; v < C / 8; ++v) {
          const uint4 u = __ldg(p + v);
          const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc = fmaf(bf(e[j]), sw[ky * K + kx][v * 8 + j], acc);
        }
      }
    }
    gimg[i] = acc;
  }
}

// out = (g1 + g2) * y * (1 - y): pre-activation gradient of the decoder's sigmoid (model.py:372)
__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ g1,
                                                          const float* __restrict__ g2, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = y[i], g = g1[i] + (g2 ? g2[i] : 0.f);
    out[i] = g * v * (1.f - v);
  }
}

// out_bf16 = g * (y > 0): pre-activation gradient of a small fp32 ReLU layer, cast for the tensor-core path
__global__ void __launch_bounds__(256) relu_mask_cast_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                             int64_t n, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn((!y || y[i] > 0.f) ? g[i] : 0.f);
}

// ---- bf16 max-pool / un-pool ------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_bf16_kernel(const __nv_bfloat16* __restrict__ full, int B, int H, int W,
                                                           int C, __nv_bfloat16* __restrict__ pooled,
                                                           uint8_t* __restrict__ idx) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t n = (int64_t)B * Hp * Wp * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t q = i / C; const int px = (int)(q % Wp); q /= Wp; const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    const __nv_bfloat16* s = full + (((int64_t)b * H + 2 * py) * W + 2 * px) * C + c;
    float best = bf(s[0]); int bi = 0;
    float v = bf(s[C]); if (v > best) { best = v; bi = 1; }
    v = bf(s[(int64_t)W * C]); if (v > best) { best = v; bi = 2; }
    v = bf(s[(int64_t)W * C + C]); if (v > best) { best = v; bi = 3; }
    pooled[i] = __float2bfloat16_rn(best);
    idx[i] = (uint8_t)bi;
  }
}
// full[b,y,x,c] = (idx[pooled] == pos) ? g[pooled] : 0
__global__ void __launch_bounds__(256) unpool_bf16_kernel(const __nv_bfloat16* __restrict__ g, const uint8_t* __restrict__ idx,
                                                          int B, int H, int W, int C, __nv_bfloat16* __restrict__ full) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t n = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t q = i / C; const int x = (int)(q % W); q /= W; const int y = (int)(q % H);
    const int b = (int)(q / H);
    const int64_t pi = (((int64_t)b * Hp + (y >> 1)) * Wp + (x >> 1)) * C + c;
    full[i] = idx[pi] == (((y & 1) << 1) | (x & 1)) ? g[pi] : __float2bfloat16_rn(0.f);
  }
}

// 8 channels (one 16-byte vector) per thread
__global__ void __launch_bounds__(256) maxpool_bf16v_kernel(const uint4* __restrict__ full, int B, int H, int W, int C8,
                                                            uint4* __restrict__ pooled, uint2* __restrict__ idx) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t n = (int64_t)B * Hp * Wp * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8); int64_t q = i / C8; const int px = (int)(q % Wp); q /= Wp; const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    const uint4* s = full + (((int64_t)b * H + 2 * py) * W + 2 * px) * C8 + c;
    uint4 v[4] = {__ldg(s), __ldg(s + C8), __ldg(s + (int64_t)W * C8), __ldg(s + (int64_t)W * C8 + C8)};
    uint4 o; uint8_t id[8];
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float best = bf(reinterpret_cast<const __nv_bfloat16*>(&v[0])[j]); int bi = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        float t = bf(reinterpret_cast<const __nv_bfloat16*>(&v[k])[j]);
        if (t > best) { best = t; bi = k; }
      }
      ob[j] = __float2bfloat16_rn(best);
      id[j] = (uint8_t)bi;
    }
    pooled[i] = o;
    idx[i] = *reinterpret_cast<uint2*>(id);
  }
}
__global__ void __launch_bounds__(256) unpool_bf16v_kernel(const uint4* __restrict__ g, const uint2* __restrict__ idx,
                                                           int B, int H, int W, int C8, uint4* __restrict__ full) {
  const int Hp = H >> 1, Wp = W >> 1;
  const int64_t n = (int64_t)B * Hp * Wp * C8;   // one thread per pooled vector, writes its 2x2 window
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8); int64_t q = i / C8; const int px = (int)(q % Wp); q /= Wp; const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    const uint4 gv = __ldg(g + i);
    const uint2 iv = __ldg(idx + i);
    const uint8_t* id = reinterpret_cast<const uint8_t*>(&iv);
    const uint16_t* gs = reinterpret_cast<const uint16_t*>(&gv);
    uint4* d = full + (((int64_t)b * H + 2 * py) * W + 2 * px) * C8 + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 o; uint16_t* os = reinterpret_cast<uint16_t*>(&o);
#pragma unroll
      for (int j = 0; j < 8; ++j) os[j] = id[j] == k ? gs[j] : (uint16_t)0;
      d[(int64_t)(k >> 1) * W * C8 + (k & 1) * C8] = o;
    }
  }
}

static inline int tgrid(int64_t n, int per = 1) {
  int64_t blocks = (n + 256LL * per - 1) / (256LL * per);
  if (blocks < 1) blocks = 1;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace livae

using namespace livae;

// 1 (default): tensor-core versions (thin_tc.cu) where the shape is eligible; 0: SIMT kernels only
static int g_thin_tc = 1;
extern "C" void livae_thin_set_tc(int mode) { g_thin_tc = mode ? 1 : 0; }

// kind: 0 = STN conv1 (C=16,K=5,S=1,+ReLU+pool), 1 = encoder c1 (C=32,K=4,S=2,+ReLU),
//       2 = decoder d4 data gradient (C=32,K=3,S=1, flipped taps, no activation)
extern "C" int livae_thin_conv1c_fwd(int kind, const float* img, const float* w, const float* bias, int B, int H,
                                     int W, void* out_bf16, uint8_t* pool_idx, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && kind >= 0 && kind <= 2, "thin_conv1c_fwd: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && w && out_bf16, "thin_conv1c_fwd: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
  LIVAE_CHECK_ARG(kind != 0 || pool_idx, "thin_conv1c_fwd: STN conv1 needs pool_idx");
  if (g_thin_tc) {
    int rc = tc::thin_tc_conv1c_fwd(kind, img, w, bias, B, H, W, out_bf16, pool_idx, st);
    if (rc != 1) return rc;
  }
  if (kind == 0) {
    LIVAE_CHECK_ARG(pool_idx && (H % 2) == 0 && (W % 2) == 0, "thin_conv1c_fwd: STN conv1 needs pool_idx and even size");
    int quads = (H / 2) * (W / 2);
    dim3 grid((quads + 255) / 256, B);
    conv1c_fwd_kernel<16, 5, 1, true><<<grid, 256, 0, st>>>(img, w, bias, B, H, W, H, W, 2, LIVAE_ACT_RELU, 0, o, pool_idx);
  } else if (kind == 1) {
    LIVAE_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "thin_conv1c_fwd: encoder c1 needs even size");
    int px = (H / 2) * (W / 2);
    dim3 grid((px + 255) / 256, B);
    conv1c_fwd_kernel<32, 4, 2, false><<<grid, 256, 0, st>>>(img, w, bias, B, H, W, H / 2, W / 2, 1, LIVAE_ACT_RELU, 0, o, nullptr);
  } else {
    int px = (H + 2) * (W + 2);
    dim3 grid((px + 255) / 256, B);
    conv1c_fwd_kernel<32, 3, 1, false><<<grid, 256, 0, st>>>(img, w, nullptr, B, H, W, H + 2, W + 2, 2, LIVAE_ACT_NONE, 1, o, nullptr);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// kind 0: STN conv1 (g = pooled pre-activation gradient bf16 [B,H/2,W/2,16] + pool_idx);
// kind 1: encoder c1 (g bf16 [B,H/2,W/2,32]).  gw [C][K*K], gb [C] fp32, written.
extern "C" int livae_thin_conv1c_wgrad(int kind, const float* img, const void* g_bf16, const uint8_t* pool_idx, int B,
                                       int H, int W, float* gw, float* gb, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (kind == 0 || kind == 1), "thin_conv1c_wgrad: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && g_bf16 && gw, "thin_conv1c_wgrad: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* g = (const __nv_bfloat16*)g_bf16;
  cudaError_t ce;
  if (kind == 0) {
    LIVAE_CHECK_ARG(pool_idx, "thin_conv1c_wgrad: STN conv1 needs pool_idx");
    if ((ce = cudaMemsetAsync(gw, 0, 16 * 25 * 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (gb && (ce = cudaMemsetAsync(gb, 0, 16 * 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (g_thin_tc) {
      int rc = tc::thin_tc_wgrad(0, img, g_bf16, pool_idx, B, H, W, gw, gb, st);
      if (rc != 1) return rc;
    }
    constexpr int RB = 256 / (5 * 4), IR = (RB - 1) + 5;
    const int IWp = (W - 1) + 5;
    const size_t smem = (size_t)((IR * IWp + 3) & ~3) * 4 + (size_t)RB * W * 16 * 2;
    LIVAE_CHECK_ARG(smem <= 200 * 1024, "thin_conv1c_wgrad: image too wide for the staged kernel");
    static OncePerDevice attr0;
    if (attr0.first()) { cudaFuncSetAttribute(conv1c_wgrad_kernel<16, 5, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);}
    int bands = (H + RB - 1) / RB;
    int by = (kNumSMs * 3 + bands - 1) / bands; if (by > B) by = B;
    conv1c_wgrad_kernel<16, 5, 1, true><<<dim3(bands, by), 256, smem, st>>>(img, g, pool_idx, B, H, W, H, W, 2, gw, gb);
  } else {
    if ((ce = cudaMemsetAsync(gw, 0, 32 * 16 * 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (gb && (ce = cudaMemsetAsync(gb, 0, 32 * 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (g_thin_tc) {
      int rc = tc::thin_tc_wgrad(1, img, g_bf16, nullptr, B, H, W, gw, gb, st);
      if (rc != 1) return rc;
    }
    constexpr int RB = 256 / (4 * 8), IR = (RB - 1) * 2 + 4;
    const int IWp = (W / 2 - 1) * 2 + 4;
    const size_t smem = (size_t)((IR * IWp + 3) & ~3) * 4 + (size_t)RB * (W / 2) * 32 * 2;
    LIVAE_CHECK_ARG(smem <= 200 * 1024, "thin_conv1c_wgrad: image too wide for the staged kernel");
    static OncePerDevice attr1;
    if (attr1.first()) { cudaFuncSetAttribute(conv1c_wgrad_kernel<32, 4, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);}
    int bands = (H / 2 + RB - 1) / RB;
    int by = (kNumSMs * 3 + bands - 1) / bands; if (by > B) by = B;
    conv1c_wgrad_kernel<32, 4, 2, false><<<dim3(bands, by), 256, smem, st>>>(img, g, nullptr, B, H, W, H / 2, W / 2, 1, gw, gb);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// encoder c1 data gradient: g bf16 [B,H/2,W/2,32] -> gimg fp32 [B,H,W]
extern "C" int livae_thin_conv1c_dgrad(const void* g_bf16, const float* w, int B, int H, int W, float* gimg,
                                       livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (H % 2) == 0 && (W % 2) == 0, "thin_conv1c_dgrad: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g_bf16 && w && gimg, "thin_conv1c_dgrad: null pointer");
  if (int e = require_sm100()) return e;
  if (g_thin_tc) {
    int rc = tc::thin_tc_col2im(1, g_bf16, w, nullptr, B, H / 2, W / 2, LIVAE_ACT_NONE, gimg, (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  conv1c_dgrad_kernel<32, 4, 2><<<tgrid((int64_t)B * H * W), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)g_bf16, w, B, H, W, H / 2, W / 2, 1, gimg);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// decoder d4 forward: x bf16 [B,H,W,32] (up-padded map) -> out fp32 [B,H-2,W-2], sigmoid
extern "C" int livae_thin_convc1_fwd(const void* x_bf16, const float* w, const float* bias, int B, int H, int W,
                                     int act, float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 2 && W > 2, "thin_convc1_fwd: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x_bf16 && w && out, "thin_convc1_fwd: null pointer");
  if (int e = require_sm100()) return e;
  if (g_thin_tc) {
    int rc = tc::thin_tc_col2im(0, x_bf16, w, bias, B, H, W, act, out, (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  convc1_fwd_kernel<32, 3><<<tgrid((int64_t)B * (H - 2) * (W - 2)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x_bf16, w, bias, B, H, W, H - 2, W - 2, act, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// decoder d4 weight/bias gradient: gw [1][32][3][3], gb [1] (written); g fp32 [B,H-2,W-2] pre-activation
extern "C" int livae_thin_convc1_wgrad(const void* x_bf16, const float* g, int B, int H, int W, float* gw, float* gb,
                                       livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 2 && W > 2, "thin_convc1_wgrad: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x_bf16 && g && gw, "thin_convc1_wgrad: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce;
  if ((ce = cudaMemsetAsync(gw, 0, 32 * 9 * 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
  if (gb && (ce = cudaMemsetAsync(gb, 0, 4, st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
  const int Ho = H - 2, Wo = W - 2;
  if (g_thin_tc) {
    int rc = tc::thin_tc_wgrad(2, g, x_bf16, nullptr, B, Ho, Wo, gw, gb, st);
    if (rc != 1) return rc;
  }
  constexpr int RWB = 4;
  const size_t smem = (size_t)(RWB + 2) * W * 32 * 2 + (size_t)RWB * Wo * 4;
  LIVAE_CHECK_ARG(smem <= 200 * 1024 && (W * 32) % 8 == 0, "thin_convc1_wgrad: map too wide for the staged kernel");
  static OncePerDevice attrd;
  if (attrd.first()) { cudaFuncSetAttribute(convc1_wgrad_kernel<32, 3, RWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);}
  int bands = (Ho + RWB - 1) / RWB;
  int by = (kNumSMs * 4 + bands - 1) / bands; if (by > B) by = B;
  convc1_wgrad_kernel<32, 3, RWB><<<dim3(bands, by), 9 * 16 * 2, smem, st>>>((const __nv_bfloat16*)x_bf16, g, B, H, W, Ho, Wo,
                                                                           gw, gb);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_sigmoid_bwd(const float* y, const float* g1, const float* g2, int64_t n, float* out,
                                 livae_stream_t stream) {
  LIVAE_CHECK_ARG(n >= 0, "sigmoid_bwd: bad size");
  if (n == 0) return 0;
  LIVAE_CHECK_ARG(y && g1 && out, "sigmoid_bwd: null pointer");
  if (int e = require_sm100()) return e;
  sigmoid_bwd_kernel<<<tgrid(n, 4), 256, 0, (cudaStream_t)stream>>>(y, g1, g2, n, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_maxpool_bf16(const void* full, int B, int H, int W, int C, void* pooled, uint8_t* idx,
                                  livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && (H % 2) == 0 && (W % 2) == 0, "maxpool_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(full && pooled && idx, "maxpool_bf16: null pointer");
  if (int e = require_sm100()) return e;
  if ((C & 7) == 0 && (((uintptr_t)full | (uintptr_t)pooled) & 15) == 0 && ((uintptr_t)idx & 7) == 0)
    maxpool_bf16v_kernel<<<tgrid((int64_t)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)full, B, H, W, C / 8, (uint4*)pooled, (uint2*)idx);
  else
    maxpool_bf16_kernel<<<tgrid((int64_t)B * (H / 2) * (W / 2) * C), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)full, B, H, W, C, (__nv_bfloat16*)pooled, idx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_unpool_bf16(const void* g_pooled, const uint8_t* idx, int B, int H, int W, int C, void* g_full,
                                 livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && (H % 2) == 0 && (W % 2) == 0, "unpool_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g_pooled && idx && g_full, "unpool_bf16: null pointer");
  if (int e = require_sm100()) return e;
  if ((C & 7) == 0 && (((uintptr_t)g_pooled | (uintptr_t)g_full) & 15) == 0 && ((uintptr_t)idx & 7) == 0)
    unpool_bf16v_kernel<<<tgrid((int64_t)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)g_pooled, (const uint2*)idx, B, H, W, C / 8, (uint4*)g_full);
  else
    unpool_bf16_kernel<<<tgrid((int64_t)B * H * W * C, 2), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)g_pooled, idx, B, H, W, C, (__nv_bfloat16*)g_full);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_relu_mask_cast_bf16(const float* g, const float* y, int64_t n, void* out_bf16,
                                         livae_stream_t stream) {
  LIVAE_CHECK_ARG(n >= 0, "relu_mask_cast_bf16: bad size");
  if (n == 0) return 0;
  LIVAE_CHECK_ARG(g && out_bf16, "relu_mask_cast_bf16: null pointer");
  if (int e = require_sm100()) return e;
  relu_mask_cast_kernel<<<tgrid(n, 4), 256, 0, (cudaStream_t)stream>>>(g, y, n, (__nv_bfloat16*)out_bf16);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// Memory-bound helpers around the GEMM layers:
//   * Upsample(x2, bilinear, align_corners=False) + ReflectionPad2d(1)      (model.py:357-370) fwd/adjoint
//   * Decoder.fc / VAEDecoder.fc: relu(Linear(L -> 256 q q)) viewed [B,256,q,q] (model.py:353,383-384; 84,108-110)
//   * global L2 norm + clip coefficient, fused AdamW on flat fp32 buffers    (train.py:396; scripts/train_rvae.py:157-159)
#include "common.cuh"

namespace livae {

// source index / lerp weight of nn.Upsample(scale 2, bilinear, align_corners=False):
// src = max(0.5*(dst+0.5)-0.5, 0); i0 = floor(src); i1 = min(i0+1, n-1); lambda = src - i0
__device__ __forceinline__ void up_src(int u, int n, int* i0, int* i1, float* f) {
  float s = fmaxf(0.5f * ((float)u + 0.5f) - 0.5f, 0.f);
  int a = (int)s;
  *i0 = a;
  *i1 = a < n - 1 ? a + 1 : a;
  *f = s - (float)a;
}
// ReflectionPad2d(1): padded index Y in [0, 2n+2) -> upsampled index in [0, 2n)
__device__ __forceinline__ int unpad(int Y, int n2) {
  int u = Y - 1;
  if (u < 0) u = -u;
  if (u >= n2) u = 2 * n2 - 2 - u;
  return u;
}

// x: [B,H,W,C] -> out: [B,2H+2,2W+2,C]
template <typename T>
__global__ void __launch_bounds__(256) upsample_pad_fwd_kernel(const T* __restrict__ x, int B, int H,
                                                               int W, int C, T* __restrict__ out) {
  int Ho = 2 * H + 2, Wo = 2 * W + 2;
  int64_t n = (int64_t)B * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t q = i / C; int X = (int)(q % Wo); q /= Wo; int Y = (int)(q % Ho);
    int b = (int)(q / Ho);
    int y0, y1, x0, x1; float fy, fx;
    up_src(unpad(Y, 2 * H), H, &y0, &y1, &fy);
    up_src(unpad(X, 2 * W), W, &x0, &x1, &fx);
    const T* s = x + (int64_t)b * H * W * C + c;
    float v00 = Cvt<T>::ld(s, ((int64_t)y0 * W + x0) * C), v01 = Cvt<T>::ld(s, ((int64_t)y0 * W + x1) * C);
    float v10 = Cvt<T>::ld(s, ((int64_t)y1 * W + x0) * C), v11 = Cvt<T>::ld(s, ((int64_t)y1 * W + x1) * C);
    // ATen: h0lambda*(w0lambda*v00 + w1lambda*v01) + h1lambda*(w0lambda*v10 + w1lambda*v11)
    Cvt<T>::st(out, i, (1.f - fy) * ((1.f - fx) * v00 + fx * v01) + fy * ((1.f - fx) * v10 + fx * v11));
  }
}

// 1-D adjoint weights: for source index i, sum over upsampled u in [2i-2, 2i+2] of
// w(u,i) * (sum of padded positions that map to u).  Done as a gather so no atomics are needed.
// gx[b,i,j,c] = relu_mask * sum_{u,v} wy(u,i) wx(v,j) G[u,v],  G[u,v] = sum_{Y in pre(u), X in pre(v)} g[Y,X]
template <typename T>
__global__ void __launch_bounds__(256) upsample_pad_bwd_kernel(const T* __restrict__ g, int B, int H,
                                                               int W, int C, const T* __restrict__ mask_y,
                                                               T* __restrict__ gx) {
  int Ho = 2 * H + 2, Wo = 2 * W + 2;
  int64_t n = (int64_t)B * H * W * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C); int64_t q = idx / C; int j = (int)(q % W); q /= W; int i = (int)(q % H);
    int b = (int)(q / H);
    if (mask_y && !(Cvt<T>::ld(mask_y, idx) > 0.f)) { Cvt<T>::st(gx, idx, 0.f); continue; }
    const T* gb = g + (int64_t)b * Ho * Wo * C + c;
    float acc = 0.f;
    for (int u = max(0, 2 * i - 2); u <= min(2 * H - 1, 2 * i + 2); ++u) {
      int y0, y1; float fy;
      up_src(u, H, &y0, &y1, &fy);
      float wy = (y0 == i ? 1.f - fy : 0.f) + (y1 == i ? fy : 0.f);
      if (wy == 0.f) continue;
      // padded rows that read upsampled row u: Y = u+1, plus the reflected border rows
      int Ys[2]; int ny = 0;
      Ys[ny++] = u + 1;
      if (u == 1) Ys[ny++] = 0;
      if (u == 2 * H - 2) Ys[ny++] = 2 * H + 1;
      for (int v = max(0, 2 * j - 2); v <= min(2 * W - 1, 2 * j + 2); ++v) {
        int x0, x1; float fx;
        up_src(v, W, &x0, &x1, &fx);
        float wx = (x0 == j ? 1.f - fx : 0.f) + (x1 == j ? fx : 0.f);
        if (wx == 0.f) continue;
        int Xs[2]; int nx = 0;
        Xs[nx++] = v + 1;
        if (v == 1) Xs[nx++] = 0;
        if (v == 2 * W - 2) Xs[nx++] = 2 * W + 1;
        float s = 0.f;
        for (int a = 0; a < ny; ++a)
          for (int d = 0; d < nx; ++d) s += Cvt<T>::ld(gb, ((int64_t)Ys[a] * Wo + Xs[d]) * C);
        acc += wy * wx * s;
      }
    }
    Cvt<T>::st(gx, idx, acc);
  }
}

// ---- bf16, 8 channels (one 16-byte vector) per thread: the index arithmetic is amortised over 8
// elements and every global access is a coalesced 128-bit transaction.
struct Up1D { int i0, i1; float f; };
__device__ __forceinline__ Up1D up1d(int Y, int n) {   // padded index -> source taps
  Up1D r;
  up_src(unpad(Y, 2 * n), n, &r.i0, &r.i1, &r.f);
  return r;
}
__device__ __forceinline__ void bf8_to_f(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 f_to_bf8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void __launch_bounds__(256) upsample_pad_fwd_bf16v_kernel(const uint4* __restrict__ x, int B, int H, int W,
                                                                     int C8, uint4* __restrict__ out) {
  const int Ho = 2 * H + 2, Wo = 2 * W + 2;
  const int64_t n = (int64_t)B * Ho * Wo * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8); int64_t q = i / C8; const int X = (int)(q % Wo); q /= Wo; const int Y = (int)(q % Ho);
    const int b = (int)(q / Ho);
    const Up1D ty = up1d(Y, H), tx = up1d(X, W);
    const uint4* s = x + (int64_t)b * H * W * C8 + c;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    bf8_to_f(__ldg(s + ((int64_t)ty.i0 * W + tx.i0) * C8), v00);
    bf8_to_f(__ldg(s + ((int64_t)ty.i0 * W + tx.i1) * C8), v01);
    bf8_to_f(__ldg(s + ((int64_t)ty.i1 * W + tx.i0) * C8), v10);
    bf8_to_f(__ldg(s + ((int64_t)ty.i1 * W + tx.i1) * C8), v11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ty.f) * ((1.f - tx.f) * v00[j] + tx.f * v01[j]) + ty.f * ((1.f - tx.f) * v10[j] + tx.f * v11[j]);
    out[i] = f_to_bf8(o);
  }
}

// 1-D adjoint taps of source index i: up to 4 upsampled rows with their lerp weights, each mapped
// to its padded row(s) (a reflected border row adds a second padded row).
struct Adj1D { int n; int Y[6]; float w[6]; };
__device__ __forceinline__ Adj1D adj1d(int i, int n) {
  Adj1D a; a.n = 0;
  const int lo = max(0, 2 * i - 1), hi = min(2 * n - 1, 2 * i + 2);
  for (int u = lo; u <= hi; ++u) {
    int i0, i1; float f;
    up_src(u, n, &i0, &i1, &f);
    const float wu = (i0 == i ? 1.f - f : 0.f) + (i1 == i ? f : 0.f);
    if (wu == 0.f) continue;
    a.Y[a.n] = u + 1; a.w[a.n] = wu; ++a.n;
    if (u == 1) { a.Y[a.n] = 0; a.w[a.n] = wu; ++a.n; }
    if (u == 2 * n - 2) { a.Y[a.n] = 2 * n + 1; a.w[a.n] = wu; ++a.n; }
  }
  return a;
}

__global__ void __launch_bounds__(256) upsample_pad_bwd_bf16v_kernel(const uint4* __restrict__ g, int B, int H, int W,
                                                                     int C8, const uint4* __restrict__ mask_y,
                                                                     uint4* __restrict__ gx) {
  const int Ho = 2 * H + 2, Wo = 2 * W + 2;
  const int64_t n = (int64_t)B * H * W * C8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8); int64_t q = idx / C8; const int j = (int)(q % W); q /= W; const int i = (int)(q % H);
    const int b = (int)(q / H);
    const Adj1D ay = adj1d(i, H), ax = adj1d(j, W);
    const uint4* gb = g + (int64_t)b * Ho * Wo * C8 + c;
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    for (int a = 0; a < ay.n; ++a) {
      float row[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) row[t] = 0.f;
      for (int d = 0; d < ax.n; ++d) {
        float v[8];
        bf8_to_f(__ldg(gb + ((int64_t)ay.Y[a] * Wo + ax.Y[d]) * C8), v);
#pragma unroll
        for (int t = 0; t < 8; ++t) row[t] = fmaf(ax.w[d], v[t], row[t]);
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(ay.w[a], row[t], acc[t]);
    }
    if (mask_y) {
      float m[8];
      bf8_to_f(__ldg(mask_y + idx), m);
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (!(m[t] > 0.f)) acc[t] = 0.f;
    }
    gx[idx] = f_to_bf8(acc);
  }
}

// ---- column-walk variants (C/8 a power of two).  The flat kernels above spend ~5 64-bit divisions per
// 16-byte vector and are instruction-bound at ~2 TB/s; one-row-per-CTA kernels were launch/latency-bound at
// 2-3 TB/s.  Here a thread owns one (column, 8-channel) slot and walks down a chunk of rows: all column taps
// and weights are computed once, vertical neighbours are carried in registers, and every step issues several
// independent 16-byte loads.
//
// Forward, "dual grid": block (r, s), r in [-1, H-1], s in [-1, W-1], is the 2x2 group of upsampled pixels
// (2r+1..2r+2, 2s+1..2s+2); all four interpolate between the same source rows {rA, rB} and columns {cA, cB}
// (up_src gives i0 = rA, i1 = rB for both), so a step loads one new source row (2 vectors) and stores 4
// (plus the reflected border copies).
static constexpr int kUpChunk = 16;      // block rows / source rows per CTA

// packed fp32 pairs (FFMA2 on sm_100): halves the instruction count of these conversion-heavy kernels
struct F8 { float2 v[4]; };
__device__ __forceinline__ F8 bf8_to_f2(const uint4& u) {
  F8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
  return r;
}
__device__ __forceinline__ uint4 f2_to_bf8(const F8& f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __float22bfloat162_rn(f.v[i]);
  return u;
}
// (1-f)*a + f*b with ATen's association: both products, then the sum (here: mul, then fma)
__device__ __forceinline__ F8 lerp8(const F8& a, const F8& b, float f) {
  F8 r;
  const float2 w0 = make_float2(1.f - f, 1.f - f), w1 = make_float2(f, f);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __ffma2_rn(w1, b.v[i], __fmul2_rn(w0, a.v[i]));
  return r;
}

__global__ void __launch_bounds__(288, 3) upsample_pad_fwd_walk_kernel(const uint4* __restrict__ x, int H, int W, int logc8,
                                                                       int rows_per_cta, uint4* __restrict__ out) {
  const int Ho = 2 * H + 2, Wo = 2 * W + 2, C8 = 1 << logc8;
  const int b = blockIdx.y;
  const int r_begin = (int)blockIdx.x * rows_per_cta - 1, r_end = min(H, r_begin + rows_per_cta);   // block rows [r_begin, r_end)
  const uint4* xb = x + (int64_t)b * H * W * C8;
  uint4* ob = out + (int64_t)b * Ho * Wo * C8;
  for (int t = threadIdx.x; t < (W + 1) * C8; t += blockDim.x) {
    const int s = (t >> logc8) - 1, c = t & (C8 - 1);
    const int cA = max(s, 0), cB = min(cA + 1, W - 1);
    // the two upsampled columns of this block, their lerp weights and padded positions
    const int v0 = 2 * s + 1, v1 = 2 * s + 2;
    const bool c0ok = v0 >= 0, c1ok = v1 < 2 * W;
    float fx0 = 0.f, fx1 = 0.f;
    { int i0, i1; if (c0ok) up_src(v0, W, &i0, &i1, &fx0); if (c1ok) up_src(v1, W, &i0, &i1, &fx1); }
    const int X0 = v0 + 1, X1 = v1 + 1;
    const int X0dup = v0 == 1 ? 0 : (v0 == 2 * W - 2 ? 2 * W + 1 : -1);
    const int X1dup = v1 == 1 ? 0 : (v1 == 2 * W - 2 ? 2 * W + 1 : -1);
    const uint4* xa = xb + (int64_t)cA * C8 + c;
    const uint4* xq = xb + (int64_t)cB * C8 + c;
    uint4* oc = ob + c;
    // 32-bit offsets inside one image (index arithmetic was 80% of this kernel's instructions)
    const int rowx = W * C8, rowo = Wo * C8;
    const int X0c = X0 * C8, X1c = X1 * C8, X0dc = X0dup * C8, X1dc = X1dup * C8;
    // horizontal lerps of the two source rows of the block (ATen's inner parentheses), carried down the walk
    F8 lo0, lo1, hi0, hi1;
    int prev_rB = -1;
#pragma unroll 1
    for (int r = r_begin; r < r_end; ++r) {
      const int rA = max(r, 0), rB = min(rA + 1, H - 1);
      if (rA == prev_rB) { lo0 = hi0; lo1 = hi1; }
      else {
        const F8 a = bf8_to_f2(__ldg(xa + rA * rowx)), q = bf8_to_f2(__ldg(xq + rA * rowx));
        lo0 = lerp8(a, q, fx0); lo1 = lerp8(a, q, fx1);
      }
      if (rB != rA) {
        const F8 a = bf8_to_f2(__ldg(xa + rB * rowx)), q = bf8_to_f2(__ldg(xq + rB * rowx));
        hi0 = lerp8(a, q, fx0); hi1 = lerp8(a, q, fx1);
      } else { hi0 = lo0; hi1 = lo1; }
      prev_rB = rB;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int u = 2 * r + 1 + k;
        if (u < 0 || u >= 2 * H) continue;
        // up_src(u): rows 2r+1 / 2r+2 sit at r + 0.25 / r + 0.75 (row 0 is clamped to the first source row)
        const float fy = k == 0 ? 0.25f : (r < 0 ? 0.f : 0.75f);
        const int offY = (u + 1) * rowo;
        const int offD = u == 1 ? 0 : (u == 2 * H - 2 ? (2 * H + 1) * rowo : -1);
        if (c0ok) {
          const uint4 q = f2_to_bf8(lerp8(lo0, hi0, fy));
          oc[offY + X0c] = q;
          if (X0dup >= 0) oc[offY + X0dc] = q;
          if (offD >= 0) {
            oc[offD + X0c] = q;
            if (X0dup >= 0) oc[offD + X0dc] = q;
          }
        }
        if (c1ok) {
          const uint4 q = f2_to_bf8(lerp8(lo1, hi1, fy));
          oc[offY + X1c] = q;
          if (X1dup >= 0) oc[offY + X1dc] = q;
          if (offD >= 0) {
            oc[offD + X1c] = q;
            if (X1dup >= 0) oc[offD + X1dc] = q;
          }
        }
      }
    }
  }
}

// weight of upsampled row u on source row i (0 outside [0, 2n))
__device__ __forceinline__ float up_w(int u, int i, int n) {
  if (u < 0 || u >= 2 * n) return 0.f;
  int i0, i1; float f;
  up_src(u, n, &i0, &i1, &f);
  return (i0 == i ? 1.f - f : 0.f) + (i1 == i ? f : 0.f);
}
// Adjoint, separable with register carry: E[u] = horizontal adjoint of the gradient row(s) that read upsampled
// row u (padded row u+1, plus the reflected copy for u = 1 and u = 2H-2); source row i needs E[2i-1..2i+2], of
// which the first two were the "new" rows of step i-1.  Thread = (source column j, 8 channels).  The column
// taps of j are the padded columns 2j..2j+3 plus, for j = 1 / j = W-2, the reflected border column; they live
// in registers (static indices only -- a dynamically built tap list went to local memory).
struct UpTaps { int X[6]; float w[6]; };
__device__ __forceinline__ UpTaps up_col_taps(int j, int W) {
  UpTaps t;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int X = 2 * j + k;
    t.X[k] = X;
    t.w[k] = up_w(unpad(X, 2 * W), j, W);      // X = 0 and X = 2W+1 are copies of upsampled columns 1 and 2W-2
  }
  t.X[4] = 0;         t.w[4] = j == 1 ? up_w(1, j, W) : 0.f;
  t.X[5] = 2 * W + 1; t.w[5] = j == W - 2 ? up_w(2 * W - 2, j, W) : 0.f;
  return t;
}
// the same with X premultiplied by the vectors per pixel (C8): 32-bit offsets inside a gradient row
__device__ __forceinline__ UpTaps up_col_taps_scaled(int j, int W, int C8) {
  UpTaps t = up_col_taps(j, W);
#pragma unroll
  for (int k = 0; k < 6; ++k) t.X[k] *= C8;
  return t;
}
__device__ __forceinline__ void up_adj_hrow(const uint4* __restrict__ grow, const UpTaps& ax, F8& h) {
  uint4 q[6];
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (ax.w[k] != 0.f) q[k] = __ldg(grow + ax.X[k]);
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (ax.w[k] != 0.f) {
      const F8 v = bf8_to_f2(q[k]);
      const float2 w = make_float2(ax.w[k], ax.w[k]);
#pragma unroll
      for (int e = 0; e < 4; ++e) h.v[e] = __ffma2_rn(w, v.v[e], h.v[e]);
    }
}
// gb (optional, fp32 [C], zeroed by the caller): column sums of the values written to gx, i.e. the bias gradient
// of the convolution whose pre-activation gradient gx is -- fused here because a separate pass re-read gx.
__global__ void __launch_bounds__(256, 3) upsample_pad_bwd_walk_kernel(const uint4* __restrict__ g, int H, int W, int logc8,
                                                                    int rows_per_cta, const uint4* __restrict__ mask_y,
                                                                    uint4* __restrict__ gx, float* __restrict__ gb) {
  __shared__ float s_colsum[256 * 8];
  const int Ho = 2 * H + 2, Wo = 2 * W + 2, C8 = 1 << logc8;
  const int b = blockIdx.y;
  const int i_begin = (int)blockIdx.x * rows_per_cta, i_end = min(H, i_begin + rows_per_cta);
  for (int t = threadIdx.x; t < W * C8; t += blockDim.x) {
    const int j = t >> logc8, c = t & (C8 - 1);
    const UpTaps ax = up_col_taps_scaled(j, W, C8);
    const int rowg = Wo * C8;
    const uint4* gcol = g + (int64_t)b * Ho * Wo * C8 + c;    // (row Y, column X) at gcol[(Y*Wo + X)*C8]
    // walk the upsampled rows u that touch source rows [i_begin, i_end): row u adds (1-f) E[u] to source row
    // i0(u) and f E[u] to i1(u); acc0 / acc1 are the running sums of source rows a and a+1
    F8 acc0, acc1, csum;
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc0.v[e] = make_float2(0.f, 0.f); acc1.v[e] = make_float2(0.f, 0.f); csum.v[e] = make_float2(0.f, 0.f); }
    int a = i_begin - 1;
    auto emit = [&]() {
      if (a < i_begin || a >= i_end) return;
      const int64_t o = (((int64_t)b * H + a) * W + j) * C8 + c;
      if (mask_y) {
        const F8 m = bf8_to_f2(__ldg(mask_y + o));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!(m.v[e].x > 0.f)) acc0.v[e].x = 0.f;
          if (!(m.v[e].y > 0.f)) acc0.v[e].y = 0.f;
        }
      }
      gx[o] = f2_to_bf8(acc0);
#pragma unroll
      for (int e = 0; e < 4; ++e) csum.v[e] = __fadd2_rn(csum.v[e], acc0.v[e]);
    };
    const int u_lo = max(0, 2 * i_begin - 1), u_hi = min(2 * H - 1, 2 * i_end);
#pragma unroll 1
    for (int u = u_lo; u <= u_hi; ++u) {
      // up_src(u): row u sits at (u-1)/2 + 0.25 / 0.75; row 0 is clamped to source row 0
      const int i0 = u == 0 ? 0 : (u - 1) >> 1, i1 = min(i0 + 1, H - 1);
      const float f = u == 0 ? 0.f : ((u & 1) ? 0.25f : 0.75f);
      if (i0 > a) {
        emit();
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc0.v[e] = acc1.v[e]; acc1.v[e] = make_float2(0.f, 0.f); }
        ++a;
      }
      F8 E;
#pragma unroll
      for (int e = 0; e < 4; ++e) E.v[e] = make_float2(0.f, 0.f);
      const int Ydup = u == 1 ? 0 : (u == 2 * H - 2 ? 2 * H + 1 : -1);
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const int Y = rep ? Ydup : u + 1;
        if (Y < 0) break;
        up_adj_hrow(gcol + Y * rowg, ax, E);
      }
      const float w0 = i1 == i0 ? 1.f : 1.f - f, w1 = i1 == i0 ? 0.f : f;
      const float2 w02 = make_float2(w0, w0), w12 = make_float2(w1, w1);
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0.v[e] = __ffma2_rn(w02, E.v[e], acc0.v[e]); acc1.v[e] = __ffma2_rn(w12, E.v[e], acc1.v[e]); }
    }
    emit();
    if (gb) {      // host guarantees one pass of the t loop (W * C8 <= blockDim.x) when gb is given
#pragma unroll
      for (int e = 0; e < 4; ++e) { s_colsum[t * 8 + 2 * e] = csum.v[e].x; s_colsum[t * 8 + 2 * e + 1] = csum.v[e].y; }
    }
  }
  if (gb) {
    __syncthreads();
    // threads t, t + C8, t + 2*C8, ... own the same 8 channels: one thread per channel sums them
    const int C = C8 * 8, nt = W * C8;
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const int c = ch >> 3, e = ch & 7;
      float sum = 0.f;
      for (int t = c; t < nt; t += C8) sum += s_colsum[t * 8 + e];
      atomicAdd(gb + ch, sum);
    }
  }
}

static inline int log2_exact(int v) {   // -1 when v is not a power of two
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

// out[b, (h,w,c)] = relu(sum_l z[b,l] * w[(c,h,w), l] + bias[(c,h,w)])  -- NHWC output of the
// reference's h.view(B, 256, q, q)
template <typename T>
__global__ void __launch_bounds__(256) decfc_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int B, int L, int C,
                                                        int HW, T* __restrict__ out) {
  int64_t n = (int64_t)B * HW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t q = i / C; int p = (int)(q % HW); int b = (int)(q / HW);
    int row = c * HW + p;
    float a = bias[row];
    for (int l = 0; l < L; ++l) a = fmaf(z[b * L + l], w[(int64_t)row * L + l], a);
    Cvt<T>::st(out, i, fmaxf(a, 0.f));
  }
}

// Weight-stationary bf16 variant: a thread owns 8 consecutive channels of one pixel (its 8 x L weights and
// biases live in registers), walks a slice of the batch and writes one 16-byte vector per sample --
// coalesced stores, no per-element index arithmetic, weights read once.  L <= 8.
// the same with 4 channels (one 8-byte vector) per thread so that 16 latent columns of weights fit in registers
__global__ void __launch_bounds__(256) decfc_fwd_ws4_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int B, int L, int C4, int HW,
                                                            uint2* __restrict__ out) {
  constexpr int LMAX = 16;
  const int v = blockIdx.x * 256 + threadIdx.x;         // vector index inside one sample: p * C4 + c4
  const int nv = HW * C4;
  if (v >= nv) return;
  const int p = v / C4, c0 = (v - p * C4) * 4;
  float wr[4][LMAX], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = (c0 + j) * HW + p;
    br[j] = bias[row];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) wr[j][l] = l < L ? w[(int64_t)row * L + l] : 0.f;
  }
  const int bchunk = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    float zl[LMAX];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) zl[l] = l < L ? __ldg(z + b * L + l) : 0.f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = br[j];
#pragma unroll
      for (int l = 0; l < LMAX; ++l) a = fmaf(zl[l], wr[j][l], a);
      o[j] = fmaxf(a, 0.f);
    }
    uint2 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
    h[0] = __floats2bfloat162_rn(o[0], o[1]); h[1] = __floats2bfloat162_rn(o[2], o[3]);
    out[(int64_t)b * nv + v] = q;
  }
}

template <int LMAX>
__global__ void __launch_bounds__(256) decfc_fwd_ws_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int B, int L, int C8, int HW,
                                                           uint4* __restrict__ out) {
  const int v = blockIdx.x * 256 + threadIdx.x;         // vector index inside one sample: p * C8 + c8
  const int nv = HW * C8;
  if (v >= nv) return;
  const int p = v / C8, c0 = (v - p * C8) * 8;
  float wr[8][LMAX], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int row = (c0 + j) * HW + p;
    br[j] = bias[row];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) wr[j][l] = l < L ? w[(int64_t)row * L + l] : 0.f;
  }
  const int bchunk = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    float zl[LMAX];
#pragma unroll
    for (int l = 0; l < LMAX; ++l) zl[l] = l < L ? __ldg(z + b * L + l) : 0.f;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = br[j];
#pragma unroll
      for (int l = 0; l < LMAX; ++l) a = fmaf(zl[l], wr[j][l], a);
      o[j] = fmaxf(a, 0.f);
    }
    out[(int64_t)b * nv + v] = f_to_bf8(o);
  }
}

// gz[b,l] = sum_n gpre[b,n] w[n,l]; one CTA per sample
// y == nullptr: gy is already the pre-activation gradient (tensor-core path convention)
template <typename T, int LT = 4>
__global__ void __launch_bounds__(256) decfc_bwd_z_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                          const float* __restrict__ w, int L, int C, int HW,
                                                          float* __restrict__ gz) {
  __shared__ float red[32];
  int b = blockIdx.x;
  int N = C * HW;
  const T* gyb = gy + (int64_t)b * N;
  const T* yb = y ? y + (int64_t)b * N : nullptr;
  for (int l0 = 0; l0 < L; l0 += LT) {          // LT latent columns per pass over gy (one pass when L <= LT)
    float a[LT];
#pragma unroll
    for (int t = 0; t < LT; ++t) a[t] = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      float g = Cvt<T>::ld(gyb, i);
      if (yb && !(Cvt<T>::ld(yb, i) > 0.f)) g = 0.f;
      int c = i % C, p = i / C;
      const float* wr = w + (int64_t)(c * HW + p) * L;
#pragma unroll
      for (int t = 0; t < LT; ++t)
        if (l0 + t < L) a[t] = fmaf(g, wr[l0 + t], a[t]);
    }
#pragma unroll
    for (int t = 0; t < LT; ++t) {
      float s = block_sum(a[t], red);
      if (threadIdx.x == 0 && l0 + t < L) gz[b * L + l0 + t] = s;
      __syncthreads();
    }
  }
}

// gw[n,l] = sum_b gpre[b,n] z[b,l], gb[n] = sum_b gpre[b,n]; thread per output feature n (NHWC
// order so that a warp reads contiguous gy), loop over the batch.  Batch is split over
// blockIdx.y and combined with atomics on the (zeroed) outputs.
template <int LT, typename T>
__global__ void __launch_bounds__(128) decfc_bwd_w_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                          const float* __restrict__ z, int B, int L, int C, int HW,
                                                          float* __restrict__ gw, float* __restrict__ gb) {
  int N = C * HW;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c = i % C, p = i / C;
  int row = c * HW + p;
  int bchunk = (B + gridDim.y - 1) / gridDim.y;
  int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  for (int l0 = 0; l0 < L; l0 += LT) {
    float a[LT];
#pragma unroll
    for (int t = 0; t < LT; ++t) a[t] = 0.f;
    float sb = 0.f;
    for (int b = b0; b < b1; ++b) {
      float g = Cvt<T>::ld(gy, (int64_t)b * N + i);
      if (y && !(Cvt<T>::ld(y, (int64_t)b * N + i) > 0.f)) g = 0.f;
      sb += g;
#pragma unroll
      for (int t = 0; t < LT; ++t)
        if (l0 + t < L) a[t] = fmaf(g, __ldg(z + b * L + l0 + t), a[t]);
    }
#pragma unroll
    for (int t = 0; t < LT; ++t)
      if (l0 + t < L) atomicAdd(gw + (int64_t)row * L + l0 + t, a[t]);
    if (l0 == 0 && gb) atomicAdd(gb + row, sb);
  }
}

// Decoder.fc backward for small latent sizes in ONE pass over gy (bf16 [B, HW, C], already the pre-activation
// gradient): gw[n,l] = sum_b g[b,n] z[b,l], gb[n] = sum_b g[b,n], gz[b,l] = sum_n g[b,n] w[n,l].  Thread = 8
// consecutive channels of one pixel (one 16-byte load per batch row), CTA = 2048 features x a slice of the batch;
// its weight rows live in registers.  gz partials: warp shuffle -> shared atomics -> one global atomic per
// (row, l, CTA).  All three outputs are zeroed by the caller.  The two-kernel form above read gy twice with
// scalar loads (0.33 ms at B = 2048, N = 16384 against 67 MB of traffic).
template <int LT>
__global__ void __launch_bounds__(256) decfc_bwd_fused_bf16_kernel(const uint4* __restrict__ gy,
                                                                   const float* __restrict__ z,
                                                                   const float* __restrict__ w, int B, int L, int C,
                                                                   int HW, int rows_per_cta, float* __restrict__ gw,
                                                                   float* __restrict__ gb, float* __restrict__ gz) {
  extern __shared__ float sgz[];                       // [rows_per_cta][LT]
  const int nvec = C * HW / 8;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = v < nvec;
  const int b0 = blockIdx.y * rows_per_cta, b1 = min(B, b0 + rows_per_cta);
  for (int i = threadIdx.x; i < rows_per_cta * LT; i += blockDim.x) sgz[i] = 0.f;
  __syncthreads();
  const int c0 = ok ? (v * 8) % C : 0, p = ok ? (v * 8) / C : 0;
  float wr[8][LT], aw[8][LT], ab[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    ab[e] = 0.f;
#pragma unroll
    for (int l = 0; l < LT; ++l) {
      aw[e][l] = 0.f;
      wr[e][l] = (ok && l < L) ? w[(int64_t)((c0 + e) * HW + p) * L + l] : 0.f;
    }
  }
  const int lane = threadIdx.x & 31;
  for (int b = b0; b < b1; ++b) {
    float g[8];
    if (ok) bf8_to_f(__ldg(gy + (int64_t)b * nvec + v), g);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = 0.f;
    }
    float zl[LT], pz[LT];
#pragma unroll
    for (int l = 0; l < LT; ++l) { zl[l] = l < L ? __ldg(z + b * L + l) : 0.f; pz[l] = 0.f; }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ab[e] += g[e];
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        aw[e][l] = fmaf(g[e], zl[l], aw[e][l]);
        pz[l] = fmaf(g[e], wr[e][l], pz[l]);
      }
    }
    if (gz) {
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        const float t = warp_sum(pz[l]);
        if (lane == 0 && l < L) atomicAdd(&sgz[(b - b0) * LT + l], t);
      }
    }
  }
  if (ok) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t row = (int64_t)(c0 + e) * HW + p;
      if (gb) atomicAdd(gb + row, ab[e]);
#pragma unroll
      for (int l = 0; l < LT; ++l)
        if (l < L) atomicAdd(gw + row * L + l, aw[e][l]);
    }
  }
  if (gz) {
    __syncthreads();
    for (int i = threadIdx.x; i < (b1 - b0) * LT; i += blockDim.x) {
      const int l = i % LT;
      if (l < L) atomicAdd(gz + (int64_t)(b0 + i / LT) * L + l, sgz[i]);
    }
  }
}

// ---- optimiser side ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n,
                                                    float* __restrict__ partial) {
  __shared__ float red[32];
  float a = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i];
    a = fmaf(v, v, a);
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = a;
}

// out[0] = norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))  (torch clip_grad_norm_)
__global__ void __launch_bounds__(256) norm_finish_kernel(const float* __restrict__ partial, int nparts,
                                                          float max_norm, float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += partial[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) {
    float nrm = sqrtf(a);
    out[0] = nrm;
    out[1] = fminf(1.f, max_norm / (nrm + 1e-6f));
  }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ g, int64_t n,
                                                    const float* __restrict__ coef) {
  float c = coef[0];
  if (c == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    g[i] *= c;
}

// torch.optim.AdamW / Adam (decoupled = 1 / 0), single flat tensor.  step_dev holds the step
// count as a float and is incremented by thread 0 AFTER all reads (separate tiny kernel order).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float lr, float b1, float b2, float eps, float wd,
                                                    int decoupled, const float* __restrict__ step_dev,
                                                    const float* __restrict__ gscale) {
  float t = step_dev[0] + 1.f;
  float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  float gs = gscale ? gscale[0] : 1.f;
  float step_size = lr / bc1;
  float bc2s = sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], gi = g[i] * gs;
    if (decoupled) pi *= (1.f - lr * wd);
    else gi = fmaf(wd, pi, gi);
    float mi = m[i] + (1.f - b1) * (gi - m[i]);       // lerp, as torch does
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) / bc2s + eps;
    p[i] = pi - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}
__global__ void inc_step_kernel(float* step_dev) { step_dev[0] += 1.f; }

static inline int sgrid(int64_t n, int per = 4) {
  int64_t blocks = (n + 256LL * per - 1) / (256LL * per);
  if (blocks < 1) blocks = 1;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace livae

using namespace livae;

extern "C" int livae_upsample_pad_fwd(const float* x, int B, int H, int W, int C, float* out,
                                      livae_stream_t stream) {
  LIVAE_CHECK_ARG(x && out && B >= 0 && H > 1 && W > 1 && C > 0, "upsample_pad_fwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  int64_t n = (int64_t)B * (2 * H + 2) * (2 * W + 2) * C;
  upsample_pad_fwd_kernel<float><<<sgrid(n, 2), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_upsample_pad_bwd(const float* g, int B, int H, int W, int C, const float* relu_mask_y,
                                      float* gx, livae_stream_t stream) {
  LIVAE_CHECK_ARG(g && gx && B >= 0 && H > 1 && W > 1 && C > 0, "upsample_pad_bwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  int64_t n = (int64_t)B * H * W * C;
  upsample_pad_bwd_kernel<float><<<sgrid(n, 1), 256, 0, (cudaStream_t)stream>>>(g, B, H, W, C, relu_mask_y, gx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_decfc_fwd(const float* z, const float* w, const float* bias, int B, int L, int C, int HW,
                               float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(z && w && bias && out && B >= 0 && L > 0 && C > 0 && HW > 0, "decfc_fwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  decfc_fwd_kernel<float><<<sgrid((int64_t)B * C * HW, 2), 256, 0, (cudaStream_t)stream>>>(z, w, bias, B, L, C, HW, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_decfc_bwd(const float* z, const float* w, const float* y, const float* gy, int B, int L,
                               int C, int HW, float* gw, float* gb, float* gz, livae_stream_t stream) {
  LIVAE_CHECK_ARG(z && w && y && gy && B >= 0 && L > 0 && C > 0 && HW > 0, "decfc_bwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int N = C * HW;
  if (gz) {
    decfc_bwd_z_kernel<float><<<B, 256, 0, st>>>(gy, y, w, L, C, HW, gz);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  if (gw) {
    cudaError_t ce;
    if ((ce = cudaMemsetAsync(gw, 0, (size_t)N * L * sizeof(float), st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (gb && (ce = cudaMemsetAsync(gb, 0, (size_t)N * sizeof(float), st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    int nb = (N + 127) / 128;
    int ysplit = (kNumSMs * 8 + nb - 1) / nb;
    if (ysplit > (B + 15) / 16) ysplit = (B + 15) / 16;
    if (ysplit < 1) ysplit = 1;
    dim3 grid(nb, ysplit);
    decfc_bwd_w_kernel<4, float><<<grid, 128, 0, st>>>(gy, y, z, B, L, C, HW, gw, gb);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  return 0;
}

// ---- bf16 variants used by the tensor-core path (gradients are pre-activation: no ReLU re-masking) ----
extern "C" int livae_upsample_pad_fwd_bf16(const void* x, int B, int H, int W, int C, void* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 1 && W > 1 && C > 0, "upsample_pad_fwd_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x && out, "upsample_pad_fwd_bf16: null pointer");
  if (int e = require_sm100()) return e;
  int64_t n = (int64_t)B * (2 * H + 2) * (2 * W + 2) * C;
  const int logc8 = (C & 7) == 0 ? log2_exact(C / 8) : -1;
  if (logc8 >= 0 && B <= 65535 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0)
  {
    const int nchunk = (H + 1 + kUpChunk - 1) / kUpChunk, rows = (H + 1 + nchunk - 1) / nchunk;
    int threads = (((W + 1) << logc8) + 31) & ~31;
    if (threads > 288) threads = 288;
    upsample_pad_fwd_walk_kernel<<<dim3(nchunk, B), threads, 0, (cudaStream_t)stream>>>((const uint4*)x, H, W, logc8, rows, (uint4*)out);
  }
  else if ((C & 7) == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0)
    upsample_pad_fwd_bf16v_kernel<<<sgrid(n / 8, 1), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, B, H, W, C / 8,
                                                                                   (uint4*)out);
  else
    upsample_pad_fwd_kernel<__nv_bfloat16><<<sgrid(n, 2), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, B, H, W, C, (__nv_bfloat16*)out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

static int upsample_pad_bwd_bf16_impl(const void* g, int B, int H, int W, int C, const void* relu_mask_y, void* gx,
                                      float* gb, livae_stream_t stream);
extern "C" int livae_upsample_pad_bwd_bf16(const void* g, int B, int H, int W, int C, const void* relu_mask_y,
                                           void* gx, livae_stream_t stream) {
  return upsample_pad_bwd_bf16_impl(g, B, H, W, C, relu_mask_y, gx, nullptr, stream);
}
// the same, plus gb[C] (fp32, written) = sum over (b, i, j) of the gx values: the bias gradient of the
// convolution that produced the up-sampled tensor's source (reference: autograd of model.py:359-367)
extern "C" int livae_upsample_pad_bwd_bias_bf16(const void* g, int B, int H, int W, int C, const void* relu_mask_y,
                                                void* gx, float* gb, livae_stream_t stream) {
  LIVAE_CHECK_ARG(gb, "upsample_pad_bwd_bias_bf16: null gb");
  if (B == 0) return 0;
  if (int e = require_sm100()) return e;
  cudaError_t ce = cudaMemsetAsync(gb, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
  if (ce != cudaSuccess) { set_error("upsample_pad_bwd_bias_bf16 memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  return upsample_pad_bwd_bf16_impl(g, B, H, W, C, relu_mask_y, gx, gb, stream);
}
static int upsample_pad_bwd_bf16_impl(const void* g, int B, int H, int W, int C, const void* relu_mask_y, void* gx,
                                      float* gb, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 1 && W > 1 && C > 0, "upsample_pad_bwd_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g && gx, "upsample_pad_bwd_bf16: null pointer");
  if (int e = require_sm100()) return e;
  int64_t n = (int64_t)B * H * W * C;
  const int logc8 = (C & 7) == 0 ? log2_exact(C / 8) : -1;
  const bool walk_ok = logc8 >= 0 && B <= 65535 && (((uintptr_t)g | (uintptr_t)gx | (uintptr_t)relu_mask_y) & 15) == 0;
  LIVAE_CHECK_ARG(!gb || (walk_ok && (W << logc8) <= 256),
                  "upsample_pad_bwd_bias_bf16: needs C/8 a power of two, W*C/8 <= 256 and 16-byte aligned pointers");
  if (walk_ok) {
    const int nchunk = (H + kUpChunk - 1) / kUpChunk, rows = (H + nchunk - 1) / nchunk;
    int threads = ((W << logc8) + 31) & ~31;
    if (threads > 256) threads = 256;
    upsample_pad_bwd_walk_kernel<<<dim3(nchunk, B), threads, 0, (cudaStream_t)stream>>>((const uint4*)g, H, W, logc8, rows,
                                                                                       (const uint4*)relu_mask_y, (uint4*)gx, gb);
  }
  else if ((C & 7) == 0 && (((uintptr_t)g | (uintptr_t)gx | (uintptr_t)relu_mask_y) & 15) == 0)
    upsample_pad_bwd_bf16v_kernel<<<sgrid(n / 8, 1), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)g, B, H, W, C / 8, (const uint4*)relu_mask_y, (uint4*)gx);
  else
    upsample_pad_bwd_kernel<__nv_bfloat16><<<sgrid(n, 1), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)g, B, H, W, C, (const __nv_bfloat16*)relu_mask_y, (__nv_bfloat16*)gx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_decfc_fwd_bf16(const float* z, const float* w, const float* bias, int B, int L, int C, int HW,
                                    void* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && L > 0 && C > 0 && HW > 0, "decfc_fwd_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(z && w && bias && out, "decfc_fwd_bf16: null pointer");
  if (int e = require_sm100()) return e;
  if (L <= 4 && (C & 7) == 0 && ((uintptr_t)out & 15) == 0) {
    const int nv = HW * (C / 8);
    const int gx = (nv + 255) / 256;
    int gy = (kNumSMs * 8 + gx - 1) / gx;
    if (gy > B) gy = B;
    decfc_fwd_ws_kernel<4><<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(z, w, bias, B, L, C / 8, HW, (uint4*)out);
  } else if (L <= 16 && (C & 3) == 0 && ((uintptr_t)out & 7) == 0) {
    const int nv = HW * (C / 4);
    const int gx = (nv + 255) / 256;
    int gy = (kNumSMs * 8 + gx - 1) / gx;
    if (gy > B) gy = B;
    decfc_fwd_ws4_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(z, w, bias, B, L, C / 4, HW, (uint2*)out);
  } else {
    decfc_fwd_kernel<__nv_bfloat16><<<sgrid((int64_t)B * C * HW, 2), 256, 0, (cudaStream_t)stream>>>(
        z, w, bias, B, L, C, HW, (__nv_bfloat16*)out);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// gy: bf16 PRE-activation gradient w.r.t. the fc output (already masked by the producer)
extern "C" int livae_decfc_bwd_bf16(const float* z, const float* w, const void* gy, int B, int L, int C, int HW,
                                    float* gw, float* gb, float* gz, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && L > 0 && C > 0 && HW > 0, "decfc_bwd_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(z && w && gy, "decfc_bwd_bf16: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* g = (const __nv_bfloat16*)gy;
  int N = C * HW;
  if (L <= 4 && gw && C % 8 == 0 && (((uintptr_t)gy) & 15) == 0) {
    cudaError_t ce;
    if ((ce = cudaMemsetAsync(gw, 0, (size_t)N * L * sizeof(float), st)) != cudaSuccess ||
        (gb && (ce = cudaMemsetAsync(gb, 0, (size_t)N * sizeof(float), st)) != cudaSuccess) ||
        (gz && (ce = cudaMemsetAsync(gz, 0, (size_t)B * L * sizeof(float), st)) != cudaSuccess)) {
      set_error("decfc_bwd_bf16: memset failed");
      return (int)ce;
    }
    const int nb = (N / 8 + 255) / 256;
    int rows = (B + 31) / 32;                     // ~32 batch slices: 8 x 32 CTAs at N = 16384
    if (rows < 16) rows = 16;
    if (rows > 256) rows = 256;
    decfc_bwd_fused_bf16_kernel<4><<<dim3(nb, (B + rows - 1) / rows), 256, rows * 4 * sizeof(float), st>>>(
        (const uint4*)gy, z, w, B, L, C, HW, rows, gw, gb, gz);
    LIVAE_CUDA_LAUNCH_CHECK();
    return 0;
  }
  if (gz) {
    if (L <= 4) decfc_bwd_z_kernel<__nv_bfloat16, 4><<<B, 256, 0, st>>>(g, (const __nv_bfloat16*)nullptr, w, L, C, HW, gz);
    else decfc_bwd_z_kernel<__nv_bfloat16, 16><<<B, 256, 0, st>>>(g, (const __nv_bfloat16*)nullptr, w, L, C, HW, gz);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  if (gw) {
    cudaError_t ce;
    if ((ce = cudaMemsetAsync(gw, 0, (size_t)N * L * sizeof(float), st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    if (gb && (ce = cudaMemsetAsync(gb, 0, (size_t)N * sizeof(float), st)) != cudaSuccess) { set_error("memset"); return (int)ce; }
    int nb = (N + 127) / 128;
    int ysplit = (kNumSMs * 8 + nb - 1) / nb;
    if (ysplit > (B + 15) / 16) ysplit = (B + 15) / 16;
    if (ysplit < 1) ysplit = 1;
    if (L <= 4)
      decfc_bwd_w_kernel<4, __nv_bfloat16><<<dim3(nb, ysplit), 128, 0, st>>>(g, (const __nv_bfloat16*)nullptr, z, B, L, C, HW, gw, gb);
    else      // 16 latent columns per pass over gy
      decfc_bwd_w_kernel<16, __nv_bfloat16><<<dim3(nb, ysplit), 128, 0, st>>>(g, (const __nv_bfloat16*)nullptr, z, B, L, C, HW, gw, gb);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  return 0;
}

// Linear weight gradient from the tensor-core kernel's [Npad][(h,w,c)] layout to torch's
// [N][(c,h,w)] (model.py:210, 321: NCHW flatten), dropping padded rows.
__global__ void permute_linear_grad_kernel(const float* __restrict__ src, int N, int C, int HW, float* __restrict__ dst) {
  int64_t n = (int64_t)N * C * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int p = (int)(i % HW); int64_t q = i / HW; int c = (int)(q % C); int r = (int)(q / C);
    dst[i] = src[((int64_t)r * HW + p) * C + c];
  }
}
extern "C" int livae_permute_linear_grad(const float* src_hwc, int N, int C, int HW, float* dst_chw, livae_stream_t stream) {
  LIVAE_CHECK_ARG(src_hwc && dst_chw && N > 0 && C > 0 && HW > 0, "permute_linear_grad: bad args");
  if (int e = require_sm100()) return e;
  permute_linear_grad_kernel<<<sgrid((int64_t)N * C * HW, 2), 256, 0, (cudaStream_t)stream>>>(src_hwc, N, C, HW, dst_chw);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t livae_l2norm_scratch_floats(void) { return kNumSMs * 16 + 8; }

extern "C" int livae_l2norm_clip(float* grads, int64_t n, float max_norm, float* out_norm_coef, float* scratch,
                                 int apply, livae_stream_t stream) {
  LIVAE_CHECK_ARG(grads && out_norm_coef && scratch && n >= 0, "l2norm_clip: bad args");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = sgrid(n, 8);
  sumsq_kernel<<<grid, 256, 0, st>>>(grads, n, scratch);
  norm_finish_kernel<<<1, 256, 0, st>>>(scratch, grid, max_norm, out_norm_coef);
  if (apply) scale_kernel<<<sgrid(n, 4), 256, 0, st>>>(grads, n, out_norm_coef + 1);
  count_launch(apply ? 2 : 1);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int decoupled, float* step_dev,
                           const float* gscale_dev, int inc_step, livae_stream_t stream) {
  LIVAE_CHECK_ARG(p && g && m && v && step_dev && n >= 0, "adamw: bad args");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (n > 0)
    adamw_kernel<<<sgrid(n, 4), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                                              step_dev, gscale_dev);
  if (inc_step) inc_step_kernel<<<1, 1, 0, st>>>(step_dev);
  if (inc_step && n > 0) count_launch(1);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

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

// ---- row-wise variants (C/8 a power of two): one CTA per output row, so the only index arithmetic in
// the element loop is a shift and a mask (the flat kernels above spend ~5 64-bit divisions per 16-byte
// vector and are instruction-bound at ~2 TB/s), and the adjoint is separable: vertical taps straight
// from global memory into an fp32 row in shared memory, horizontal taps from there.
static constexpr int kUpRows = 1;   // output rows per CTA (more rows per CTA measured slower: less parallelism)

__global__ void __launch_bounds__(256) upsample_pad_fwd_row_kernel(const uint4* __restrict__ x, int H, int W, int logc8,
                                                                   uint4* __restrict__ out) {
  const int Ho = 2 * H + 2, Wo = 2 * W + 2, C8 = 1 << logc8;
  const int b = blockIdx.y;
  const int nv = Wo * C8;
  for (int Y = blockIdx.x * kUpRows; Y < min(Ho, (int)(blockIdx.x + 1) * kUpRows); ++Y) {
    const Up1D ty = up1d(Y, H);
    const uint4* r0 = x + ((int64_t)b * H + ty.i0) * W * C8;
    const uint4* r1 = x + ((int64_t)b * H + ty.i1) * W * C8;
    uint4* o = out + ((int64_t)b * Ho + Y) * nv;
    const float wy1 = ty.f, wy0 = 1.f - ty.f;
    for (int t = threadIdx.x; t < nv; t += 256) {
      const int X = t >> logc8, c = t & (C8 - 1);
      const Up1D tx = up1d(X, W);
      float v00[8], v01[8], v10[8], v11[8], r[8];
      bf8_to_f(__ldg(r0 + tx.i0 * C8 + c), v00);
      bf8_to_f(__ldg(r0 + tx.i1 * C8 + c), v01);
      bf8_to_f(__ldg(r1 + tx.i0 * C8 + c), v10);
      bf8_to_f(__ldg(r1 + tx.i1 * C8 + c), v11);
      const float wx1 = tx.f, wx0 = 1.f - tx.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = wy0 * (wx0 * v00[j] + wx1 * v01[j]) + wy1 * (wx0 * v10[j] + wx1 * v11[j]);
      o[t] = f_to_bf8(r);
    }
  }
}

static constexpr int kUpBwdRows = 1;   // low-res rows per CTA

__global__ void __launch_bounds__(256) upsample_pad_bwd_row_kernel(const uint4* __restrict__ g, int H, int W, int logc8,
                                                                   const uint4* __restrict__ mask_y,
                                                                   uint4* __restrict__ gx) {
  extern __shared__ __align__(16) float srow[];     // [Wo][C] fp32: vertical taps already applied
  const int Ho = 2 * H + 2, Wo = 2 * W + 2, C8 = 1 << logc8;
  const int b = blockIdx.y;
  const uint4* gb = g + (int64_t)b * Ho * Wo * C8;
  for (int i = blockIdx.x * kUpBwdRows; i < min(H, (int)(blockIdx.x + 1) * kUpBwdRows); ++i) {
    const Adj1D ay = adj1d(i, H);
    __syncthreads();                                  // the previous row's horizontal pass is done with srow
    for (int t = threadIdx.x; t < Wo * C8; t += 256) {
      uint4 q[6];
#pragma unroll
      for (int a = 0; a < 6; ++a)
        if (a < ay.n) q[a] = __ldg(gb + (int64_t)ay.Y[a] * Wo * C8 + t);
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
      for (int a = 0; a < 6; ++a)
        if (a < ay.n) {
          float v[8];
          bf8_to_f(q[a], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(ay.w[a], v[e], acc[e]);
        }
      float4* d = reinterpret_cast<float4*>(srow + (size_t)t * 8);
      d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    const int64_t obase = ((int64_t)b * H + i) * W * C8;
    for (int t = threadIdx.x; t < W * C8; t += 256) {
      const int j = t >> logc8, c = t & (C8 - 1);
      uint4 mk = make_uint4(0, 0, 0, 0);
      if (mask_y) mk = __ldg(mask_y + obase + t);
      const Adj1D ax = adj1d(j, W);
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      for (int d = 0; d < ax.n; ++d) {
        const float4* sp = reinterpret_cast<const float4*>(srow + ((size_t)ax.Y[d] * C8 + c) * 8);
        const float4 v0 = sp[0], v1 = sp[1];
        const float w = ax.w[d];
        acc[0] = fmaf(w, v0.x, acc[0]); acc[1] = fmaf(w, v0.y, acc[1]); acc[2] = fmaf(w, v0.z, acc[2]); acc[3] = fmaf(w, v0.w, acc[3]);
        acc[4] = fmaf(w, v1.x, acc[4]); acc[5] = fmaf(w, v1.y, acc[5]); acc[6] = fmaf(w, v1.z, acc[6]); acc[7] = fmaf(w, v1.w, acc[7]);
      }
      if (mask_y) {
        float m[8];
        bf8_to_f(mk, m);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!(m[e] > 0.f)) acc[e] = 0.f;
      }
      gx[obase + t] = f_to_bf8(acc);
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
template <typename T>
__global__ void __launch_bounds__(256) decfc_bwd_z_kernel(const T* __restrict__ gy, const T* __restrict__ y,
                                                          const float* __restrict__ w, int L, int C, int HW,
                                                          float* __restrict__ gz) {
  __shared__ float red[32];
  int b = blockIdx.x;
  int N = C * HW;
  const T* gyb = gy + (int64_t)b * N;
  const T* yb = y ? y + (int64_t)b * N : nullptr;
  for (int l0 = 0; l0 < L; l0 += 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      float g = Cvt<T>::ld(gyb, i);
      if (yb && !(Cvt<T>::ld(yb, i) > 0.f)) g = 0.f;
      int c = i % C, p = i / C;
      const float* wr = w + (int64_t)(c * HW + p) * L;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (l0 + t < L) a[t] = fmaf(g, wr[l0 + t], a[t]);
    }
    for (int t = 0; t < 4; ++t) {
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
    upsample_pad_fwd_row_kernel<<<dim3((2 * H + 2 + kUpRows - 1) / kUpRows, B), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, H, W, logc8, (uint4*)out);
  else if ((C & 7) == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0)
    upsample_pad_fwd_bf16v_kernel<<<sgrid(n / 8, 1), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, B, H, W, C / 8,
                                                                                   (uint4*)out);
  else
    upsample_pad_fwd_kernel<__nv_bfloat16><<<sgrid(n, 2), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, B, H, W, C, (__nv_bfloat16*)out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_upsample_pad_bwd_bf16(const void* g, int B, int H, int W, int C, const void* relu_mask_y,
                                           void* gx, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && H > 1 && W > 1 && C > 0, "upsample_pad_bwd_bf16: bad args");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g && gx, "upsample_pad_bwd_bf16: null pointer");
  if (int e = require_sm100()) return e;
  int64_t n = (int64_t)B * H * W * C;
  const int logc8 = (C & 7) == 0 ? log2_exact(C / 8) : -1;
  const size_t row_smem = (size_t)(2 * W + 2) * C * 4;
  if (logc8 >= 0 && B <= 65535 && row_smem <= 48 * 1024 && (((uintptr_t)g | (uintptr_t)gx | (uintptr_t)relu_mask_y) & 15) == 0)
    upsample_pad_bwd_row_kernel<<<dim3((H + kUpBwdRows - 1) / kUpBwdRows, B), 256, row_smem, (cudaStream_t)stream>>>((const uint4*)g, H, W, logc8,
                                                                                     (const uint4*)relu_mask_y, (uint4*)gx);
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
  if (gz) {
    decfc_bwd_z_kernel<__nv_bfloat16><<<B, 256, 0, st>>>(g, (const __nv_bfloat16*)nullptr, w, L, C, HW, gz);
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
    decfc_bwd_w_kernel<4, __nv_bfloat16><<<dim3(nb, ysplit), 128, 0, st>>>(g, (const __nv_bfloat16*)nullptr, z, B, L, C, HW,
                                                                            gw, gb);
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

// Engine 0: fp32 gather-GEMM for every convolution / linear layer of the step (reference
// model.py:203-214 STN, 289-303 encoder, 353-373 decoder, 29-43 / 87-98 plain VAE).
// It is the exact-fp32 engine (parity 1e-4 against the reference's ATen ops) and the home of the
// layers that are too thin for tensor cores (Cin = 1, Cout = 1, Linear heads); the 16-bit tcgen05
// engine (conv_tc.cu) takes the dense mid layers.  One register-tiled SIMT GEMM kernel is
// instantiated over three operand-gather functors (see conv_f32.cuh for P1/P2/P3); activation
// derivatives and max-pool routing are applied while LOADING the gradient operand, so neither a
// masked gradient nor an un-pooled tensor is ever written to HBM.
#include "conv_f32.cuh"

namespace livae {

static constexpr int BK = 16;

// ---------------------------------------------------------------------------------------------
struct P1 {  // big -> small (Conv2d forward / ConvTranspose2d dgrad)
  ConvGeom g; TensorRef big; const float* w; const float* bias; float* out; int act;
  static constexpr bool kAContigK = true, kBContigK = true;
  __device__ int M() const { return g.B * g.Hs * g.Ws; }
  __device__ int N() const { return g.Cs; }
  __device__ int K() const { return g.kh * g.kw * g.Cb; }
  __device__ float A(int m, int k) const {
    int ox = m % g.Ws; int t = m / g.Ws; int oy = t % g.Hs; int b = t / g.Hs;
    int cb = k % g.Cb; int tap = k / g.Cb; int kx = tap % g.kw; int ky = tap / g.kw;
    int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
    if (iy < 0 || iy >= g.Hb || ix < 0 || ix >= g.Wb) return 0.f;
    return ref_load(big, b, iy, ix, cb, g.Hb, g.Wb, g.Cb);
  }
  __device__ float Bv(int k, int n) const {
    int cb = k % g.Cb; int tap = k / g.Cb;
    return __ldg(w + ((int64_t)n * g.Cb + cb) * (g.kh * g.kw) + tap);
  }
  __device__ void store(int m, int n, float v) const {
    if (bias) v += bias[n];
    if (act == LIVAE_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == LIVAE_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    out[(int64_t)m * g.Cs + n] = v;
  }
};

struct P2 {  // small -> big (Conv2d dgrad / ConvTranspose2d forward)
  ConvGeom g; TensorRef small; const float* w; const float* bias; float* out; int act;
  static constexpr bool kAContigK = true, kBContigK = true;
  __device__ int M() const { return g.B * g.Hb * g.Wb; }
  __device__ int N() const { return g.Cb; }
  __device__ int K() const { return g.kh * g.kw * g.Cs; }
  __device__ float A(int m, int k) const {
    int ix = m % g.Wb; int t = m / g.Wb; int iy = t % g.Hb; int b = t / g.Hb;
    int cs = k % g.Cs; int tap = k / g.Cs; int kx = tap % g.kw; int ky = tap / g.kw;
    int ty = iy + g.pad - ky, tx = ix + g.pad - kx;
    if (ty < 0 || tx < 0) return 0.f;
    int oy = ty / g.stride, ox = tx / g.stride;
    if (oy * g.stride != ty || ox * g.stride != tx || oy >= g.Hs || ox >= g.Ws) return 0.f;
    return ref_load(small, b, oy, ox, cs, g.Hs, g.Ws, g.Cs);
  }
  __device__ float Bv(int k, int n) const {
    int cs = k % g.Cs; int tap = k / g.Cs;
    return __ldg(w + ((int64_t)cs * g.Cb + n) * (g.kh * g.kw) + tap);
  }
  __device__ void store(int m, int n, float v) const {
    if (bias) v += bias[n];
    if (act == LIVAE_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == LIVAE_ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    out[(int64_t)m * g.Cb + n] = v;
  }
};

// P2 specialised for a "full-map" kernel (Linear over a flattened feature map: Hs = Ws = 1,
// kh = Hb, kw = Wb, pad 0): each big-side pixel matches exactly one tap, so the data gradient is
// the plain GEMM [B, Cs] x [Cs, Cb*taps] with the output index permuted to NHWC.
struct P2Lin {
  ConvGeom g; TensorRef small; const float* w; float* out;
  static constexpr bool kAContigK = true, kBContigK = false;
  __device__ int M() const { return g.B; }
  __device__ int N() const { return g.Cb * g.kh * g.kw; }
  __device__ int K() const { return g.Cs; }
  __device__ float A(int m, int k) const { return ref_load(small, m, 0, 0, k, 1, 1, g.Cs); }
  __device__ float Bv(int k, int n) const { return __ldg(w + (int64_t)k * g.Cb * g.kh * g.kw + n); }
  __device__ void store(int m, int n, float v) const {
    int taps = g.kh * g.kw;
    int cb = n / taps, tap = n - cb * taps;
    out[((int64_t)m * taps + tap) * g.Cb + cb] = v;
  }
};

struct P3 {  // weight gradient, reduction over (b, oy, ox); split-K with fp32 atomics
  ConvGeom g; TensorRef big; TensorRef small; float* gw;
  static constexpr bool kAContigK = false, kBContigK = false;
  __device__ int M() const { return g.kh * g.kw * g.Cb; }
  __device__ int N() const { return g.Cs; }
  __device__ int K() const { return g.B * g.Hs * g.Ws; }
  __device__ float A(int m, int r) const {
    int cb = m % g.Cb; int tap = m / g.Cb; int kx = tap % g.kw; int ky = tap / g.kw;
    int ox = r % g.Ws; int t = r / g.Ws; int oy = t % g.Hs; int b = t / g.Hs;
    int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
    if (iy < 0 || iy >= g.Hb || ix < 0 || ix >= g.Wb) return 0.f;
    return ref_load(big, b, iy, ix, cb, g.Hb, g.Wb, g.Cb);
  }
  __device__ float Bv(int r, int n) const {
    int ox = r % g.Ws; int t = r / g.Ws; int oy = t % g.Hs; int b = t / g.Hs;
    return ref_load(small, b, oy, ox, n, g.Hs, g.Ws, g.Cs);
  }
  __device__ void store(int m, int n, float v) const {
    int cb = m % g.Cb; int tap = m / g.Cb;
    atomicAdd(gw + ((int64_t)n * g.Cb + cb) * (g.kh * g.kw) + tap, v);
  }
};

template <int BM, int BN, int TM, int TN, class P>
__global__ void __launch_bounds__(256) gemm_f32_kernel(P p) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int M = p.M(), N = p.N(), K = p.K();
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int kchunk = (K + gridDim.z - 1) / gridDim.z;
  kchunk = (kchunk + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int kk, mm;
      if (P::kAContigK) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < kend) ? p.A(m, k) : 0.f;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int kk, nn;
      if (P::kBContigK) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
      int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < kend) ? p.Bv(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (kbeg >= kend) return;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < N) p.store(m, n, acc[i][j]);
    }
  }
}

template <class P>
static int launch_gemm(const P& p, int M, int N, int K, bool splitk, cudaStream_t st) {
  auto zsplit = [&](int ctas) {
    if (!splitk) return 1;
    int z = (kNumSMs * 4 + ctas - 1) / ctas;
    int zmax = (K + BK * 8 - 1) / (BK * 8);
    if (z > zmax) z = zmax;
    if (z < 1) z = 1;
    if (z > 65535) z = 65535;
    return z;
  };
  if (N <= 8) {
    dim3 grid((M + 255) / 256, (N + 7) / 8, 1);
    grid.z = zsplit(grid.x * grid.y);
    gemm_f32_kernel<256, 8, 4, 2, P><<<grid, 256, 0, st>>>(p);
  } else if (N <= 32) {
    dim3 grid((M + 127) / 128, (N + 31) / 32, 1);
    grid.z = zsplit(grid.x * grid.y);
    gemm_f32_kernel<128, 32, 4, 4, P><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
    grid.z = zsplit(grid.x * grid.y);
    gemm_f32_kernel<64, 64, 4, 4, P><<<grid, 256, 0, st>>>(p);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// column sums of an [R, C]-shaped operand seen through a TensorRef: gb[c] += sum_r t(r, c)
__global__ void __launch_bounds__(256) colsum_kernel(TensorRef t, int B, int H, int W, int C,
                                                     float* __restrict__ gb, int rows_per_cta) {
  __shared__ float red[256];
  int R = B * H * W;
  int cols = C < 256 ? C : 256;
  int rgs = 256 / cols;
  int tid = threadIdx.x;
  int cl = tid % cols, rg = tid / cols;
  int r0 = blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  for (int c0 = 0; c0 < C; c0 += cols) {
    int c = c0 + cl;
    float a = 0.f;
    if (rg < rgs && c < C) {
      for (int r = r0 + rg; r < r1; r += rgs) {
        int x = r % W; int q = r / W; int y = q % H; int b = q / H;
        a += ref_load(t, b, y, x, c, H, W, C);
      }
    }
    red[tid] = a;
    __syncthreads();
    if (rg == 0 && c < C) {
      for (int j = 1; j < rgs; ++j) a += red[j * cols + cl];
      atomicAdd(gb + c, a);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ full, int B, int H,
                                                          int W, int C, float* __restrict__ pooled,
                                                          uint8_t* __restrict__ idx) {
  int Hp = H >> 1, Wp = W >> 1;
  int64_t n = (int64_t)B * Hp * Wp * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t q = i / C; int px = (int)(q % Wp); q /= Wp; int py = (int)(q % Hp);
    int b = (int)(q / Hp);
    const float* s = full + (((int64_t)b * H + 2 * py) * W + 2 * px) * C + c;
    // torch max_pool2d scan order (h then w), strictly-greater replaces
    float best = s[0]; int bi = 0;
    float v = s[C]; if (v > best) { best = v; bi = 1; }
    v = s[(int64_t)W * C]; if (v > best) { best = v; bi = 2; }
    v = s[(int64_t)W * C + C]; if (v > best) { best = v; bi = 3; }
    pooled[i] = best;
    idx[i] = (uint8_t)bi;
  }
}

static ConvGeom make_geom(const livae_conv_desc* d, int* Ho, int* Wo) {
  ConvGeom g;
  g.B = d->B; g.kh = d->kh; g.kw = d->kw; g.stride = d->stride; g.pad = d->pad;
  if (d->kind == LIVAE_CONV) {
    g.Hb = d->Hin; g.Wb = d->Win; g.Cb = d->Cin;
    g.Hs = (d->Hin + 2 * d->pad - d->kh) / d->stride + 1;
    g.Ws = (d->Win + 2 * d->pad - d->kw) / d->stride + 1;
    g.Cs = d->Cout;
    *Ho = g.Hs; *Wo = g.Ws;
  } else {
    g.Hs = d->Hin; g.Ws = d->Win; g.Cs = d->Cin;
    g.Hb = (d->Hin - 1) * d->stride - 2 * d->pad + d->kh;
    g.Wb = (d->Win - 1) * d->stride - 2 * d->pad + d->kw;
    g.Cb = d->Cout;
    *Ho = g.Hb; *Wo = g.Wb;
  }
  return g;
}

static int check_desc(const livae_conv_desc* d) {
  LIVAE_CHECK_ARG(d, "conv: null descriptor");
  LIVAE_CHECK_ARG(d->kind == LIVAE_CONV || d->kind == LIVAE_CONVT, "conv: bad kind %d", d->kind);
  LIVAE_CHECK_ARG(d->B >= 0 && d->Hin > 0 && d->Win > 0 && d->Cin > 0 && d->Cout > 0 && d->kh > 0 &&
                      d->kw > 0 && d->stride > 0 && d->pad >= 0,
                  "conv: bad sizes");
  LIVAE_CHECK_ARG(!d->pool || d->kind == LIVAE_CONV, "conv: pool only with LIVAE_CONV");
  LIVAE_CHECK_ARG((int64_t)d->B * d->Hin * d->Win * (int64_t)(d->Cin > d->Cout ? d->Cin : d->Cout) * 16 <
                      (int64_t)1 << 40, "conv: tensor too large");
  return 0;
}

}  // namespace livae

using namespace livae;

extern "C" void livae_conv_out_shape(const livae_conv_desc* d, int* Ho, int* Wo) {
  int h, w;
  make_geom(d, &h, &w);
  if (d->pool) { h >>= 1; w >>= 1; }
  *Ho = h; *Wo = w;
}

extern "C" int64_t livae_conv_fwd_ws_bytes(const livae_conv_desc* d) {
  if (!d || !d->pool) return 0;
  int h, w;
  ConvGeom g = make_geom(d, &h, &w);
  return (int64_t)g.B * g.Hs * g.Ws * g.Cs * 4;
}

extern "C" int livae_conv_fwd(const livae_conv_desc* d, const float* x, const float* w, const float* bias,
                              float* y, uint8_t* pool_idx, void* ws, livae_stream_t stream) {
  if (int e = check_desc(d)) return e;
  if (d->B == 0) return 0;
  LIVAE_CHECK_ARG(x && w && y, "conv_fwd: null pointer");
  LIVAE_CHECK_ARG(!d->pool || (pool_idx && ws), "conv_fwd: pool needs pool_idx and workspace");
  if (int e = require_sm100()) return e;
  if (d->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo;
  ConvGeom g = make_geom(d, &Ho, &Wo);
  TensorRef in{x, nullptr, nullptr, 0};
  if (d->kind == LIVAE_CONV) {
    LIVAE_CHECK_ARG(g.Hs > 0 && g.Ws > 0, "conv_fwd: empty output");
    LIVAE_CHECK_ARG(!d->pool || ((g.Hs & 1) == 0 && (g.Ws & 1) == 0), "conv_fwd: pool needs even output");
    float* out = d->pool ? (float*)ws : y;
    P1 p{g, in, w, bias, out, d->act};
    if (int e = launch_gemm(p, g.B * g.Hs * g.Ws, g.Cs, g.kh * g.kw * g.Cb, false, st)) return e;
    if (d->pool) {
      int64_t n = (int64_t)g.B * (g.Hs / 2) * (g.Ws / 2) * g.Cs;
      int grid = (int)((n + 255) / 256 < kNumSMs * 16 ? (n + 255) / 256 : kNumSMs * 16);
      maxpool_fwd_kernel<<<grid, 256, 0, st>>>(out, g.B, g.Hs, g.Ws, g.Cs, y, pool_idx);
      LIVAE_CUDA_LAUNCH_CHECK();
    }
  } else {
    P2 p{g, in, w, bias, y, d->act};
    if (int e = launch_gemm(p, g.B * g.Hb * g.Wb, g.Cb, g.kh * g.kw * g.Cs, false, st)) return e;
  }
  return 0;
}

// gy: gradient w.r.t. the layer's post-activation (post-pool) output y.  gw/gb are written (zeroed
// here, then accumulated with atomics); gx is written.  Any of gw, gb, gx may be NULL.
extern "C" int livae_conv_bwd(const livae_conv_desc* d, const float* x, const float* w, const float* y,
                              const float* gy, const uint8_t* pool_idx, float* gw, float* gb, float* gx,
                              livae_stream_t stream) {
  if (int e = check_desc(d)) return e;
  if (d->B == 0) return 0;
  LIVAE_CHECK_ARG(x && w && gy, "conv_bwd: null pointer");
  LIVAE_CHECK_ARG(d->act == LIVAE_ACT_NONE || y, "conv_bwd: activation needs y");
  LIVAE_CHECK_ARG(!d->pool || pool_idx, "conv_bwd: pool needs pool_idx");
  if (int e = require_sm100()) return e;
  if (d->B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo;
  ConvGeom g = make_geom(d, &Ho, &Wo);
  TensorRef grad{gy, d->act == LIVAE_ACT_NONE ? nullptr : y, d->pool ? pool_idx : nullptr, d->act};
  if (d->pool && !grad.yact) grad.yact = y ? y : gy;  // routing only (act none): yact unused by apply_dact
  TensorRef in{x, nullptr, nullptr, 0};
  size_t wbytes = (size_t)g.Cs * g.Cb * g.kh * g.kw * sizeof(float);
  cudaError_t ce;
  if (d->kind == LIVAE_CONV) {
    if (gw) {
      if ((ce = cudaMemsetAsync(gw, 0, wbytes, st)) != cudaSuccess) { set_error("memset gw"); return (int)ce; }
      P3 p{g, in, grad, gw};
      if (int e = launch_gemm(p, g.kh * g.kw * g.Cb, g.Cs, g.B * g.Hs * g.Ws, true, st)) return e;
    }
    if (gb) {
      if ((ce = cudaMemsetAsync(gb, 0, g.Cs * sizeof(float), st)) != cudaSuccess) { set_error("memset gb"); return (int)ce; }
      int R = g.B * g.Hs * g.Ws;
      int rows = (R + kNumSMs * 4 - 1) / (kNumSMs * 4);
      if (rows < 64) rows = 64;
      colsum_kernel<<<(R + rows - 1) / rows, 256, 0, st>>>(grad, g.B, g.Hs, g.Ws, g.Cs, gb, rows);
      LIVAE_CUDA_LAUNCH_CHECK();
    }
    if (gx) {
      if (g.Hs == 1 && g.Ws == 1 && g.pad == 0 && g.kh == g.Hb && g.kw == g.Wb && !d->pool) {
        P2Lin p{g, grad, w, gx};
        if (int e = launch_gemm(p, g.B, g.Cb * g.kh * g.kw, g.Cs, false, st)) return e;
      } else {
        P2 p{g, grad, w, nullptr, gx, LIVAE_ACT_NONE};
        if (int e = launch_gemm(p, g.B * g.Hb * g.Wb, g.Cb, g.kh * g.kw * g.Cs, false, st)) return e;
      }
    }
  } else {
    if (gw) {
      if ((ce = cudaMemsetAsync(gw, 0, wbytes, st)) != cudaSuccess) { set_error("memset gw"); return (int)ce; }
      P3 p{g, grad, in, gw};
      if (int e = launch_gemm(p, g.kh * g.kw * g.Cb, g.Cs, g.B * g.Hs * g.Ws, true, st)) return e;
    }
    if (gb) {
      if ((ce = cudaMemsetAsync(gb, 0, g.Cb * sizeof(float), st)) != cudaSuccess) { set_error("memset gb"); return (int)ce; }
      int R = g.B * g.Hb * g.Wb;
      int rows = (R + kNumSMs * 4 - 1) / (kNumSMs * 4);
      if (rows < 64) rows = 64;
      colsum_kernel<<<(R + rows - 1) / rows, 256, 0, st>>>(grad, g.B, g.Hb, g.Wb, g.Cb, gb, rows);
      LIVAE_CUDA_LAUNCH_CHECK();
    }
    if (gx) {
      P1 p{g, grad, w, nullptr, gx, LIVAE_ACT_NONE};
      if (int e = launch_gemm(p, g.B * g.Hs * g.Ws, g.Cs, g.kh * g.kw * g.Cb, false, st)) return e;
    }
  }
  return 0;
}

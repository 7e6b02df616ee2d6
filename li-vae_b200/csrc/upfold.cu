// Decoder blocks d1-d3: Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv2d(3x3) -> ReLU (reference
// model.py:356-368) WITHOUT the up-sampled tensor ("phase folding").
//
// Up-sampling and padding are linear and act per channel, so each of the four output phases (py,px) of the block
// is a 3x3 convolution of the LOW-resolution input with folded weights
//     Wf[(py,px,co)][ci][DY][DX] = sum_{ky,kx} A[py][DY][ky] * A[px][DX][kx] * w[co][ci][ky][kx]
// (A = the 0.25 / 0.75 bilinear weights of align_corners=False, below): ONE convolution Cin -> 4*Cout over a
// quarter of the pixels, same MACs, N = 4*Cout instead of Cout on the tensor cores (the N <= 64 layers were bound
// by the UMMA operand read), and the 4x larger up-sampled activation and its gradient are never written or read.
// With zero padding of the low-resolution input (TMA out-of-bounds fill) the folded form F0 is exact wherever no
// border rule is involved.  The borders: write P = Ry (x) Rx [x] for the true up-sampled + reflect-padded tensor
// (Ry, Rx the 1-D operators incl. index clamping and reflection) and P0 = R0y (x) R0x [x] for what F0 implicitly
// convolves (interior formula on the zero-extended input).  Then
//     P - P0 = (Ry - R0y) (x) Rx  +  R0y (x) (Rx - R0x) = T1 + T2,
// T1 lives on padded rows {-1, 0, 2h-1, 2h}, T2 on padded columns {-1, 0, 2w-1, 2w}: two batches of 4-row STRIPS
// (the lr strips stored transposed so that both batches have one shape), convolved with the layer's own 3x3
// weights by the ordinary kernels (pad 0: [*,4,2w+2,Cin] -> [*,2,2w,Cout]); the result corrects the two outermost
// output rows / columns and is added in the folded convolution's epilogue before the ReLU (epilogue_upfold).
// Backward: every piece is linear, so the data gradient is F0^T gz (one convolution 4*Cout -> Cin whose operand is
// gz read block-wise through a strided tensor map) plus the strips' adjoint scattered into the two outermost
// rows / columns of gx; the weight gradient is the folded one mapped back through A plus the strips' own.
//
// The algebra (folded weights, the four correction terms, fold / unfold adjoint pair) is restated in numpy in
// oracle/folded_upconv.py and pinned on the CPU against torch's Upsample -> ReflectionPad2d -> Conv2d.
#include "tc_common.cuh"

namespace livae {
namespace tc {

// A[p][d + 1][k]: weight of filter tap k (0..2) of output phase p on the low-resolution neighbour at offset d.
// Output 2i+p reads up-sampled positions 2i+p-1+k; U(2m) = 0.25 x[m-1] + 0.75 x[m], U(2m+1) = 0.75 x[m] + 0.25 x[m+1].
__device__ __forceinline__ float fold_a(int p, int d1, int k) {
  const float T[2][3][3] = {{{0.75f, 0.25f, 0.f}, {0.25f, 0.75f, 0.75f}, {0.f, 0.f, 0.25f}},
                            {{0.25f, 0.f, 0.f}, {0.75f, 0.75f, 0.25f}, {0.f, 0.25f, 0.75f}}};
  return T[p][d1][k];
}

// wf bf16 [9][4*Cout][Cin] (forward: N = (py,px,co), K = ci), wd bf16 [9][Cin][4*Cout] (data gradient: N = ci, K = (py,px,co));
// tap = (DY+1)*3 + (DX+1)
__global__ void upfold_pack_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd) {
  const int N = 4 * Cout, total = 9 * N * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin; int t = i / Cin; const int n = t % N; const int tap = t / N;
    const int d1y = tap / 3, d1x = tap % 3;
    const int ph = n / Cout, co = n - ph * Cout;
    const float* wp = w + ((int64_t)co * Cin + ci) * 9;
    float s = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) s += fold_a(ph >> 1, d1y, ky) * fold_a(ph & 1, d1x, kx) * wp[ky * 3 + kx];
    const __nv_bfloat16 b = __float2bfloat16_rn(s);
    wf[i] = b;
    wd[((int64_t)tap * Cin + ci) * N + n] = b;
  }
}

// gw[co][ci][ky][kx] = sum_{py,px,DY,DX} A[py][DY][ky] A[px][DX][kx] acc[tap][ci][(py,px,co)]  (+ the strips' gradients)
__global__ void upfold_unfold_kernel(const float* __restrict__ acc, const float* __restrict__ gw_tb,
                                     const float* __restrict__ gw_lr, int Cout, int Cin, float* __restrict__ gw) {
  const int total = Cout * Cin * 9, N = 4 * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kx = i % 3; int t = i / 3; const int ky = t % 3; t /= 3; const int ci = t % Cin; const int co = t / Cin;
    float s = 0.f;
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int d1y = 0; d1y < 3; ++d1y) {
        const float ay = fold_a(py, d1y, ky);
        if (ay == 0.f) continue;
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
          for (int d1x = 0; d1x < 3; ++d1x) {
            const float ax = fold_a(px, d1x, kx);
            if (ax == 0.f) continue;
            s += ay * ax * acc[((int64_t)(d1y * 3 + d1x) * Cin + ci) * N + (py * 2 + px) * Cout + co];
          }
      }
    if (gw_tb) s += gw_tb[i];
    if (gw_lr) s += gw_lr[((int64_t)co * Cin + ci) * 9 + kx * 3 + ky];      // computed on the transposed strips
    gw[i] = s;
  }
}

__device__ __forceinline__ void bf8_axpy(float (&a)[8], float s, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[2 * i] += s * __uint_as_float(w[i] << 16);
    a[2 * i + 1] += s * __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 bf8_pack(const float (&a)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Weight of low-resolution element j (0..n-1) in the TRUE up-sampled + reflect-padded value at padded position X
// (-1..2n): bilinear with index clamping, positions -1 and 2n mirrored onto 1 and 2n-2.
__device__ __forceinline__ float up_true_w(int X, int j, int n) {
  const int Xc = X < 0 ? 1 : (X >= 2 * n ? 2 * n - 2 : X);
  const int m = Xc >> 1;
  if (Xc & 1) return (m == j ? 0.75f : 0.f) + ((m + 1 < n ? m + 1 : n - 1) == j ? 0.25f : 0.f);
  return ((m > 0 ? m - 1 : 0) == j ? 0.25f : 0.f) + (m == j ? 0.75f : 0.f);
}
// The same for the interior formula on the zero-extended input (what the folded convolution implies), X in -1..2n
__device__ __forceinline__ float up_zero_w(int X, int j) {
  const int m = (X + 2) / 2 - 1;            // floor(X / 2) for X >= -2
  if (X - 2 * m) return (m == j ? 0.75f : 0.f) + (m + 1 == j ? 0.25f : 0.f);
  return (m - 1 == j ? 0.25f : 0.f) + (m == j ? 0.75f : 0.f);
}

// Border strips of the layer input (see the header): s_tb [2][B][4][2w+2][Cin], s_lr [2][B][4][2h+2][Cin] (bf16).
//   top    rows: T1(-1) = Rx[0.5 x[0,:] + 0.25 x[1,:]],  T1(0) = Rx[0.25 x[0,:]], 0, 0
//   bottom rows: 0, 0, T1(2h-1) = Rx[0.25 x[h-1,:]],  T1(2h) = Rx[0.25 x[h-2,:] + 0.5 x[h-1,:]]
//   left / right: the same with rows and columns exchanged and R0y (zero-extended) in place of Rx
// One thread = one strip position and 8 channels: the up-sampled values u0 / u1 of the outermost and the second
// low-resolution line at that position, from which both live rows follow; the two zero rows are written as well
// (every strip is a slice of one tall image for the convolution kernels, see livae/ops.py).
__global__ void __launch_bounds__(256) upfold_strips_kernel(const uint4* __restrict__ x, int B, int h, int w, int C8,
                                                            uint4* __restrict__ s_tb, uint4* __restrict__ s_lr) {
  const int Ltb = 2 * w + 2, Llr = 2 * h + 2;
  const int64_t n_tb = (int64_t)2 * B * Ltb * C8, n_lr = (int64_t)2 * B * Llr * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tb + n_lr; i += (int64_t)gridDim.x * blockDim.x) {
    const bool lr = i >= n_tb;
    int64_t t = lr ? i - n_tb : i;
    const int L = lr ? Llr : Ltb, n = lr ? h : w, m = lr ? w : h;  // strip length, low-res extent along / across the strip
    const int c = (int)(t % C8); t /= C8;
    const int pos = (int)(t % L); t /= L;
    const int b = (int)(t % B); const int side = (int)(t / B);
    const int l0 = side ? m - 1 : 0, l1 = side ? m - 2 : 1;       // outermost and second line (rows for tb, columns for lr)
    const int X = pos - 1;
    const int Xe = lr ? X : (X < 0 ? 1 : (X >= 2 * n ? 2 * n - 2 : X));   // tb: the padding ring mirrors
    const int jm = (Xe + 2) / 2 - 1;                               // floor(Xe / 2)
    float u0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, u1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = jm - 1; j <= jm + 1; ++j) {                       // at most two neighbours carry weight
      if (j < 0 || j >= n) continue;
      const float wj = lr ? up_zero_w(X, j) : up_true_w(X, j, n);
      if (wj == 0.f) continue;
      const int64_t p0 = lr ? ((int64_t)b * h + j) * w + l0 : ((int64_t)b * h + l0) * w + j;
      const int64_t p1 = lr ? ((int64_t)b * h + j) * w + l1 : ((int64_t)b * h + l1) * w + j;
      bf8_axpy(u0, wj, __ldg(x + p0 * C8 + c));
      bf8_axpy(u1, wj, __ldg(x + p1 * C8 + c));
    }
    float ra[8], rb[8];                                            // ra: the row next to the image (0.25 u0), rb: the outer one
#pragma unroll
    for (int k = 0; k < 8; ++k) { ra[k] = 0.25f * u0[k]; rb[k] = 0.5f * u0[k] + 0.25f * u1[k]; }
    uint4* dst = (lr ? s_lr : s_tb) + ((((int64_t)side * B + b) * 4) * L + pos) * C8 + c;
    const int64_t rs = (int64_t)L * C8;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (side == 0) { dst[0] = bf8_pack(rb); dst[rs] = bf8_pack(ra); dst[2 * rs] = z; dst[3 * rs] = z; }
    else { dst[0] = z; dst[rs] = z; dst[2 * rs] = bf8_pack(ra); dst[3 * rs] = bf8_pack(rb); }
  }
}

// g_tb [2][B][4][2w][C]: rows 0,1 = output rows {0,1} / {2h-2,2h-1} of gz [B,2h,2w,C], rows 2,3 = 0 (the slack rows of the
// tall image); g_lr [2][B][4][2h][C] = columns {0,1} / {2w-2,2w-1}, transposed
__global__ void __launch_bounds__(256) upfold_gather_kernel(const uint4* __restrict__ gz, int B, int H, int W, int C8,
                                                            uint4* __restrict__ g_tb, uint4* __restrict__ g_lr) {
  const int64_t n_tb = (int64_t)2 * B * 4 * W * C8, n_lr = (int64_t)2 * B * 4 * H * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tb + n_lr; i += (int64_t)gridDim.x * blockDim.x) {
    const bool lr = i >= n_tb;
    int64_t t = lr ? i - n_tb : i;
    const int L = lr ? H : W;
    const int c = (int)(t % C8); t /= C8;
    const int pos = (int)(t % L); t /= L;
    const int r = (int)(t & 3); t >>= 2;
    const int b = (int)(t % B); const int side = (int)(t / B);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < 2) {
      const int Y = lr ? pos : (side ? H - 2 + r : r), X = lr ? (side ? W - 2 + r : r) : pos;
      v = __ldg(gz + (((int64_t)b * H + Y) * W + X) * C8 + c);
    }
    (lr ? g_lr : g_tb)[lr ? i - n_tb : i] = v;
  }
}

// The two outermost rows / columns of y [B,2h,2w,C] hold acc + bias (epilogue_upfold): y = ReLU(y + correction).
// corr_tb fp32 [2][B][4][2w][C] (rows 0,1 used: output rows {0,1} / {2h-2,2h-1}), corr_lr fp32 [2][B][4][2h][C] (columns)
__global__ void __launch_bounds__(256) upfold_ring_kernel(const float4* __restrict__ corr_tb, const float4* __restrict__ corr_lr,
                                                          int B, int H, int W, int C8, uint4* __restrict__ y) {
  const int nring = 4 * W + 4 * (H - 4);
  const int64_t total = (int64_t)B * nring * C8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int c = (int)(t % C8); t /= C8;
    const int r = (int)(t % nring); const int b = (int)(t / nring);
    int Y, X;
    if (r < 4 * W) { const int k = r / W; Y = k < 2 ? k : H - 4 + k; X = r - k * W; }
    else { const int q = r - 4 * W; const int k = q & 3; Y = 2 + (q >> 2); X = k < 2 ? k : W - 4 + k; }
    uint4* yp = y + (((int64_t)b * H + Y) * W + X) * C8 + c;
    const uint4 v = *yp;
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
    float a[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[2 * k] = __uint_as_float(vw[k] << 16); a[2 * k + 1] = __uint_as_float(vw[k] & 0xffff0000u); }
    auto add = [&](const float4* p) {
      const float4 u = __ldg(p), w4 = __ldg(p + 1);
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += w4.x; a[5] += w4.y; a[6] += w4.z; a[7] += w4.w;
    };
    if (Y < 2) add(corr_tb + ((((int64_t)b * 4 + Y) * W + X) * C8 + c) * 2);
    if (Y >= H - 2) add(corr_tb + (((((int64_t)B + b) * 4 + (Y - (H - 2))) * W + X) * C8 + c) * 2);
    if (X < 2) add(corr_lr + ((((int64_t)b * 4 + X) * H + Y) * C8 + c) * 2);
    if (X >= W - 2) add(corr_lr + (((((int64_t)B + b) * 4 + (X - (W - 2))) * H + Y) * C8 + c) * 2);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], 0.f);
    *yp = bf8_pack(a);
  }
}

// gx[b,i,j,:] += (x > 0) * adjoint of upfold_strips_kernel applied to the strips' gradients
//   gs_tb fp32 [2][B][4][2w+2][Cin], gs_lr fp32 [2][B][4][2h+2][Cin]; touches rows {0,1,h-2,h-1} and columns {0,1,w-2,w-1}
__global__ void __launch_bounds__(256) upfold_patch_kernel(const float4* __restrict__ gs_tb, const float4* __restrict__ gs_lr,
                                                           const uint4* __restrict__ xmask, int B, int h, int w, int C8,
                                                           uint4* __restrict__ gx) {
  const int Ltb = 2 * w + 2, Llr = 2 * h + 2;
  const int nring = 4 * w + 4 * (h - 4);                 // the two outermost rows / columns (h, w >= 4)
  const int64_t total = (int64_t)B * nring * C8;
  for (int64_t tix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tix < total; tix += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = tix;
    const int c = (int)(t % C8); t /= C8;
    const int rr = (int)(t % nring); const int b = (int)(t / nring);
    int i, j;
    if (rr < 4 * w) { const int k = rr / w; i = k < 2 ? k : h - 4 + k; j = rr - k * w; }
    else { const int q = rr - 4 * w; const int k = q & 3; i = 2 + (q >> 2); j = k < 2 ? k : w - 4 + k; }
    const int64_t idx = (((int64_t)b * h + i) * w + j) * C8 + c;
    const bool rowb = i < 2 || i >= h - 2, colb = j < 2 || j >= w - 2;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto add = [&](const float4* base, float s) {
      const float4 u = __ldg(base), v = __ldg(base + 1);
      a[0] += s * u.x; a[1] += s * u.y; a[2] += s * u.z; a[3] += s * u.w;
      a[4] += s * v.x; a[5] += s * v.y; a[6] += s * v.z; a[7] += s * v.w;
    };
    if (rowb) {
      // strip row r of side s carries weight cw on low-resolution row i (the transpose of the table in upfold_strips_kernel)
      for (int sr = 0; sr < 4; ++sr) {
        const int side = sr >> 1, r = side ? 2 + (sr & 1) : (sr & 1);
        const int l0 = side ? h - 1 : 0, l1 = side ? h - 2 : 1;
        const float c0 = side ? (r == 3 ? 0.5f : 0.25f) : (r == 0 ? 0.5f : 0.25f);
        const float c1 = side ? (r == 3 ? 0.25f : 0.f) : (r == 0 ? 0.25f : 0.f);
        const float cw = (i == l0 ? c0 : 0.f) + (i == l1 ? c1 : 0.f);
        if (cw == 0.f) continue;
        const float4* row = gs_tb + ((((int64_t)side * B + b) * 4 + r) * Ltb) * (2 * C8) + 2 * c;
        for (int X = max(-1, 2 * j - 3); X <= min(2 * w, 2 * j + 4); ++X) {
          const float wj = up_true_w(X, j, w);
          if (wj != 0.f) add(row + (int64_t)(X + 1) * (2 * C8), cw * wj);
        }
      }
    }
    if (colb) {
      for (int sr = 0; sr < 4; ++sr) {
        const int side = sr >> 1, r = side ? 2 + (sr & 1) : (sr & 1);
        const int l0 = side ? w - 1 : 0, l1 = side ? w - 2 : 1;
        const float c0 = side ? (r == 3 ? 0.5f : 0.25f) : (r == 0 ? 0.5f : 0.25f);
        const float c1 = side ? (r == 3 ? 0.25f : 0.f) : (r == 0 ? 0.25f : 0.f);
        const float cw = (j == l0 ? c0 : 0.f) + (j == l1 ? c1 : 0.f);
        if (cw == 0.f) continue;
        const float4* row = gs_lr + ((((int64_t)side * B + b) * 4 + r) * Llr) * (2 * C8) + 2 * c;
        for (int Y = max(-1, 2 * i - 1); Y <= min(2 * h, 2 * i + 2); ++Y) {
          const float wi = up_zero_w(Y, i);
          if (wi != 0.f) add(row + (int64_t)(Y + 1) * (2 * C8), cw * wi);
        }
      }
    }
    const uint4 m = __ldg(xmask + idx);
    const uint4 g = gx[idx];
    const uint32_t mw[4] = {m.x, m.y, m.z, m.w}, gw_[4] = {g.x, g.y, g.z, g.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float m0 = __uint_as_float(mw[k] << 16), m1 = __uint_as_float(mw[k] & 0xffff0000u);
      o[2 * k] = __uint_as_float(gw_[k] << 16) + (m0 > 0.f ? a[2 * k] : 0.f);
      o[2 * k + 1] = __uint_as_float(gw_[k] & 0xffff0000u) + (m1 > 0.f ? a[2 * k + 1] : 0.f);
    }
    gx[idx] = bf8_pack(o);
  }
}

static void block_taps(int sign, int* tdy, int* tdx, int* tw) {
  for (int t = 0; t < 9; ++t) { tdy[t] = sign * (t / 3 - 1); tdx[t] = sign * (t % 3 - 1); tw[t] = t; }
}
static int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;

// N = 4*Cout must be one UMMA (<= 256 columns) for the weight gradient; Cout a multiple of 32 for the epilogue
extern "C" int livae_upfold_supported(int B, int h, int w, int Cin, int Cout) {
  return B > 0 && h >= 8 && w >= 8 && h <= 120 && w <= 120 && h % 8 == 0 && w % 8 == 0 &&     // strip kernels: 2w, 2h multiples of 16
         Cin % 64 == 0 && Cin <= 512 && (Cout == 32 || Cout == 64) ? 1 : 0;
}

extern "C" int livae_upfold_pack(const float* w, int Cout, int Cin, void* wf, void* wd, livae_stream_t stream) {
  LIVAE_CHECK_ARG(w && wf && wd && Cout > 0 && Cin > 0, "upfold_pack: bad args");
  if (int e = require_sm100()) return e;
  const int n = 9 * 4 * Cout * Cin;
  upfold_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_upfold_strips(const void* x, int B, int h, int w, int Cin, void* s_tb, void* s_lr, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x && s_tb && s_lr && h >= 2 && w >= 2 && (Cin & 7) == 0, "upfold_strips: bad args");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)s_tb | (uintptr_t)s_lr) & 15) == 0, "upfold_strips: alignment");
  if (int e = require_sm100()) return e;
  const int64_t n = (int64_t)2 * B * (2 * w + 2 + 2 * h + 2) * (Cin / 8);
  upfold_strips_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, B, h, w, Cin / 8, (uint4*)s_tb, (uint4*)s_lr);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// y bf16 [B,2h,2w,Cout] = ReLU(F0(x) + bias), except the two outermost rows / columns = F0(x) + bias (finish with livae_upfold_ring)
extern "C" int livae_upfold_fwd(const void* x, const void* wf, const float* bias, int B, int h, int w, int Cin, int Cout,
                                void* y, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(livae_upfold_supported(B, h, w, Cin, Cout), "upfold_fwd: shape not supported");
  LIVAE_CHECK_ARG(x && wf && y, "upfold_fwd: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)wf | (uintptr_t)y) & 15) == 0, "upfold_fwd: alignment");
  if (int e = require_sm100()) return e;
  int tdy[9], tdx[9], tw[9];
  block_taps(1, tdy, tdx, tw);
  int rc = launch_conv_tc_halo(x, B, h, w, Cin, wf, 9, 4 * Cout, h, w, h, w, 1, 0, 0, 1, 9, tdy, tdx, tw, y, 0, bias,
                               LIVAE_ACT_RELU, nullptr, (cudaStream_t)stream, HaloOpts{0, 3, nullptr, Cout});
  if (rc == 1) { set_error("upfold_fwd: shape rejected by the halo kernel"); return -1; }
  return rc;
}

// the two outermost rows / columns of y: y = ReLU(y + correction); corr_tb fp32 [2B,4,2w,Cout], corr_lr fp32 [2B,4,2h,Cout]
extern "C" int livae_upfold_ring(const float* corr_tb, const float* corr_lr, int B, int h, int w, int Cout, void* y,
                                 livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(corr_tb && corr_lr && y && h >= 2 && w >= 2 && (Cout & 7) == 0, "upfold_ring: bad args");
  LIVAE_CHECK_ARG((((uintptr_t)corr_tb | (uintptr_t)corr_lr | (uintptr_t)y) & 15) == 0, "upfold_ring: alignment");
  if (int e = require_sm100()) return e;
  const int64_t n = (int64_t)B * (8 * w + 8 * h - 16) * (Cout / 8);
  upfold_ring_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float4*)corr_tb, (const float4*)corr_lr, B, 2 * h, 2 * w,
                                                                    Cout / 8, (uint4*)y);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_upfold_gather(const void* gz, int B, int h, int w, int Cout, void* g_tb, void* g_lr, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(gz && g_tb && g_lr && h >= 1 && w >= 1 && (Cout & 7) == 0, "upfold_gather: bad args");
  LIVAE_CHECK_ARG((((uintptr_t)gz | (uintptr_t)g_tb | (uintptr_t)g_lr) & 15) == 0, "upfold_gather: alignment");
  if (int e = require_sm100()) return e;
  const int64_t n = (int64_t)2 * B * 4 * (2 * w + 2 * h) * (Cout / 8);
  upfold_gather_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const uint4*)gz, B, 2 * h, 2 * w, Cout / 8, (uint4*)g_tb, (uint4*)g_lr);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// gx bf16 [B,h,w,Cin] = F0^T gz * (x_mask > 0); gz bf16 [B,2h,2w,Cout] read block-wise; the borders are completed by livae_upfold_patch
extern "C" int livae_upfold_dgrad(const void* gz, const void* wd, const void* x_mask, int B, int h, int w, int Cin, int Cout,
                                  void* gx, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(livae_upfold_supported(B, h, w, Cin, Cout), "upfold_dgrad: shape not supported");
  LIVAE_CHECK_ARG(gz && wd && gx, "upfold_dgrad: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)gz | (uintptr_t)wd | (uintptr_t)gx | (uintptr_t)x_mask) & 15) == 0, "upfold_dgrad: alignment");
  if (int e = require_sm100()) return e;
  int tdy[9], tdx[9], tw[9];
  block_taps(-1, tdy, tdx, tw);          // gx[i] = sum_D Wf[.,.,D]^T gz_blocks[i - D]
  int rc = launch_conv_tc_halo(gz, B, h, w, 4 * Cout, wd, 9, Cin, h, w, h, w, 1, 0, 0, 1, 9, tdy, tdx, tw, gx, 0, nullptr,
                               LIVAE_ACT_NONE, x_mask, (cudaStream_t)stream, HaloOpts{1, 0, nullptr, 0});
  if (rc == 1) { set_error("upfold_dgrad: shape rejected by the halo kernel"); return -1; }
  return rc;
}

extern "C" int livae_upfold_patch(const float* gs_tb, const float* gs_lr, const void* x_mask, int B, int h, int w, int Cin,
                                  void* gx, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(gs_tb && gs_lr && x_mask && gx && h >= 4 && w >= 4 && (Cin & 7) == 0, "upfold_patch: bad args");
  LIVAE_CHECK_ARG((((uintptr_t)gs_tb | (uintptr_t)gs_lr | (uintptr_t)x_mask | (uintptr_t)gx) & 15) == 0, "upfold_patch: alignment");
  if (int e = require_sm100()) return e;
  const int64_t n = (int64_t)B * (4 * w + 4 * h - 16) * (Cin / 8);
  upfold_patch_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float4*)gs_tb, (const float4*)gs_lr, (const uint4*)x_mask,
                                                                     B, h, w, Cin / 8, (uint4*)gx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t livae_upfold_wgrad_ws_bytes(int Cin, int Cout) { return (int64_t)9 * Cin * 4 * Cout * 4; }

// gw fp32 [Cout][Cin][3][3] = A^T (sum x^T gz_blocks) + gw_tb + transpose(gw_lr)   (gw_tb / gw_lr may be NULL)
extern "C" int livae_upfold_wgrad(const void* x, const void* gz, const float* gw_tb, const float* gw_lr, int B, int h, int w,
                                  int Cin, int Cout, float* gw, void* ws, livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(livae_upfold_supported(B, h, w, Cin, Cout), "upfold_wgrad: shape not supported");
  LIVAE_CHECK_ARG(x && gz && gw && ws, "upfold_wgrad: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)gz | (uintptr_t)ws) & 15) == 0, "upfold_wgrad: alignment");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemsetAsync(ws, 0, (size_t)livae_upfold_wgrad_ws_bytes(Cin, Cout), st);
  if (ce != cudaSuccess) { set_error("upfold_wgrad memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  livae_tc_conv_desc d;
  d.B = B; d.Hin = h; d.Win = w; d.Cin = Cin; d.Cout = 4 * Cout; d.kh = 3; d.kw = 3; d.stride = 1; d.pad = 1; d.act = 0; d.out_f32 = 0;
  int rc = launch_wgrad_halo(&d, x, gz, (float*)ws, h, w, st, 2);
  if (rc == 1) { set_error("upfold_wgrad: shape rejected by the halo kernel"); return -1; }
  if (rc != 0) return rc;
  const int n = Cout * Cin * 9;
  upfold_unfold_kernel<<<(n + 255) / 256, 256, 0, st>>>((const float*)ws, gw_tb, gw_lr, Cout, Cin, gw);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// a3: default_transform on the device (reference data.py:78-116) and the paired random rotation + centre crop +
// per-patch min-max of PairedAdaptiveLatticeDataset.__getitem__ (data.py:694-735).
//
// The random draws stay on the host (the reference takes them from Python's `random`; livae/data.py draws
// them in the same order) and arrive as per-patch parameter arrays.  The resampling follows torchvision's
// tensor path, which is what the reference executes: _get_inverse_affine_matrix in double, the matrix rounded
// to fp32, _gen_affine_grid + grid_sample(bilinear, zeros, align_corners=False) in fp32; TF.rotate(fill=0)
// multiplies the sampled value by the identically sampled all-ones mask (_apply_grid_transform).
//
// Both kernels are HBM-bound gathers: one CTA per patch walks its rows with coalesced stores; the four-tap
// reads of one output row hit at most a few source rows that stay in L1.  Algorithmic bytes per patch:
// augment: 2*S*S*4 (read + write); rotate_crop: S*S*4 read (the part of the source a rotated P-crop touches
// is at most P*sqrt2 wide) + P*P*4 written, read back from L2 once for the min-max.
#include "common.cuh"

namespace livae {

// torchvision's fp32 grid: g = (xb*m0 + yb*m1 + m2) with m pre-divided by S/2, then ((g+1)*S-1)/2
__device__ __forceinline__ float tv_unnormalise(float xb, float yb, float m0, float m1, float m2, int S) {
  const float g = fmaf(yb, m1, xb * m0) + m2;
  return ((g + 1.f) * (float)S - 1.f) * 0.5f;
}

struct Tap { float v, mask; };

__device__ __forceinline__ Tap bilinear_zeros(const float* __restrict__ src, int S, float ix, float iy) {
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float tx = ix - fx0, ty = iy - fy0;
  const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  const bool xa = x0 >= 0 && x0 < S, xb = x0 + 1 >= 0 && x0 + 1 < S;
  const bool ya = y0 >= 0 && y0 < S, yb = y0 + 1 >= 0 && y0 + 1 < S;
  Tap t{0.f, 0.f};
  if (ya && xa) { t.v += w00 * src[y0 * S + x0]; t.mask += w00; }
  if (ya && xb) { t.v += w01 * src[y0 * S + x0 + 1]; t.mask += w01; }
  if (yb && xa) { t.v += w10 * src[(y0 + 1) * S + x0]; t.mask += w10; }
  if (yb && xb) { t.v += w11 * src[(y0 + 1) * S + x0 + 1]; t.mask += w11; }
  return t;
}

// out[i, j] = flip(scale(in))[(i - shift_y) mod S, (j - shift_x) mod S]            (data.py:85-93, 105-114)
__global__ void __launch_bounds__(256) augment_kernel(const float* __restrict__ in, int N, int S,
                                                      const float* __restrict__ scale,
                                                      const int32_t* __restrict__ flags,
                                                      const int32_t* __restrict__ shift,
                                                      float* __restrict__ out) {
  const int n = blockIdx.x;
  const float* src = in + (int64_t)n * S * S;
  float* dst = out + (int64_t)n * S * S;
  // matrix [1/s, 0, 0; 0, 1/s, 0] in double like Python, rounded to fp32, divided by S/2 in fp32
  const float m = (float)(1.0 / (double)scale[n]) / (0.5f * (float)S);
  const int fl = flags[n];
  int sy = shift[2 * n] % S, sx = shift[2 * n + 1] % S;
  if (sy < 0) sy += S;
  if (sx < 0) sx += S;
  const float c = 0.5f * (float)(S - 1);
  const int rows_per_blk = (S + gridDim.y - 1) / gridDim.y;
  const int r_beg = blockIdx.y * rows_per_blk, r_end = min(S, r_beg + rows_per_blk);
  for (int i = r_beg * S + threadIdx.x; i < r_end * S; i += blockDim.x) {
    const int r = i / S, col = i - r * S;
    int yy = r - sy, xx = col - sx;               // undo the roll
    if (yy < 0) yy += S;
    if (xx < 0) xx += S;
    if (fl & 2) yy = S - 1 - yy;                  // undo vflip
    if (fl & 1) xx = S - 1 - xx;                  // undo hflip
    const float ix = tv_unnormalise((float)xx - c, (float)yy - c, m, 0.f, 0.f, S);
    const float iy = tv_unnormalise((float)xx - c, (float)yy - c, 0.f, m, 0.f, S);
    dst[i] = (fl & 4) ? src[yy * S + xx] : bilinear_zeros(src, S, ix, iy).v;   // bit2: flips + roll only
  }
}

// mode 0: centre crop; mode 1: TF.rotate(angle_deg, bilinear, fill=0) then centre crop.  normalise: per-patch
// min-max to [0,1] (all-zero patch if constant, data.py:716-730).  One CTA per patch.
__global__ void __launch_bounds__(256) rotate_crop_kernel(const float* __restrict__ in, int N, int S, int P,
                                                          const double* __restrict__ angle_deg, int mode,
                                                          int normalise, float* __restrict__ out) {
  __shared__ float red_lo[8], red_hi[8];
  const int n = blockIdx.x;
  const float* src = in + (int64_t)n * S * S;
  float* dst = out + (int64_t)n * P * P;
  const int off = (S - P) / 2;                    // TF.center_crop: int(round((S - P) / 2.0)), S - P even
  float m0 = 0.f, m1 = 0.f;
  if (mode == 1) {
    // _get_inverse_affine_matrix(center 0, -angle): [cos r, sin r, 0, -sin r, cos r, 0], r = radians(-angle)
    const double r = -angle_deg[n] * (3.14159265358979323846 / 180.0);
    m0 = (float)cos(r) / (0.5f * (float)S);
    m1 = (float)sin(r) / (0.5f * (float)S);
  }
  const float c = 0.5f * (float)(S - 1);
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
    const int r = i / P + off, col = i - (i / P) * P + off;
    float v;
    if (mode == 1) {
      const float xb = (float)col - c, yb = (float)r - c;
      const float ix = tv_unnormalise(xb, yb, m0, m1, 0.f, S);
      const float iy = tv_unnormalise(xb, yb, -m1, m0, 0.f, S);
      const Tap t = bilinear_zeros(src, S, ix, iy);
      v = t.v * t.mask;                           // + (1 - mask) * fill, fill = 0
    } else {
      v = src[r * S + col];
    }
    dst[i] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  if (!normalise) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { red_lo[threadIdx.x >> 5] = lo; red_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  lo = red_lo[0]; hi = red_hi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) { lo = fminf(lo, red_lo[w]); hi = fmaxf(hi, red_hi[w]); }
  const bool flat = !(hi > lo);
  const float range = hi - lo;
  // each thread re-reads exactly the elements it wrote
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) dst[i] = flat ? 0.f : (dst[i] - lo) / range;
}

}  // namespace livae

extern "C" int livae_augment(const float* in, int N, int S, const float* scale, const int32_t* flags,
                             const int32_t* shift, float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(N >= 0 && S > 0, "augment: bad sizes");
  if (N == 0) return 0;
  LIVAE_CHECK_ARG(in && scale && flags && shift && out && in != out, "augment: null or aliased pointer");
  if (int e = livae::require_sm100()) return e;
  const int bands = N >= 148 * 8 ? 1 : (S >= 64 ? 4 : 1);
  livae::augment_kernel<<<dim3(N, bands), 256, 0, (cudaStream_t)stream>>>(in, N, S, scale, flags, shift, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_rotate_crop(const float* in, int N, int S, int P, const double* angle_deg, int mode,
                                 int normalise, float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(N >= 0 && S > 0 && P > 0 && P <= S && ((S - P) & 1) == 0, "rotate_crop: bad sizes (S - P even)");
  LIVAE_CHECK_ARG(mode == 0 || mode == 1, "rotate_crop: mode must be 0 (crop) or 1 (rotate + crop)");
  if (N == 0) return 0;
  LIVAE_CHECK_ARG(in && out && in != out && (mode == 0 || angle_deg), "rotate_crop: null or aliased pointer");
  if (int e = livae::require_sm100()) return e;
  livae::rotate_crop_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(in, N, S, P, angle_deg, mode, normalise, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

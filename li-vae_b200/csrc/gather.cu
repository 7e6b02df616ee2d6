// a1: peak-centred integer patch gather (bit exact) and per-patch min-max.
// Memory-bound: per patch read P*P*sizeof(src) + write P*P*4 bytes.
#include "common.cuh"

namespace livae {

// One CTA per (patch, 32-row band); each warp copies rows.  Source rows are P
// contiguous elements at an arbitrary (unaligned) x offset, so loads are scalar but
// fully coalesced (a warp reads 32 consecutive elements); stores are float4.
template <typename S>
__global__ void __launch_bounds__(256) patch_gather_kernel(
    const S* __restrict__ images, int n_img, int H, int W, const int32_t* __restrict__ sites,
    int N, int P, float* __restrict__ out) {
  int n = blockIdx.x;
  int img = sites[3 * n + 0], cy = sites[3 * n + 1], cx = sites[3 * n + 2];
  int y0 = cy - P / 2, x0 = cx - P / 2;
  const S* src = images + (int64_t)img * H * W;
  float* dst = out + (int64_t)n * P * P;
  int rows_per_blk = (P + gridDim.y - 1) / gridDim.y;
  int r_beg = blockIdx.y * rows_per_blk;
  int r_end = min(P, r_beg + rows_per_blk);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  bool img_ok = img >= 0 && img < n_img;
  for (int r = r_beg + warp; r < r_end; r += nwarp) {
    int y = y0 + r;
    bool row_ok = img_ok && y >= 0 && y < H;
    const S* srow = src + (int64_t)y * W;
    for (int c = lane; c < P; c += 32) {
      int x = x0 + c;
      float v = 0.f;
      if (row_ok && x >= 0 && x < W) v = (float)srow[x];   // (float)double == numpy .astype(float32), RN
      dst[(int64_t)r * P + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256) patch_minmax_kernel(float* __restrict__ p, int N, int P) {
  __shared__ float smin[32], smax[32];
  float* d = p + (int64_t)blockIdx.x * P * P;
  int n = P * P;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { float v = d[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { smin[w] = lo; smax[w] = hi; }
  __syncthreads();
  lo = smin[0]; hi = smax[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
  // data.py:553-558: (p - min) / (max - min) if max > min else zeros
  if (hi > lo) {
    float den = hi - lo;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = (d[i] - lo) / den;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = 0.f;
  }
}

// a2: sub-pixel gather (AdaptiveLatticeDataset.__getitem__ with transform=None, reference data.py:478-551):
// the patch centred on the FLOAT site (cy, cx) is the bilinear resampling of float32(img), zero outside the image,
// at (cy - P/2 + r, cx - P/2 + c) -- what the reference's integer ROI + sub-pixel TF.affine translate + centre
// crops compute (for even P; its fp32 affine grid differs from this exact form by <= 2e-5).  Coordinates and
// weights in double (a site near 4096 has an fp32 ulp of 5e-4 pixels), one rounding to float at the end.
template <typename S>
__global__ void __launch_bounds__(256) patch_gather_subpixel_kernel(
    const S* __restrict__ images, int n_img, int H, int W, const int32_t* __restrict__ img_idx,
    const double* __restrict__ yx, int N, int P, int roi, float* __restrict__ out) {
  const int n = blockIdx.x;
  const int img = img_idx[n];
  const double cy = yx[2 * n], cx = yx[2 * n + 1];
  const S* src = images + (int64_t)img * H * W;
  float* dst = out + (int64_t)n * P * P;
  const bool img_ok = img >= 0 && img < n_img;
  // roi > 0: the reference's integer ROI window of `roi` pixels around round-half-even(c) (data.py:496-511):
  // nothing outside it is read, which matters when the crop is as large as the window (P + 2*padding)
  int wy0 = 0, wy1 = H, wx0 = 0, wx1 = W;
  if (roi > 0) {
    const int yi = (int)rint(cy), xi = (int)rint(cx);
    wy0 = max(0, yi - roi / 2); wy1 = min(H, yi - roi / 2 + roi);
    wx0 = max(0, xi - roi / 2); wx1 = min(W, xi - roi / 2 + roi);
  }
  const int rows_per_blk = (P + gridDim.y - 1) / gridDim.y;
  const int r_beg = blockIdx.y * rows_per_blk, r_end = min(P, r_beg + rows_per_blk);
  for (int i = r_beg * P + threadIdx.x; i < r_end * P; i += blockDim.x) {
    const int r = i / P, c = i - r * P;
    const double ys = cy - (double)(P / 2) + r, xs = cx - (double)(P / 2) + c;
    const double fy0 = floor(ys), fx0 = floor(xs);
    const int y0 = (int)fy0, x0 = (int)fx0;
    const double fy = ys - fy0, fx = xs - fx0;
    double acc = 0.0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        if (img_ok && yy >= wy0 && yy < wy1 && xx >= wx0 && xx < wx1)
          acc += (dy ? fy : 1.0 - fy) * (dx ? fx : 1.0 - fx) * (double)(float)src[(int64_t)yy * W + xx];
      }
    dst[i] = (float)acc;
  }
}

template <typename S>
int patch_gather_subpixel(const S* images, int n_img, int H, int W, const int32_t* img_idx, const double* yx, int N,
                          int P, int roi, float* out, cudaStream_t st) {
  LIVAE_CHECK_ARG(N >= 0 && P > 0 && (P & 1) == 0 && H > 0 && W > 0 && n_img > 0, "patch_gather_subpixel: bad sizes (P even)");
  LIVAE_CHECK_ARG(roi == 0 || (roi >= P && (roi & 1) == 0), "patch_gather_subpixel: roi must be 0 or even and >= P");
  if (N == 0) return 0;
  LIVAE_CHECK_ARG(images && img_idx && yx && out, "patch_gather_subpixel: null pointer");
  if (int e = require_sm100()) return e;
  const int bands = N >= 148 * 8 ? 1 : (P >= 64 ? 4 : 1);
  patch_gather_subpixel_kernel<S><<<dim3(N, bands), 256, 0, st>>>(images, n_img, H, W, img_idx, yx, N, P, roi, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

template <typename S>
int patch_gather(const S* images, int n_img, int H, int W, const int32_t* sites, int N, int P,
                 float* out, cudaStream_t st) {
  LIVAE_CHECK_ARG(N >= 0 && P > 0 && H > 0 && W > 0 && n_img > 0, "patch_gather: bad sizes");
  if (N == 0) return 0;   // empty site list: nothing to do (pointers may be null)
  LIVAE_CHECK_ARG(images && sites && out, "patch_gather: null pointer");
  if (int e = require_sm100()) return e;
  // enough CTAs to fill 148 SMs several times over even for small N
  int bands = N >= 148 * 8 ? 1 : (P >= 64 ? 4 : 1);
  dim3 grid(N, bands);
  patch_gather_kernel<S><<<grid, 256, 0, st>>>(images, n_img, H, W, sites, N, P, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}
}  // namespace livae

extern "C" int livae_patch_gather_f32(const float* images, int n_img, int H, int W, const int32_t* sites,
                                      int N, int P, float* out, livae_stream_t stream) {
  return livae::patch_gather<float>(images, n_img, H, W, sites, N, P, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_gather_f64(const double* images, int n_img, int H, int W, const int32_t* sites,
                                      int N, int P, float* out, livae_stream_t stream) {
  return livae::patch_gather<double>(images, n_img, H, W, sites, N, P, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_gather_subpixel_f32(const float* images, int n_img, int H, int W, const int32_t* img_idx,
                                               const double* yx, int N, int P, float* out, livae_stream_t stream) {
  return livae::patch_gather_subpixel<float>(images, n_img, H, W, img_idx, yx, N, P, 0, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_gather_subpixel_f64(const double* images, int n_img, int H, int W, const int32_t* img_idx,
                                               const double* yx, int N, int P, float* out, livae_stream_t stream) {
  return livae::patch_gather_subpixel<double>(images, n_img, H, W, img_idx, yx, N, P, 0, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_gather_roi_f32(const float* images, int n_img, int H, int W, const int32_t* img_idx,
                                          const double* yx, int N, int S, int roi, float* out, livae_stream_t stream) {
  return livae::patch_gather_subpixel<float>(images, n_img, H, W, img_idx, yx, N, S, roi, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_gather_roi_f64(const double* images, int n_img, int H, int W, const int32_t* img_idx,
                                          const double* yx, int N, int S, int roi, float* out, livae_stream_t stream) {
  return livae::patch_gather_subpixel<double>(images, n_img, H, W, img_idx, yx, N, S, roi, out, (cudaStream_t)stream);
}
extern "C" int livae_patch_minmax(float* patches, int N, int P, livae_stream_t stream) {
  LIVAE_CHECK_ARG(N >= 0 && P > 0, "patch_minmax: bad args");
  if (N == 0) return 0;
  LIVAE_CHECK_ARG(patches, "patch_minmax: null pointer");
  if (int e = livae::require_sm100()) return e;
  livae::patch_minmax_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(patches, N, P);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

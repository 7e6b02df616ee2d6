// a6: fused rotate + bilinear sample == F.affine_grid + F.grid_sample(bilinear, reflection,
// align_corners=False) at reference model.py:254-258, model.py:467-470, train.py:675-677.
//
// Memory-bound.  Algorithmic bytes per [H,W] fp32 image: forward read H*W*4 + write H*W*4;
// backward read gout + read img (+ write gimg).  No [B,H,W,2] grid tensor is ever
// materialised: the source coordinate of every output pixel is recomputed in registers
// from the per-sample (cos, sin).
//
// Forward: one CTA per (sample, channel) image.  The source image is staged ONCE into shared
// memory with coalesced float4 loads (64 KB at 128x128, 3 CTAs resident per SM), the four
// bilinear taps are shared-memory gathers, the output is written with coalesced float4 stores.
//
// Backward: one CTA per sample.  grad_input is a scatter; instead of global atomics the CTA
// accumulates into a shared-memory tile (shared atomics, no HBM traffic) and writes the tile out
// once with float4 stores, so HBM sees exactly one coalesced write of grad_input.  The per-sample
// dL/d(cos), dL/d(sin) are reduced with warp shuffles + one shared exchange.
#include "common.cuh"

namespace livae {

struct SrcCoord {
  float v;     // source coordinate after reflect + clip
  float mult;  // d(v)/d(normalised grid coordinate)
};

// ATen grid_sampler: unnormalise (align_corners=False), reflect about [-0.5, n-0.5], clip to
// [0, n-1] (gradient zero where clipped, clip_coordinates_set_grad uses <= / >=).
__device__ __forceinline__ SrcCoord src_coord(float g, int n) {
  float fn = (float)n;
  float u = ((g + 1.f) * fn - 1.f) * 0.5f;
  float mult = fn * 0.5f;
  float v = u + 0.5f;
  float r;
  if (v >= 0.f && v < fn) {
    // inside the image (all but the corner pixels of a rotated patch): flips = 0 and fmodf(v, fn) = v
    // exactly, so this is bit-identical to the general path below without its division and fmodf
    r = v - 0.5f;
  } else {
    if (v < 0.f) { v = -v; mult = -mult; }
    float flips = floorf(v / fn);
    float extra = fmodf(v, fn);
    if (((int)flips & 1) == 0) {
      r = extra - 0.5f;
    } else {
      r = fn - extra - 0.5f;
      mult = -mult;
    }
  }
  if (r <= 0.f) { r = 0.f; mult = 0.f; }
  else if (r >= fn - 1.f) { r = fn - 1.f; mult = 0.f; }
  SrcCoord o; o.v = r; o.mult = mult;
  return o;
}

template <bool kSmem>
__global__ void __launch_bounds__(256) rot_sample_fwd_kernel(
    const float* __restrict__ img, const float* __restrict__ cs, float sgn, int C, int H, int W,
    float* __restrict__ out) {
  extern __shared__ __align__(16) float s_img[];
  int bc = blockIdx.x;
  int b = bc / C;
  const float* src = img + (int64_t)bc * H * W;
  float* dst = out + (int64_t)bc * H * W;
  int n = H * W;
  if (kSmem) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(s_img);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) s_img[i] = __ldg(src + i);
    __syncthreads();
  }
  const float* tap = kSmem ? s_img : src;
  float c = cs[2 * b], s = sgn * cs[2 * b + 1];
  float invW = 1.f / (float)W, invH = 1.f / (float)H;
  int wq = (W + 3) / 4;
  for (int q = threadIdx.x; q < H * wq; q += blockDim.x) {
    int i = q / wq, j0 = (q - i * wq) * 4;
    float ys = (2.f * i + 1.f) * invH - 1.f;
    float o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int j = j0 + t;
      float xs = (2.f * j + 1.f) * invW - 1.f;
      SrcCoord cx = src_coord(c * xs - s * ys, W);
      SrcCoord cy = src_coord(s * xs + c * ys, H);
      float fx0 = floorf(cx.v), fy0 = floorf(cy.v);
      int x0 = (int)fx0, y0 = (int)fy0;
      float fx = cx.v - fx0, fy = cy.v - fy0;
      // coordinates are clipped to [0,n-1]: only the +1 taps can fall outside
      bool x1ok = x0 + 1 < W, y1ok = y0 + 1 < H;
      int x1 = x1ok ? x0 + 1 : x0, y1 = y1ok ? y0 + 1 : y0;
      float v00 = tap[y0 * W + x0];
      float v01 = x1ok ? tap[y0 * W + x1] : 0.f;
      float v10 = y1ok ? tap[y1 * W + x0] : 0.f;
      float v11 = (x1ok && y1ok) ? tap[y1 * W + x1] : 0.f;
      // same association as ATen: nw*(1-fx)(1-fy) + ne*fx(1-fy) + sw*(1-fx)fy + se*fx*fy
      o[t] = v00 * ((1.f - fx) * (1.f - fy)) + v01 * (fx * (1.f - fy)) +
             v10 * ((1.f - fx) * fy) + v11 * (fx * fy);
    }
    if ((W & 3) == 0) {
      *reinterpret_cast<float4*>(dst + i * W + j0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int t = 0; t < 4 && j0 + t < W; ++t) dst[i * W + j0 + t] = o[t];
    }
  }
}

// kSmemGrad: accumulate grad_input in a shared tile; otherwise (image too large for shared
// memory) fall back to global atomics on a pre-zeroed buffer.  kSmemImg: the source image is staged in
// shared memory as in the forward kernel (the four taps of a pixel are gathers; from global memory they
// miss L1, which the gradient tile has squeezed to a few KB, on every pixel).
template <bool kSmemGrad, bool kSmemImg>
__global__ void __launch_bounds__(512) rot_sample_bwd_kernel(
    const float* __restrict__ img, const float* __restrict__ cs, float sgn,
    const float* __restrict__ gout, int C, int H, int W, float* __restrict__ gimg,
    float* __restrict__ gcs) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_img = s_dyn;                                   // [H*W] when kSmemImg
  float* s_g = s_dyn + (kSmemImg ? H * W : 0);            // [H*W] when kSmemGrad && gimg
  __shared__ float red[2][32];
  int b = blockIdx.x;
  int n = H * W;
  float c = cs[2 * b], s = sgn * cs[2 * b + 1];
  float invW = 1.f / (float)W, invH = 1.f / (float)H;
  float acc_c = 0.f, acc_s = 0.f;
  for (int ch = 0; ch < C; ++ch) {
    const float* src = img + ((int64_t)b * C + ch) * n;
    const float* go = gout + ((int64_t)b * C + ch) * n;
    float* gi = gimg ? gimg + ((int64_t)b * C + ch) * n : nullptr;
    if (kSmemImg) {
      __syncthreads();
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(s_img);
      for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
      for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) s_img[i] = __ldg(src + i);
    }
    if (kSmemGrad && gi)
      for (int i = threadIdx.x; i < n; i += blockDim.x) s_g[i] = 0.f;
    if (kSmemImg || (kSmemGrad && gi)) __syncthreads();
    const float* tap = kSmemImg ? s_img : src;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
      int i = p / W, j = p - i * W;
      float ys = (2.f * i + 1.f) * invH - 1.f;
      float xs = (2.f * j + 1.f) * invW - 1.f;
      SrcCoord cx = src_coord(c * xs - s * ys, W);
      SrcCoord cy = src_coord(s * xs + c * ys, H);
      float fx0 = floorf(cx.v), fy0 = floorf(cy.v);
      int x0 = (int)fx0, y0 = (int)fy0;
      float fx = cx.v - fx0, fy = cy.v - fy0;
      bool x1ok = x0 + 1 < W, y1ok = y0 + 1 < H;
      int x1 = x1ok ? x0 + 1 : x0, y1 = y1ok ? y0 + 1 : y0;
      float g = __ldg(go + p);
      float v00 = tap[y0 * W + x0];
      float v01 = x1ok ? tap[y0 * W + x1] : 0.f;
      float v10 = y1ok ? tap[y1 * W + x0] : 0.f;
      float v11 = (x1ok && y1ok) ? tap[y1 * W + x1] : 0.f;
      if (gi) {
        float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy);
        float w10 = (1.f - fx) * fy, w11 = fx * fy;
        float* t = kSmemGrad ? s_g : gi;
        atomicAdd(t + y0 * W + x0, w00 * g);
        if (x1ok) atomicAdd(t + y0 * W + x1, w01 * g);
        if (y1ok) atomicAdd(t + y1 * W + x0, w10 * g);
        if (x1ok && y1ok) atomicAdd(t + y1 * W + x1, w11 * g);
      }
      // d(out)/d(ix), d(out)/d(iy)
      float gix = (-(1.f - fy) * v00 + (1.f - fy) * v01 - fy * v10 + fy * v11) * g;
      float giy = (-(1.f - fx) * v00 - fx * v01 + (1.f - fx) * v10 + fx * v11) * g;
      float ggx = gix * cx.mult, ggy = giy * cy.mult;
      // affine_grid backward: base_grid^T @ grad_grid for [[c,-s],[s,c]]
      acc_c += ggx * xs + ggy * ys;
      acc_s += -ggx * ys + ggy * xs;
    }
    if (kSmemGrad && gi) {
      __syncthreads();
      if ((n & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(gi);
        const float4* s4 = reinterpret_cast<const float4*>(s_g);
        for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = s4[i];
      } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) gi[i] = s_g[i];
      }
      __syncthreads();
    }
  }
  acc_c = warp_sum(acc_c);
  acc_s = warp_sum(acc_s);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = acc_c; red[1][w] = acc_s; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    float a = lane < nw ? red[0][lane] : 0.f;
    float d = lane < nw ? red[1][lane] : 0.f;
    a = warp_sum(a);
    d = warp_sum(d);
    if (lane == 0 && gcs) { gcs[2 * b] = a; gcs[2 * b + 1] = sgn * d; }
  }
}

static constexpr int kMaxSmemImage = 200 * 1024;  // bytes of one staged image / gradient tile

}  // namespace livae

extern "C" int livae_rot_sample_fwd(const float* img, const float* cs, float sgn, int B, int C, int H,
                                    int W, float* out, livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0, "rot_sample_fwd: bad sizes");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && cs && out, "rot_sample_fwd: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  size_t bytes = (size_t)H * W * sizeof(float);
  if (bytes <= (size_t)kMaxSmemImage && ((uintptr_t)img & 15) == 0) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(rot_sample_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kMaxSmemImage);
      attr_done = true;
    }
    rot_sample_fwd_kernel<true><<<B * C, 256, bytes, st>>>(img, cs, sgn, C, H, W, out);
  } else {
    rot_sample_fwd_kernel<false><<<B * C, 256, 0, st>>>(img, cs, sgn, C, H, W, out);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_rot_sample_bwd(const float* img, const float* cs, float sgn, const float* gout,
                                    int B, int C, int H, int W, float* gimg, float* gcs,
                                    livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0, "rot_sample_bwd: bad sizes");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && cs && gout, "rot_sample_bwd: null pointer");
  LIVAE_CHECK_ARG(gimg || gcs, "rot_sample_bwd: nothing to compute");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  size_t bytes = (size_t)H * W * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(rot_sample_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemImage);
    cudaFuncSetAttribute(rot_sample_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemImage);
    attr_done = true;
  }
  const bool aligned = ((uintptr_t)img & 15) == 0;
  if (!gimg) {
    if (aligned && bytes <= (size_t)kMaxSmemImage)
      rot_sample_bwd_kernel<true, true><<<B, 512, bytes, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
    else
      rot_sample_bwd_kernel<true, false><<<B, 512, 0, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  } else if (aligned && 2 * bytes <= (size_t)kMaxSmemImage) {
    rot_sample_bwd_kernel<true, true><<<B, 512, 2 * bytes, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  } else if (bytes <= (size_t)kMaxSmemImage) {
    rot_sample_bwd_kernel<true, false><<<B, 512, bytes, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  } else {
    cudaError_t e = cudaMemsetAsync(gimg, 0, (size_t)B * C * bytes, st);
    if (e != cudaSuccess) { set_error("rot_sample_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    rot_sample_bwd_kernel<false, false><<<B, 512, 0, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

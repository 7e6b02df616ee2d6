// Box-filter SSIM of the per-step metric block (reference train.py:606-667: five avg_pool2d(win, stride 1,
// pad win/2, zero padding counted) passes + elementwise map + mean), fused into one pass over the two images:
// a CTA stages a strip of rows (+ halo) of both images in shared memory, forms the five horizontal window sums
// (a, b, a*a, b*b, a*b), then the vertical sums, the SSIM map and its partial sum.  HBM traffic: both images once.
#include "common.cuh"

namespace livae {

static constexpr int kSsimRows = 16;     // output rows per CTA

__global__ void __launch_bounds__(256) ssim_box_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                       int win, float c1, float c2, float* __restrict__ partial) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int half = win >> 1;
  const int R = kSsimRows + 2 * half;                  // staged rows
  float* sa = sm;                                      // [R][W]
  float* sb = sa + R * W;
  float* sh = sb + R * W;                              // [5][R][W] horizontal window sums
  const int plane = blockIdx.y, r0 = blockIdx.x * kSsimRows;
  const float* pa = a + (int64_t)plane * H * W;
  const float* pb = b + (int64_t)plane * H * W;
  for (int i = threadIdx.x; i < R * W; i += blockDim.x) {
    const int r = i / W, x = i - r * W, y = r0 - half + r;
    const bool ok = y >= 0 && y < H;
    sa[i] = ok ? __ldg(pa + y * W + x) : 0.f;
    sb[i] = ok ? __ldg(pb + y * W + x) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * W; i += blockDim.x) {
    const int r = i / W, x = i - r * W;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
    const int x0 = max(0, x - half), x1 = min(W - 1, x + half);
    for (int k = x0; k <= x1; ++k) {
      const float u = sa[r * W + k], v = sb[r * W + k];
      s0 += u; s1 += v; s2 = fmaf(u, u, s2); s3 = fmaf(v, v, s3); s4 = fmaf(u, v, s4);
    }
    sh[i] = s0; sh[R * W + i] = s1; sh[2 * R * W + i] = s2; sh[3 * R * W + i] = s3; sh[4 * R * W + i] = s4;
  }
  __syncthreads();
  const float inv = 1.f / (float)(win * win);          // count_include_pad: always win*win
  float acc = 0.f;
  for (int i = threadIdx.x; i < kSsimRows * W; i += blockDim.x) {
    const int r = i / W, x = i - r * W;
    if (r0 + r >= H) continue;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < win; ++k) {
#pragma unroll
      for (int q = 0; q < 5; ++q) s[q] += sh[q * R * W + (r + k) * W + x];
    }
    const float mu1 = s[0] * inv, mu2 = s[1] * inv;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sig1 = s[2] * inv - mu1_sq, sig2 = s[3] * inv - mu2_sq, sig12 = s[4] * inv - mu12;
    acc += ((2.f * mu12 + c1) * (2.f * sig12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sig1 + sig2 + c2));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// deterministic finish: one CTA sums the partials in a fixed order
__global__ void __launch_bounds__(256) ssim_finish_kernel(const float* __restrict__ partial, int64_t n, float scale,
                                                          float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

}  // namespace livae

using namespace livae;

extern "C" int64_t livae_ssim_box_ws_floats(int64_t planes, int H) { return planes * ((H + kSsimRows - 1) / kSsimRows); }

// out[0] = mean SSIM map of a, b: fp32 [planes, H, W] (planes = B*C), odd window `win`; ws: livae_ssim_box_ws_floats
extern "C" int livae_ssim_box(const float* a, const float* b, int64_t planes, int H, int W, int win, float c1, float c2,
                              float* ws, float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && win > 0 && (win & 1) == 1, "ssim_box: bad sizes (odd window)");
  LIVAE_CHECK_ARG(out, "ssim_box: null out");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 0) { cudaMemsetAsync(out, 0, sizeof(float), st); return 0; }
  LIVAE_CHECK_ARG(a && b && ws, "ssim_box: null pointer");
  LIVAE_CHECK_ARG(planes <= 65535, "ssim_box: too many planes (%lld)", (long long)planes);
  const int R = kSsimRows + 2 * (win / 2);
  const size_t smem = (size_t)7 * R * W * sizeof(float);
  LIVAE_CHECK_ARG(smem <= 200 * 1024, "ssim_box: image too wide for the shared-memory strip (W = %d, window %d)", W, win);
  static bool attr_done = false;
  if (!attr_done) { cudaFuncSetAttribute(ssim_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
  const int strips = (H + kSsimRows - 1) / kSsimRows;
  ssim_box_kernel<<<dim3(strips, (unsigned)planes), 256, smem, st>>>(a, b, H, W, win, c1, c2, ws);
  LIVAE_CUDA_LAUNCH_CHECK();
  ssim_finish_kernel<<<1, 256, 0, st>>>(ws, planes * strips, 1.f / ((float)planes * H * W), out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// Box-filter SSIM of the per-step metric block (reference train.py:606-667: five avg_pool2d(win, stride 1,
// pad win/2, zero padding counted) passes + elementwise map + mean), fused into one pass over the two images:
// a CTA stages a strip of rows (+ halo) of both images in shared memory, forms the five horizontal window sums
// (a, b, a*a, b*b, a*b), then the vertical sums, the SSIM map and its partial sum.  HBM traffic: both images once.
#include "common.cuh"

namespace livae {

static constexpr int kSsimRows = 16;     // output rows per CTA

// WIN > 0: compile-time window (the reference's 11): rows are staged with `half` zero columns on either side so the
// window loops have no bounds and unroll fully -- with run-time bounds every iteration waited out its own
// shared-memory latency (4 warps per scheduler at this footprint), 1.17 ms per call against 0.27 GB of traffic.
// Summation order is unchanged (left to right, top to bottom, zeros included), so results are bit-identical.
template <int WIN>
__global__ void __launch_bounds__(256) ssim_box_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W,
                                                       int win_rt, float c1, float c2, float* __restrict__ partial) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int win = WIN > 0 ? WIN : win_rt;
  const int half = win >> 1;
  const int R = kSsimRows + 2 * half;                  // staged rows
  const int Wp = W + 2 * half;                         // staged row pitch (zero columns left and right)
  float* sa = sm;                                      // [R][Wp]
  float* sb = sa + R * Wp;
  float* sh = sb + R * Wp;                             // [5][R][W] horizontal window sums
  const int plane = blockIdx.y, r0 = blockIdx.x * kSsimRows;
  const float* pa = a + (int64_t)plane * H * W;
  const float* pb = b + (int64_t)plane * H * W;
  for (int i = threadIdx.x; i < R * Wp; i += blockDim.x) {
    const int r = i / Wp, xp = i - r * Wp, x = xp - half, y = r0 - half + r;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    sa[i] = ok ? __ldg(pa + y * W + x) : 0.f;
    sb[i] = ok ? __ldg(pb + y * W + x) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * W; i += blockDim.x) {
    const int r = i / W, x = i - r * W;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
    const float* ra = sa + r * Wp + x;                 // window [x - half, x + half] = staged columns x .. x + 2 half
    const float* rb = sb + r * Wp + x;
    if (WIN > 0) {
#pragma unroll
      for (int k = 0; k < (WIN > 0 ? WIN : 1); ++k) {
        const float u = ra[k], v = rb[k];
        s0 += u; s1 += v; s2 = fmaf(u, u, s2); s3 = fmaf(v, v, s3); s4 = fmaf(u, v, s4);
      }
    } else {
      for (int k = 0; k < win; ++k) {
        const float u = ra[k], v = rb[k];
        s0 += u; s1 += v; s2 = fmaf(u, u, s2); s3 = fmaf(v, v, s3); s4 = fmaf(u, v, s4);
      }
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
    if (WIN > 0) {
#pragma unroll
      for (int k = 0; k < (WIN > 0 ? WIN : 1); ++k) {
#pragma unroll
        for (int q = 0; q < 5; ++q) s[q] += sh[q * R * W + (r + k) * W + x];
      }
    } else {
      for (int k = 0; k < win; ++k) {
#pragma unroll
        for (int q = 0; q < 5; ++q) s[q] += sh[q * R * W + (r + k) * W + x];
      }
    }
    const float mu1 = s[0] * inv, mu2 = s[1] * inv;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sig1 = s[2] * inv - mu1_sq, sig2 = s[3] * inv - mu2_sq, sig12 = s[4] * inv - mu12;
    acc += ((2.f * mu12 + c1) * (2.f * sig12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sig1 + sig2 + c2));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// Window 11, W % 4 == 0: the same sums in the same order, but every thread produces 4 (horizontal pass) or 2 x 4
// (vertical pass) neighbouring outputs from float4 shared-memory loads held in registers: 6x fewer shared-memory
// instructions than one scalar load per tap (the scalar form was bound by the LDS issue rate).
// Staged rows have 8 zero columns on the left (float4-aligned; the window of x starts at staged column x + 3)
// and 8 on the right.
__global__ void __launch_bounds__(256) ssim_box11_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, int H,
                                                             int W, float c1, float c2, float* __restrict__ partial) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[32];
  constexpr int WIN = 11, HALF = 5, PADL = 8;
  const int R = kSsimRows + 2 * HALF;
  const int Wp = W + 2 * PADL;
  float* sa = sm;                                      // [R][Wp]
  float* sb = sa + R * Wp;
  float* sh = sb + R * Wp;                             // [5][R][W]
  const int plane = blockIdx.y, r0 = blockIdx.x * kSsimRows;
  const float* pa = a + (int64_t)plane * H * W;
  const float* pb = b + (int64_t)plane * H * W;
  const int wq = Wp >> 2;
  for (int i = threadIdx.x; i < R * wq; i += blockDim.x) {
    const int r = i / wq, x = ((i - r * wq) << 2) - PADL, y = r0 - HALF + r;
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (y >= 0 && y < H && x >= 0 && x < W) {          // W % 4 == 0: a chunk is entirely inside or outside
      va = __ldg(reinterpret_cast<const float4*>(pa + y * W + x));
      vb = __ldg(reinterpret_cast<const float4*>(pb + y * W + x));
    }
    *reinterpret_cast<float4*>(sa + r * Wp + x + PADL) = va;
    *reinterpret_cast<float4*>(sb + r * Wp + x + PADL) = vb;
  }
  __syncthreads();
  const int xq = W >> 2;
  for (int i = threadIdx.x; i < R * xq; i += blockDim.x) {
    const int r = i / xq, x = (i - r * xq) << 2;
    float u[20], v[20];                                // staged columns x .. x + 19; output x + j uses 3 + j .. 13 + j
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float4 ta = *reinterpret_cast<const float4*>(sa + r * Wp + x + 4 * c);
      const float4 tb = *reinterpret_cast<const float4*>(sb + r * Wp + x + 4 * c);
      u[4 * c] = ta.x; u[4 * c + 1] = ta.y; u[4 * c + 2] = ta.z; u[4 * c + 3] = ta.w;
      v[4 * c] = tb.x; v[4 * c + 1] = tb.y; v[4 * c + 2] = tb.z; v[4 * c + 3] = tb.w;
    }
    float o[5][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float uu = u[3 + j + k], vv = v[3 + j + k];
        s0 += uu; s1 += vv; s2 = fmaf(uu, uu, s2); s3 = fmaf(vv, vv, s3); s4 = fmaf(uu, vv, s4);
      }
      o[0][j] = s0; o[1][j] = s1; o[2][j] = s2; o[3][j] = s3; o[4][j] = s4;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q)
      *reinterpret_cast<float4*>(sh + q * R * W + r * W + x) = make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
  }
  __syncthreads();
  const float inv = 1.f / (float)(WIN * WIN);
  float acc = 0.f;
  for (int i = threadIdx.x; i < (kSsimRows / 2) * xq; i += blockDim.x) {
    const int rp = i / xq, x = (i - rp * xq) << 2, r = 2 * rp;        // output rows r, r + 1 of the strip
    float s[2][5][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[t][q][j] = 0.f;
#pragma unroll
    for (int k = 0; k < WIN + 1; ++k) {
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 h4 = *reinterpret_cast<const float4*>(sh + q * R * W + (r + k) * W + x);
        const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k < WIN) s[0][q][j] += hv[j];            // rows r .. r + 10
          if (k > 0) s[1][q][j] += hv[j];              // rows r + 1 .. r + 11
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (r0 + r + t >= H) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float mu1 = s[t][0][j] * inv, mu2 = s[t][1][j] * inv;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float sig1 = s[t][2][j] * inv - mu1_sq, sig2 = s[t][3][j] * inv - mu2_sq, sig12 = s[t][4][j] * inv - mu12;
        acc += ((2.f * mu12 + c1) * (2.f * sig12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sig1 + sig2 + c2));
      }
    }
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
  const size_t smem = ((size_t)2 * R * (W + 2 * (win / 2)) + (size_t)5 * R * W) * sizeof(float);
  LIVAE_CHECK_ARG(smem <= 200 * 1024, "ssim_box: image too wide for the shared-memory strip (W = %d, window %d)", W, win);
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(ssim_box_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ssim_box_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int strips = (H + kSsimRows - 1) / kSsimRows;
  const size_t smem_vec = ((size_t)2 * R * (W + 16) + (size_t)5 * R * W) * sizeof(float);
  if (win == 11 && (W & 3) == 0 && smem_vec <= 200 * 1024 && ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0) {
    static OncePerDevice vec_attr;
    if (vec_attr.first()) { cudaFuncSetAttribute(ssim_box11_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);}
    ssim_box11_vec_kernel<<<dim3(strips, (unsigned)planes), 256, smem_vec, st>>>(a, b, H, W, c1, c2, ws);
  } else if (win == 11) ssim_box_kernel<11><<<dim3(strips, (unsigned)planes), 256, smem, st>>>(a, b, H, W, win, c1, c2, ws);
  else ssim_box_kernel<0><<<dim3(strips, (unsigned)planes), 256, smem, st>>>(a, b, H, W, win, c1, c2, ws);
  LIVAE_CUDA_LAUNCH_CHECK();
  ssim_finish_kernel<<<1, 256, 0, st>>>(ws, planes * strips, 1.f / ((float)planes * H * W), out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

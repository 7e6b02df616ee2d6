// a8 / a12 / a13 / a4-tail: small memory- or latency-bound pieces of the step.
//   * rotation head: F.normalize(vec, eps=1e-6) -> (cos, sin), theta = atan2   (model.py:245-261)
//   * get_rotation_matrix(theta) -> (cos, sin)                                 (model.py:220-235)
//   * reparameterisation z = mu + eps * exp(0.5 logvar)                        (model.py:426-440)
//   * fused ELBO reductions  sum((r-x)^2)  and  sum(-0.5 (1 + lv - mu^2 - e^lv))  in ONE launch
//     (loss.py:116-119, 162-169; train.py:391-393) and their fused backward
//   * cycle-consistency loss mean(1 - cos(th_r - th + angle))                  (loss.py:52-94)
// ELBO forward: algorithmic bytes = 2 * n_pix * 4 (read recon, x); backward = 2 reads + 1 write
// (+1 write when the target also needs a gradient).  Reductions are two-stage and deterministic:
// per-CTA partials in a scratch buffer, the last CTA to finish (atomic ticket) sums them in order.
#include "common.cuh"

namespace livae {

__global__ void stn_head_fwd_kernel(const float* __restrict__ vec, int B, float* __restrict__ cs,
                                    float* __restrict__ theta) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a = vec[2 * b], d = vec[2 * b + 1];
  float n = fmaxf(sqrtf(a * a + d * d), 1e-6f);
  float c = a / n, s = d / n;
  cs[2 * b] = c;
  cs[2 * b + 1] = s;
  if (theta) theta[b] = atan2f(s, c);
}

__global__ void stn_head_bwd_kernel(const float* __restrict__ vec, const float* __restrict__ gcs,
                                    const float* __restrict__ gtheta, int B, float* __restrict__ gvec) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a = vec[2 * b], d = vec[2 * b + 1];
  float nr = sqrtf(a * a + d * d);
  float n = fmaxf(nr, 1e-6f);
  float c = a / n, s = d / n;
  float gc = gcs ? gcs[2 * b] : 0.f, gs = gcs ? gcs[2 * b + 1] : 0.f;
  if (gtheta) {
    float r2 = c * c + s * s;  // atan2(s, c): d/dc = -s/r2, d/ds = c/r2
    float gt = gtheta[b];
    gc += -s / r2 * gt;
    gs += c / r2 * gt;
  }
  if (nr > 1e-6f) {
    float dot = c * gc + s * gs;
    gvec[2 * b] = (gc - c * dot) / n;
    gvec[2 * b + 1] = (gs - s * dot) / n;
  } else {  // clamp_min branch of F.normalize: denominator is the constant eps
    gvec[2 * b] = gc / n;
    gvec[2 * b + 1] = gs / n;
  }
}

// STN tail in one launch each way: Linear(32 -> 2) (model.py:213) + F.normalize + atan2 (model.py:245-261).
// One warp per sample, lane = input feature.  Replaces a GEMM launch on a [B,32]x[32,2] problem plus the head.
__global__ void __launch_bounds__(256) stn_tail_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ w9,
                                                           const float* __restrict__ b9, int B,
                                                           float* __restrict__ vec, float* __restrict__ cs,
                                                           float* __restrict__ theta) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  const float w0 = w9[lane], w1 = w9[32 + lane];
  for (int b = warp; b < B; b += nwarp) {
    const float f = f1[b * 32 + lane];
    const float a = warp_sum(f * w0) + b9[0], d = warp_sum(f * w1) + b9[1];
    if (lane == 0) {
      vec[2 * b] = a; vec[2 * b + 1] = d;
      const float n = fmaxf(sqrtf(a * a + d * d), 1e-6f);
      const float c = a / n, s = d / n;
      cs[2 * b] = c; cs[2 * b + 1] = s;
      if (theta) theta[b] = atan2f(s, c);
    }
  }
}

// backward of the above: gvec from (gcs, gtheta) as stn_head_bwd_kernel, then gw9 = gvec^T f1, gb9 = sum gvec
// (atomics on zeroed outputs, one per CTA and element) and gf1 = (gvec w9) * (f1 > 0) written as bf16 for the
// tensor-core fc1 backward.
__global__ void __launch_bounds__(256) stn_tail_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ w9,
                                                           const float* __restrict__ vec,
                                                           const float* __restrict__ gcs,
                                                           const float* __restrict__ gtheta, int B,
                                                           float* __restrict__ gw9, float* __restrict__ gb9,
                                                           __nv_bfloat16* __restrict__ gf1b) {
  __shared__ float sw[8][66];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  const float w0 = w9[lane], w1 = w9[32 + lane];
  float aw0 = 0.f, aw1 = 0.f, ab0 = 0.f, ab1 = 0.f;
  for (int b = warp; b < B; b += nwarp) {
    const float a = vec[2 * b], d = vec[2 * b + 1];
    const float nr = sqrtf(a * a + d * d);
    const float n = fmaxf(nr, 1e-6f);
    const float c = a / n, s = d / n;
    float gc = gcs ? gcs[2 * b] : 0.f, gs = gcs ? gcs[2 * b + 1] : 0.f;
    if (gtheta) {
      const float r2 = c * c + s * s, gt = gtheta[b];
      gc += -s / r2 * gt;
      gs += c / r2 * gt;
    }
    float g0, g1;
    if (nr > 1e-6f) {
      const float dot = c * gc + s * gs;
      g0 = (gc - c * dot) / n; g1 = (gs - s * dot) / n;
    } else {
      g0 = gc / n; g1 = gs / n;
    }
    const float f = f1[b * 32 + lane];
    aw0 = fmaf(g0, f, aw0); aw1 = fmaf(g1, f, aw1);
    ab0 += g0; ab1 += g1;
    gf1b[b * 32 + lane] = __float2bfloat16_rn(f > 0.f ? fmaf(g0, w0, g1 * w1) : 0.f);
  }
  sw[wid][lane] = aw0; sw[wid][32 + lane] = aw1;
  if (lane == 0) { sw[wid][64] = ab0; sw[wid][65] = ab1; }
  __syncthreads();
  if (threadIdx.x < 66) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sw[w][threadIdx.x];
    atomicAdd(threadIdx.x < 64 ? gw9 + threadIdx.x : gb9 + (threadIdx.x - 64), t);
  }
}

__global__ void angle_to_cs_kernel(const float* __restrict__ theta, int B, float* __restrict__ cs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s, c;
  sincosf(theta[b], &s, &c);
  cs[2 * b] = c;
  cs[2 * b + 1] = s;
}

__global__ void angle_to_cs_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ gcs,
                                       int B, float* __restrict__ gtheta) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s, c;
  sincosf(theta[b], &s, &c);
  gtheta[b] = -s * gcs[2 * b] + c * gcs[2 * b + 1];
}

__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ eps, int n, float* __restrict__ z) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = mu[i] + eps[i] * expf(0.5f * lv[i]);
}

__global__ void reparam_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ lv,
                                   const float* __restrict__ eps, int n, float* __restrict__ gmu,
                                   float* __restrict__ glv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float g = gz[i];
    gmu[i] = g;
    glv[i] = g * eps[i] * (0.5f * expf(0.5f * lv[i]));
  }
}

static constexpr int kElboMaxBlocks = 148 * 8;

// sums[0] = sum((r-x)^2) over n_pix, sums[1] = sum(-0.5(1+lv-mu^2-e^lv)) over n_lat
__global__ void __launch_bounds__(256) elbo_fwd_kernel(
    const float* __restrict__ recon, const float* __restrict__ x, int64_t n_pix,
    const float* __restrict__ mu, const float* __restrict__ lv, int n_lat, float* __restrict__ sums,
    float* __restrict__ scratch) {
  __shared__ float red[32];
  __shared__ bool is_last;
  float a = 0.f;
  int64_t n4 = n_pix >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 r = __ldg(r4 + i), t = __ldg(x4 + i);
    float d0 = r.x - t.x, d1 = r.y - t.y, d2 = r.z - t.z, d3 = r.w - t.w;
    a += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += stride) {
    float d = recon[i] - x[i];
    a += d * d;
  }
  float k = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lat; i += stride) {
    float m = mu[i], l = lv[i];
    k += -0.5f * (1.f + l - m * m - expf(l));
  }
  a = block_sum(a, red);
  k = block_sum(k, red);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + 2 * kElboMaxBlocks);
  if (threadIdx.x == 0) {
    scratch[2 * blockIdx.x] = a;
    scratch[2 * blockIdx.x + 1] = k;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float sa = 0.f, sk = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      sa += __ldcg(scratch + 2 * i);
      sk += __ldcg(scratch + 2 * i + 1);
    }
    sa = block_sum(sa, red);
    sk = block_sum(sk, red);
    if (threadIdx.x == 0) {
      sums[0] = sa;
      sums[1] = sk;
      *ticket = 0u;  // self-resetting
    }
  }
}

// d_recon = g[0]*2(r-x), d_x = -d_recon (optional), d_mu = g[1]*mu, d_lv = g[1]*0.5(e^lv-1)
__global__ void __launch_bounds__(256) elbo_bwd_kernel(
    const float* __restrict__ recon, const float* __restrict__ x, int64_t n_pix,
    const float* __restrict__ mu, const float* __restrict__ lv, int n_lat,
    const float* __restrict__ g, float* __restrict__ d_recon, float* __restrict__ d_x,
    float* __restrict__ d_mu, float* __restrict__ d_lv) {
  float g0 = 2.f * g[0];
  int64_t n4 = n_pix >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 r = __ldg(r4 + i), t = __ldg(x4 + i);
    float4 o = make_float4(g0 * (r.x - t.x), g0 * (r.y - t.y), g0 * (r.z - t.z), g0 * (r.w - t.w));
    if (d_recon) reinterpret_cast<float4*>(d_recon)[i] = o;
    if (d_x) reinterpret_cast<float4*>(d_x)[i] = make_float4(-o.x, -o.y, -o.z, -o.w);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += stride) {
    float o = g0 * (recon[i] - x[i]);
    if (d_recon) d_recon[i] = o;
    if (d_x) d_x[i] = -o;
  }
  if (n_lat > 0) {
    float g1 = g[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lat; i += stride) {
      d_mu[i] = g1 * mu[i];
      d_lv[i] = g1 * 0.5f * (expf(lv[i]) - 1.f);
    }
  }
}

__global__ void __launch_bounds__(1024) cycle_fwd_kernel(const float* __restrict__ th,
                                                         const float* __restrict__ thr,
                                                         const float* __restrict__ ang, int B,
                                                         float* __restrict__ loss) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) a += 1.f - cosf((thr[i] - th[i]) + ang[i]);
  a = block_sum(a, red);
  if (threadIdx.x == 0) loss[0] = a / (float)B;
}

__global__ void cycle_bwd_kernel(const float* __restrict__ th, const float* __restrict__ thr,
                                 const float* __restrict__ ang, const float* __restrict__ g, int B,
                                 float* __restrict__ d_th, float* __restrict__ d_thr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float v = sinf((thr[i] - th[i]) + ang[i]) * g[0] / (float)B;
  if (d_th) d_th[i] = -v;
  if (d_thr) d_thr[i] = v;
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ a,
                             const float* __restrict__ y, const float* __restrict__ b, int64_t n,
                             float* __restrict__ out) {
  float fa = a ? a[0] : 1.f, fb = b ? b[0] : 1.f;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = fa * x[i] + (y ? fb * y[i] : 0.f);
}

static inline int grid_for(int64_t n, int per_thread = 1) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (blocks < 1) blocks = 1;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace livae

using namespace livae;

extern "C" int livae_stn_head_fwd(const float* vec, int B, float* cs, float* theta, livae_stream_t stream) {
  LIVAE_CHECK_ARG(vec && cs && B >= 0, "stn_head_fwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  stn_head_fwd_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(vec, B, cs, theta);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_stn_tail_fwd(const float* f1, const float* w9, const float* b9, int B, int K, float* vec,
                                  float* cs, float* theta, livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && K == 32, "stn_tail_fwd: the localisation head is Linear(32 -> 2) (model.py:213)");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(f1 && w9 && b9 && vec && cs, "stn_tail_fwd: null pointer");
  if (int e = livae::require_sm100()) return e;
  const int grid = B >= 8 * 296 ? 296 : (B + 7) / 8;
  livae::stn_tail_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(f1, w9, b9, B, vec, cs, theta);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_stn_tail_bwd(const float* f1, const float* w9, const float* vec, const float* gcs,
                                  const float* gtheta, int B, int K, float* gw9, float* gb9, void* gf1_bf16,
                                  livae_stream_t stream) {
  LIVAE_CHECK_ARG(B >= 0 && K == 32, "stn_tail_bwd: the localisation head is Linear(32 -> 2) (model.py:213)");
  LIVAE_CHECK_ARG(gw9 && gb9, "stn_tail_bwd: null pointer");
  cudaError_t ce;
  if ((ce = cudaMemsetAsync(gw9, 0, 64 * sizeof(float), (cudaStream_t)stream)) != cudaSuccess ||
      (ce = cudaMemsetAsync(gb9, 0, 2 * sizeof(float), (cudaStream_t)stream)) != cudaSuccess) {
    livae::set_error("stn_tail_bwd: memset failed");
    return (int)ce;
  }
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(f1 && w9 && vec && gf1_bf16, "stn_tail_bwd: null pointer");
  if (int e = livae::require_sm100()) return e;
  const int grid = B >= 8 * 64 ? 64 : (B + 7) / 8;
  livae::stn_tail_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(f1, w9, vec, gcs, gtheta, B, gw9, gb9,
                                                                     (__nv_bfloat16*)gf1_bf16);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_stn_head_bwd(const float* vec, const float* gcs, const float* gtheta, int B,
                                  float* gvec, livae_stream_t stream) {
  LIVAE_CHECK_ARG(vec && gvec && B >= 0, "stn_head_bwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  stn_head_bwd_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(vec, gcs, gtheta, B, gvec);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_angle_to_cs(const float* theta, int B, float* cs, livae_stream_t stream) {
  LIVAE_CHECK_ARG(theta && cs && B >= 0, "angle_to_cs: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  angle_to_cs_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta, B, cs);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_angle_to_cs_bwd(const float* theta, const float* gcs, int B, float* gtheta,
                                     livae_stream_t stream) {
  LIVAE_CHECK_ARG(theta && gcs && gtheta && B >= 0, "angle_to_cs_bwd: bad args");
  if (int e = require_sm100()) return e;
  if (B == 0) return 0;
  angle_to_cs_bwd_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta, gcs, B, gtheta);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_reparam_fwd(const float* mu, const float* logvar, const float* eps, int n, float* z,
                                 livae_stream_t stream) {
  LIVAE_CHECK_ARG(mu && logvar && eps && z && n >= 0, "reparam_fwd: bad args");
  if (int e = require_sm100()) return e;
  if (n == 0) return 0;
  reparam_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mu, logvar, eps, n, z);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_reparam_bwd(const float* gz, const float* logvar, const float* eps, int n,
                                 float* gmu, float* glv, livae_stream_t stream) {
  LIVAE_CHECK_ARG(gz && logvar && eps && gmu && glv && n >= 0, "reparam_bwd: bad args");
  if (int e = require_sm100()) return e;
  if (n == 0) return 0;
  reparam_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gz, logvar, eps, n, gmu, glv);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t livae_elbo_scratch_floats(void) { return 2 * kElboMaxBlocks + 4; }

extern "C" int livae_elbo_fwd(const float* recon, const float* x, int64_t n_pix, const float* mu,
                              const float* logvar, int n_lat, float* sums, float* scratch,
                              livae_stream_t stream) {
  LIVAE_CHECK_ARG(recon && x && sums && scratch && n_pix >= 0, "elbo_fwd: bad args");
  LIVAE_CHECK_ARG(n_lat == 0 || (mu && logvar), "elbo_fwd: n_lat > 0 needs mu and logvar");
  LIVAE_CHECK_ARG((((uintptr_t)recon | (uintptr_t)x) & 15) == 0, "elbo_fwd: pointers must be 16-byte aligned");
  if (int e = require_sm100()) return e;
  int grid = grid_for(n_pix, 16);
  if (grid > kElboMaxBlocks) grid = kElboMaxBlocks;
  elbo_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(recon, x, n_pix, mu, logvar, n_lat, sums, scratch);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_elbo_bwd(const float* recon, const float* x, int64_t n_pix, const float* mu,
                              const float* logvar, int n_lat, const float* g, float* d_recon, float* d_x,
                              float* d_mu, float* d_lv, livae_stream_t stream) {
  LIVAE_CHECK_ARG(recon && x && g && n_pix >= 0, "elbo_bwd: bad args");
  LIVAE_CHECK_ARG(n_lat == 0 || (mu && logvar && d_mu && d_lv), "elbo_bwd: n_lat > 0 needs mu/logvar/d_mu/d_lv");
  LIVAE_CHECK_ARG((((uintptr_t)recon | (uintptr_t)x | (uintptr_t)d_recon | (uintptr_t)d_x) & 15) == 0,
                  "elbo_bwd: pointers must be 16-byte aligned");
  if (int e = require_sm100()) return e;
  elbo_bwd_kernel<<<grid_for(n_pix, 8), 256, 0, (cudaStream_t)stream>>>(recon, x, n_pix, mu, logvar, n_lat, g,
                                                                        d_recon, d_x, d_mu, d_lv);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_cycle_fwd(const float* theta, const float* theta_rot, const float* angle, int B,
                               float* loss, livae_stream_t stream) {
  LIVAE_CHECK_ARG(theta && theta_rot && angle && loss && B > 0, "cycle_fwd: bad args");
  if (int e = require_sm100()) return e;
  cycle_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(theta, theta_rot, angle, B, loss);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_cycle_bwd(const float* theta, const float* theta_rot, const float* angle,
                               const float* g, int B, float* d_theta, float* d_theta_rot,
                               livae_stream_t stream) {
  LIVAE_CHECK_ARG(theta && theta_rot && angle && g && B > 0, "cycle_bwd: bad args");
  if (int e = require_sm100()) return e;
  cycle_bwd_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(theta, theta_rot, angle, g, B, d_theta,
                                                                      d_theta_rot);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_axpby_dev(const float* x, const float* a_dev, const float* y, const float* b_dev,
                               int64_t n, float* out, livae_stream_t stream) {
  LIVAE_CHECK_ARG(x && out && n >= 0, "axpby: bad args");
  if (int e = require_sm100()) return e;
  if (n == 0) return 0;
  axpby_kernel<<<grid_for(n, 4), 256, 0, (cudaStream_t)stream>>>(x, a_dev, y, b_dev, n, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

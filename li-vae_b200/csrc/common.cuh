// Shared helpers for liblivae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/livae_b200.h"

namespace livae {

void set_error(const char* fmt, ...);
// number of kernels this library has launched (reported by bench.py as gpu_launches)
void count_launch(int n = 1);

#define LIVAE_CHECK_ARG(cond, ...)                     \
  do {                                                 \
    if (!(cond)) {                                     \
      livae::set_error(__VA_ARGS__);                   \
      return -1;                                       \
    }                                                  \
  } while (0)

#define LIVAE_CUDA_LAUNCH_CHECK()                                                   \
  do {                                                                              \
    livae::count_launch(1);                                                         \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      livae::set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__));   \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

// refuse to run anywhere but sm_100 (no fallback paths)
int require_sm100();

// "do this once PER DEVICE": cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the compute-capability check are
// per-device state, so a process driving several GPUs must repeat them on each (a process-wide flag left every
// device but the first without the > 48 KB shared-memory opt-in).
struct OncePerDevice {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

static constexpr int kNumSMs = 148;

// Caller-owned device scratch (livae_set_scratch, one buffer per device): where kernels that split a reduction over
// CTAs park their per-CTA partial sums so that a second pass can add them in a FIXED order (bit-reproducible results;
// red.global.add / atomicAdd commit in whatever order the CTAs retire).  Valid for work enqueued on ONE stream at a
// time.  nullptr when no (or too small a) buffer was registered: callers then fall back to atomics.
float* scratch_floats(int64_t nfloats);

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  __device__ __forceinline__ static float ld(const float* p, int64_t i) { return p[i]; }
  __device__ __forceinline__ static void st(float* p, int64_t i, float v) { p[i] = v; }
};
template <> struct Cvt<__half> {
  __device__ __forceinline__ static float ld(const __half* p, int64_t i) { return __half2float(p[i]); }
  __device__ __forceinline__ static void st(__half* p, int64_t i, float v) { p[i] = __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ __forceinline__ static float ld(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
  __device__ __forceinline__ static void st(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

inline size_t dt_size(int dt) { return dt == LIVAE_F32 ? 4 : 2; }

}  // namespace livae

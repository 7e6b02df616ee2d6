// Micro-benchmark: cycles per tcgen05.mma for the operand layouts of the thin 1-channel kernels (thin_tc.cu):
//   fold : conv1_wgrad_fold_kernel -- A MN-major 128B-swizzled (two M atoms of 64, LBO 2048, SBO 1024), B K-major
//          32-byte rows (swizzle 32B, SBO 256), M = 128, N = 64, K = 16
//   win  : conv1c_tc_kernel<0> -- A and B K-major 128-byte rows, N = 64
//   n16  : the previous conv1c_tc_kernel<0> -- K-major 64-byte rows, N = 16
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I li-vae_b200/csrc -I include tools/umma_fold_rate.cu -o tools/bin/umma_fold_rate -lcuda
#include <cstdio>
#include "tc_common.cuh"
using namespace livae::tc;

__global__ void __launch_bounds__(192) rate_kernel(int mode, int reps, int issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp >= 1 && warp <= issuers) {
    const bool leader = elect_one();
    const int w = warp - 1;
    const uint32_t a0 = smem_u32(smem) + (uint32_t)w * 16384u, b0 = smem_u32(smem) + 64 * 1024;
    uint32_t idesc; uint64_t ad, bd; uint32_t astep, bstep;
    if (mode == 0) {        // fold: 4 row tiles of 4096 B per stage, B window moves by 12 rows of 32 B per MMA pair
      idesc = make_idesc_bf16(128, 64, 1, 0);
      ad = make_smem_desc(a0, 2048u, 1024u, 2u); bd = make_smem_desc(b0, 16u, 256u, 6u);
      astep = 4096u >> 4; bstep = (12u * 32u) >> 4;
    } else if (mode == 1) { // win
      idesc = make_idesc_f16(128, 64, 0, 0);
      ad = make_smem_desc(a0, 16u, 1024u, 2u); bd = make_smem_desc(b0, 16u, 1024u, 2u);
      astep = 2u; bstep = 2u;
    } else {                // n16
      idesc = make_idesc_f16(128, 16, 0, 0);
      ad = make_smem_desc(a0, 16u, 512u, 4u); bd = make_smem_desc(b0, 16u, 512u, 4u);
      astep = 2u; bstep = 2u;
    }
    const uint32_t d0 = tmem_base + (uint32_t)w * 128u;
    for (int i = 0; i < 8; ++i) if (leader) umma_f16(d0, ad, bd, idesc, 1u);
    if (leader) umma_commit(&bar[w]);
    mbar_wait(&bar[w], 0);
    tc_fence_after();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i += 4) {
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) if (leader) umma_f16(d0, ad + (k & 3u) * astep, bd + k * bstep, idesc, 1u);
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&bar[w]);
    mbar_wait(&bar[w], 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && w == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int reps = 2048;
  const char* names[3] = {"fold (A MN-major 128B, B K-major 32B rows, N=64)", "win  (K-major 128B rows, N=64)", "n16  (K-major 64B rows, N=16)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int issuers : {1, 2}) {
      rate_kernel<<<148, 192, 170 * 1024>>>(mode, reps, issuers, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-52s issuers %d : issue %.1f  complete %.1f cycles per MMA per SM\n", names[mode], issuers,
             (double)h[0] / reps / issuers, (double)h[1] / reps / issuers);
    }
  return 0;
}

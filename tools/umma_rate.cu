// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N and of the operand
// majors, with both operands in shared memory (the SS form every kernel of this repo uses).  Zero data;
// the descriptors use the same LBO/SBO/swizzle conventions as conv_tc.cu (K-major) and wgrad_tc.cu (MN-major).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I li-vae_b200/csrc -I include tools/umma_rate.cu -o tools/bin/umma_rate -lcuda
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace livae::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <bool UNIFORM>
__global__ void __launch_bounds__(192) rate_kernel(int N, int a_mn, int b_mn, int reps, int issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp >= 1 && warp <= issuers && (UNIFORM || lane == 0)) {
    // UNIFORM: the whole warp runs the loop (descriptor arithmetic stays warp-uniform), one elected lane issues
    const bool leader = UNIFORM ? elect_one() : true;
    const int w = warp - 1;
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(smem) + (uint32_t)w * 8192u, b0 = smem_u32(smem) + 48 * 1024 + (uint32_t)w * 8192u;
    const int kcs = N >= 64 ? 64 : N;
    const uint32_t rbg = kcs * 2;
    // K-major: 128-byte rows (kc = 64), LBO 16, SBO 8 rows.  MN-major: boxes of 64 pixels x 64 channels
    uint64_t ad = a_mn ? make_smem_desc(a0, 8192u, 1024u, 2u) : make_smem_desc(a0, 16u, 1024u, 2u);
    uint64_t bd = b_mn ? make_smem_desc(b0, 64u * rbg, 8u * rbg, rbg == 128 ? 2u : 4u) : make_smem_desc(b0, 16u, 1024u, 2u);
    // warm-up
    const uint32_t d0 = tmem_base + (uint32_t)w * 128u;
    for (int i = 0; i < 8; ++i) if (leader) umma_f16(d0, ad, bd, idesc, 1u);
    if (leader) umma_commit(&bar[w]);
    mbar_wait(&bar[w], 0);
    tc_fence_after();
    const long long t0 = clock64();
    const uint32_t d1 = d0 + (N <= 64 ? 64u : 0u);
#pragma unroll 1
    for (int i = 0; i < reps; i += 8) {
      // +32 bytes per K step, as in a real K loop; two accumulators alternate
#pragma unroll
      for (int k = 0; k < 4; ++k) if (leader) umma_f16(d0, ad + 2u * k, bd + 2u * k, idesc, 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (leader) umma_f16(d1, ad + 2u * k, bd + 2u * k, idesc, 1u);
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

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2048;
  printf("cycles per tcgen05.mma (M=128, K=16, bf16, SS), %d back-to-back MMAs, 148 CTAs; ideal = N/2\n", reps);
  for (int uni = 0; uni < 2; ++uni)
  for (int issuers : {1, 2})
    for (int maj : {0}) {
      const int a_mn = maj & 1, b_mn = maj >> 1;
      for (int N : {16, 32, 64, 128}) {
        if (issuers > 1 && N > 128) continue;
        if (uni) rate_kernel<true><<<148, 192, 100 * 1024>>>(N, a_mn, b_mn, reps, issuers, d);
        else rate_kernel<false><<<148, 192, 100 * 1024>>>(N, a_mn, b_mn, reps, issuers, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s issuers %d  A %s  B %s  N=%3d : issue %.1f  complete %.1f cyc per MMA per SM (ideal %.0f)\n", uni ? "warp-uniform loop" : "lane-0 loop      ", issuers, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ",
               N, (double)h[0] / reps / issuers, (double)h[1] / reps / issuers, N / 2.0);
      }
    }
  return 0;
}

// Data gradient of the two wide nn.Linear layers (STN fc1: 32*(P/4)^2 -> 32, model.py:211; latent heads:
// 256*(P/16)^2 -> 2L, model.py:302-303) -- gx[b][k] = (mask[b][k] > 0) * sum_j g[b][j] * w[k][j], J = 16 / 32 / 64.
//
// HBM-bound: the contraction is J <= 64 long, the cost is writing B x K bf16 (134 MB at B = 2048, P = 128) and
// reading the equally large ReLU mask; 2*B*K*J FLOP is noise.  As a 1x1 convolution on the tcgen05 engine every
// epilogue thread owned one 64 KB-strided row and wrote / read it in 32-byte pieces (0.87 TB/s, 13 % of the copy
// rate).  Here the math runs on warp-level mma.sync (m16n8k16, bf16 in, fp32 accumulate -- the same operand
// rounding as before) and each warp passes its 32 x 64 result through shared memory so that mask reads and output
// writes are 128-byte row segments.
#include "common.cuh"

namespace livae {

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// block = 8 warps: 32 rows (b) x 512 columns (k); warp w owns columns [64 w, 64 w + 64)
template <int J>
__global__ void __launch_bounds__(256) linear_dgrad_kernel(const __nv_bfloat16* __restrict__ g,      // [B][J]
                                                           const __nv_bfloat16* __restrict__ w,      // [K][J]
                                                           const __nv_bfloat16* __restrict__ mask,   // [B][K] or null
                                                           int B, int K, __nv_bfloat16* __restrict__ out) {
  constexpr int KS = J / 16;
  __shared__ __align__(16) __nv_bfloat16 stage[8][32][72];       // +8 columns of padding: conflict-free 16-byte rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int b0 = blockIdx.y * 32, k0 = blockIdx.x * 512 + warp * 64;
  if (k0 >= K) return;
  // A fragments: rows b0 + mt*16 + {gid, gid + 8}, columns ks*16 + tig*2 + {0, 1, 8, 9}
  uint32_t a[2][KS][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = b0 + mt * 16 + gid + h * 8;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(g + (int64_t)(r < B ? r : 0) * J + ks * 16 + tig * 2);
        a[mt][ks][h] = r < B ? __ldg(row) : 0u;
        a[mt][ks][2 + h] = r < B ? __ldg(row + 4) : 0u;
      }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int kcol = k0 + nt * 8 + gid;                         // B fragment: n = gid, k = tig*2 + {0,1} (+8)
    float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t bb0 = 0u, bb1 = 0u;
      if (kcol < K) {
        const uint32_t* wr = reinterpret_cast<const uint32_t*>(w + (int64_t)kcol * J + ks * 16 + tig * 2);
        bb0 = __ldg(wr); bb1 = __ldg(wr + 4);
      }
      mma_16816(d[0], a[0][ks], bb0, bb1);
      mma_16816(d[1], a[1][ks], bb0, bb1);
    }
    // C fragment: rows gid / gid + 8 of each 16-row tile, columns nt*8 + tig*2 + {0, 1}
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      *reinterpret_cast<__nv_bfloat162*>(&stage[warp][mt * 16 + gid][nt * 8 + tig * 2]) = __floats2bfloat162_rn(d[mt][0], d[mt][1]);
      *reinterpret_cast<__nv_bfloat162*>(&stage[warp][mt * 16 + gid + 8][nt * 8 + tig * 2]) = __floats2bfloat162_rn(d[mt][2], d[mt][3]);
    }
  }
  __syncwarp();
  // 8 lanes x 16 bytes = one 128-byte row segment; 4 rows per pass
  const int seg = lane & 7, rsub = lane >> 3;
  const int kc = k0 + seg * 8;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int r = pass * 4 + rsub, b = b0 + r;
    if (b >= B || kc >= K) continue;
    uint4 v = *reinterpret_cast<const uint4*>(&stage[warp][r][seg * 8]);
    if (mask) {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(mask + (int64_t)b * K + kc));
      const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m);
      __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(&v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (!(__bfloat162float(mb[i]) > 0.f)) vb[i] = __float2bfloat16_rn(0.f);
    }
    *reinterpret_cast<uint4*>(out + (int64_t)b * K + kc) = v;
  }
}

}  // namespace livae

using namespace livae;

// gx bf16 [B][K] = (mask > 0) * (g bf16 [B][J] @ w bf16 [K][J]^T); w is livae_tc_pack_weights mode 4 ([tap][Cb][J]).
// J in {16, 32, 64}; K a multiple of 8; mask may be NULL.
extern "C" int livae_linear_dgrad(const void* g, const void* w_kj, const void* mask, int B, int K, int J, void* gx,
                                  livae_stream_t stream) {
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(g && w_kj && gx && B > 0 && K > 0 && (K & 7) == 0, "linear_dgrad: bad args");
  LIVAE_CHECK_ARG(J == 16 || J == 32 || J == 64, "linear_dgrad: J must be 16, 32 or 64");
  LIVAE_CHECK_ARG((((uintptr_t)g | (uintptr_t)w_kj | (uintptr_t)gx | (uintptr_t)mask) & 15) == 0, "linear_dgrad: alignment");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((K + 511) / 512, (B + 31) / 32);
  const __nv_bfloat16 *gp = (const __nv_bfloat16*)g, *wp = (const __nv_bfloat16*)w_kj, *mp = (const __nv_bfloat16*)mask;
  if (J == 16) linear_dgrad_kernel<16><<<grid, 256, 0, st>>>(gp, wp, mp, B, K, (__nv_bfloat16*)gx);
  else if (J == 32) linear_dgrad_kernel<32><<<grid, 256, 0, st>>>(gp, wp, mp, B, K, (__nv_bfloat16*)gx);
  else linear_dgrad_kernel<64><<<grid, 256, 0, st>>>(gp, wp, mp, B, K, (__nv_bfloat16*)gx);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

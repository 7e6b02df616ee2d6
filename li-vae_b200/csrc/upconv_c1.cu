// Decoder layer d4 = Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv3x3(32 -> 1) (reference model.py:369-372),
// backward w.r.t. its LOW-RESOLUTION input, without materialising the up-sampled gradient.
//
// Before: thin_conv1c_fwd[2] wrote the gradient of the padded up-sampled tensor (bf16 [B,2H+2,2W+2,32], 2.2 GB at
// B = 2048, H = 64) and upsample_pad_bwd read it back: 4.4 GB of HBM traffic for 0.54 GB of result.  Here one CTA
// owns a 16x16 tile of low-res pixels:
//   1. stage the 36x36 window of the 1-channel pre-activation gradient g it depends on (rounded to tf32);
//   2. gU[Yp,Xp,ci] = sum_{ky,kx} g[Yp-ky, Xp-kx] w[ci,ky,kx] for the 34x34 padded up-sampled positions of the
//      tile: a [1156 x 9 taps] x [9 x 32] GEMM on the warp-level tensor-core path (mma.sync m16n8k8 tf32,
//      fp32 accumulate; A fragments gathered from the g window, B fragments = the weights, resident in registers).
//      Rounded to bf16 into shared memory -- the same rounding the materialised tensor had;
//   3. adjoint of upsample + reflect-pad as a separable gather over gU (per thread: 10 padded rows -> 4 low-res
//      rows of one column, weights 0.25/0.75/1 in closed form incl. the clamped and reflected borders), ReLU
//      mask of the layer below, bf16 store, and the bias-gradient column sums of what was written.
// tcgen05 is not used here: K = 9 and the operand is a gather of a 1-channel image -- building an smem operand
// tile for UMMA costs more than the whole MMA work (292 warp-level MMAs per CTA).
// Algorithmic bytes per patch: g 4*(2H)(2W) + mask 2*32*H*W + out 2*32*H*W  (64 KB + 256 KB + 256 KB at H = 64).
#include "common.cuh"

namespace livae {

namespace {
constexpr int kT = 16;                 // low-res tile edge
constexpr int kR = 2 * kT + 2;         // 34: padded up-sampled rows / cols per tile
constexpr int kG = kR + 2;             // 36: g window edge
constexpr int kPos = kR * kR;          // 1156 positions
constexpr int kPitch = 80;             // bytes per position in the gU tile (64 B of bf16 + pad: conflict-free stores)
constexpr int kC = 32;

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
}  // namespace

__global__ void __launch_bounds__(256, 2) upconv_c1_bwd_data_kernel(const float* __restrict__ gpre,
                                                                    const float* __restrict__ w,
                                                                    const __nv_bfloat16* __restrict__ ymask, int H,
                                                                    int W, __nv_bfloat16* __restrict__ gy,
                                                                    float* __restrict__ gb) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* gU = smem;                                               // [kPos][kPitch]
  uint32_t* gs = reinterpret_cast<uint32_t*>(smem + kPos * kPitch);       // [kG][kG] tf32 bit patterns
  float* colsum = reinterpret_cast<float*>(gs + kG * kG);                 // [8][32]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.z, i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
  const int H2 = 2 * H, W2 = 2 * W;

  // ---- 1. g window: padded position (Yp, Xp) feeds output (Yp - ky, Xp - kx); window origin = (2 i0 - 2, 2 j0 - 2)
  const float* gimg = gpre + (int64_t)b * H2 * W2;
  for (int e = tid; e < kG * kG; e += 256) {
    const int r = e / kG, c = e - r * kG;
    const int Y = 2 * i0 - 2 + r, X = 2 * j0 - 2 + c;
    float v = 0.f;
    if (Y >= 0 && Y < H2 && X >= 0 && X < W2) v = __ldg(gimg + (int64_t)Y * W2 + X);
    gs[e] = to_tf32(v);
  }
  // ---- B fragments (m16n8k8 tf32): B[k = tap][n = ci] = w[ci][tap]; b0 = (k = t, n = g), b1 = (k = t + 4, n = g);
  // second k-step: tap 8 at k = 0, nothing else
  const int gq = lane >> 2, tq = lane & 3;
  uint32_t bA[4], bB[4], bC[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float* wr = w + (nt * 8 + gq) * 9;
    bA[nt] = to_tf32(__ldg(wr + tq));
    bB[nt] = to_tf32(__ldg(wr + tq + 4));
    bC[nt] = tq == 0 ? to_tf32(__ldg(wr + 8)) : 0u;
  }
  // window offset of tap k relative to position (r, c): (r + 2 - ky) * kG + (c + 2 - kx)
  const int kb = tq + 4;
  const int offa = (2 - tq / 3) * kG + (2 - tq % 3), offb = (2 - kb / 3) * kG + (2 - kb % 3);
  // the ReLU-mask vectors of this thread's four output pixels, requested before the tile maths
  const int o = tid & 3, lj = (tid >> 2) & (kT - 1), seg = tid >> 6;      // channel octet, column, 4-row segment
  const int j = j0 + lj;
  uint4 mk[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + 4 * seg + q;
    mk[q] = make_uint4(0u, 0u, 0u, 0u);
    if (i < H && j < W) mk[q] = __ldg(reinterpret_cast<const uint4*>(ymask + (((int64_t)b * H + i) * W + j) * kC) + o);
  }
  __syncthreads();

  // ---- 2. gU tile by warp-level MMA
  for (int mt = wid; mt * 16 < kPos; mt += 8) {
    const int m0 = mt * 16 + gq, m1 = m0 + 8;
    const bool v0 = m0 < kPos, v1 = m1 < kPos;
    const int p0 = v0 ? (m0 / kR) * kG + (m0 % kR) : 0, p1 = v1 ? (m1 / kR) * kG + (m1 % kR) : 0;
    uint32_t a[4], a8[4];
    a[0] = gs[p0 + offa]; a[1] = gs[p1 + offa]; a[2] = gs[p0 + offb]; a[3] = gs[p1 + offb];
    a8[0] = tq == 0 ? gs[p0] : 0u; a8[1] = tq == 0 ? gs[p1] : 0u; a8[2] = 0u; a8[3] = 0u;     // tap 8: ky = kx = 2
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      mma_tf32_1688(d, a, bA[nt], bB[nt]);
      mma_tf32_1688(d, a8, bC[nt], 0u);
      if (v0) *reinterpret_cast<uint32_t*>(gU + m0 * kPitch + nt * 16 + tq * 4) = pack_bf16(d[0], d[1]);
      if (v1) *reinterpret_cast<uint32_t*>(gU + m1 * kPitch + nt * 16 + tq * 4) = pack_bf16(d[2], d[3]);
    }
  }
  __syncthreads();

  // ---- 3. adjoint of upsample + reflect-pad.  Thread = (4 consecutive low-res rows, one column, 8 channels): it
  // walks the 10 padded rows those outputs read, takes the horizontal adjoint h of each row ONCE (4 column taps
  // 2j..2j+3 with weights .25 .75 .75 .25, border columns 0 / W-1 absorb the clamped tap, columns 1 / W-2 also
  // read the reflected pad column) and deals it to the <= 2 outputs that read the row.
  float2 wx[4] = {{0.25f, 0.25f}, {0.75f, 0.75f}, {0.75f, 0.75f}, {0.25f, 0.25f}};
  if (j == 0) { wx[0] = make_float2(0.75f, 0.75f); wx[1] = make_float2(1.f, 1.f); }
  if (j == W - 1) { wx[2] = make_float2(1.f, 1.f); wx[3] = make_float2(0.75f, 0.75f); }
  const int xe = j == 1 ? 0 : (j == W - 2 ? 2 * W + 1 - 2 * j0 : -1);    // local column of the reflected pad tap
  float2 acc[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[q][e] = make_float2(0.f, 0.f);
  const int ibase = i0 + 4 * seg;
  const unsigned char* gcol = gU + (2 * lj) * kPitch + o * 16;
#pragma unroll
  for (int rr = 0; rr < 10; ++rr) {
    const int r = 8 * seg + rr, Yp = 2 * i0 + r;
    if (Yp > 2 * H + 1) continue;
    float2 h[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    const unsigned char* grow = gcol + r * (kR * kPitch);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(grow + k * kPitch);
      const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __ffma2_rn(wx[k], make_float2(__uint_as_float(vv[e] << 16), __uint_as_float(vv[e] & 0xffff0000u)), h[e]);
    }
    if (xe >= 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(gU + (r * kR + xe) * kPitch + o * 16);
      const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
      const float2 qw = make_float2(0.25f, 0.25f);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h[e] = __ffma2_rn(qw, make_float2(__uint_as_float(vv[e] << 16), __uint_as_float(vv[e] & 0xffff0000u)), h[e]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = rr - 2 * q;                     // static after unrolling
      const int i = ibase + q;
      float wy = 0.f;
      if (k >= 0 && k < 4) {
        wy = (k == 0 || k == 3) ? 0.25f : 0.75f;
        if (i == 0 && k < 2) wy = k == 0 ? 0.75f : 1.f;
        if (i == H - 1 && k >= 2) wy = k == 2 ? 1.f : 0.75f;
      }
      if (rr == 0 && q == 1 && Yp == 0) wy = 0.25f;                       // i == 1 reads the reflected pad row 0
      if (rr == 2 * q + 5 && Yp == 2 * H + 1 && i == H - 2) wy = 0.25f;   // i == H-2 reads pad row 2H+1
      if ((k >= 0 && k < 4) || (rr == 0 && q == 1) || rr == 2 * q + 5) {
        const float2 w2 = make_float2(wy, wy);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][e] = __ffma2_rn(w2, h[e], acc[q][e]);
      }
    }
  }
  float cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ibase + q;
    if (i >= H || j >= W) continue;
    const uint32_t mm[4] = {mk[q].x, mk[q].y, mk[q].z, mk[q].w};
    uint4 outv;
    uint32_t* ov = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // y is a post-ReLU activation: > 0 <=> sign bit clear and magnitude non-zero
      const bool k0 = (mm[e] & 0x8000u) == 0 && (mm[e] & 0x7fffu) != 0;
      const bool k1 = (mm[e] & 0x80000000u) == 0 && (mm[e] & 0x7fff0000u) != 0;
      const __nv_bfloat162 v = __floats2bfloat162_rn(k0 ? acc[q][e].x : 0.f, k1 ? acc[q][e].y : 0.f);
      ov[e] = *reinterpret_cast<const uint32_t*>(&v);
      const float2 back = __bfloat1622float2(v);
      cs[2 * e] += back.x;
      cs[2 * e + 1] += back.y;
    }
    *(reinterpret_cast<uint4*>(gy + (((int64_t)b * H + i) * W + j) * kC) + o) = outv;
  }
  if (gb) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = cs[e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) colsum[wid * 32 + lane * 8 + e] = v;
    }
    __syncthreads();
    if (tid < 32) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += colsum[k * 32 + tid];
      atomicAdd(gb + tid, t);
    }
  }
}

}  // namespace livae

extern "C" int livae_upconv_c1_bwd_data(const float* gpre, const float* w, const void* y_bf16, int B, int H, int W,
                                        void* gy_bf16, float* gb, livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && H >= 4 && W >= 4 && H % 4 != 1 && W % 16 != 1,
                  "upconv_c1_bwd_data: bad sizes (H, W >= 4, H not 1 mod 4, W not 1 mod 16)");
  if (gb) {
    cudaError_t ce = cudaMemsetAsync(gb, 0, kC * sizeof(float), (cudaStream_t)stream);
    if (ce != cudaSuccess) { set_error("upconv_c1_bwd_data: memset failed"); return (int)ce; }
  }
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(gpre && w && y_bf16 && gy_bf16, "upconv_c1_bwd_data: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)y_bf16 | (uintptr_t)gy_bf16) & 15) == 0, "upconv_c1_bwd_data: 16-byte alignment");
  LIVAE_CHECK_ARG(B <= 65535, "upconv_c1_bwd_data: B > 65535");
  if (int e = require_sm100()) return e;
  const size_t smem = (size_t)kPos * kPitch + kG * kG * sizeof(uint32_t) + 8 * 32 * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(upconv_c1_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem);
    if (ce != cudaSuccess) { set_error("upconv_c1_bwd_data: cannot set %zu B of shared memory", smem); return (int)ce; }
    attr_done = true;
  }
  dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, B);
  upconv_c1_bwd_data_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      gpre, w, (const __nv_bfloat16*)y_bf16, H, W, (__nv_bfloat16*)gy_bf16, gb);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

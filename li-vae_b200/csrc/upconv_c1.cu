// Decoder layer d4 = Upsample(x2, bilinear) -> ReflectionPad2d(1) -> Conv3x3(32 -> 1) -> Sigmoid (reference
// model.py:369-372), forward and backward, computed from / to the LOW-RESOLUTION tensor.
//
// The materialising form (livae_upsample_pad_fwd_bf16 + thin 32->1 kernels) wrote the up-sampled, padded map
// (bf16 [B,2H+2,2W+2,32], 2.2 GB at B = 2048, H = 64) once and read it twice, and did the same with its gradient:
// 3.4 ms of a 22 ms step, all of it HBM traffic on tensors 4x larger than the layer's real input.  Up-sampling and
// reflect-padding act per channel, so they commute with the channel contraction of a convolution; with ONE output
// channel that turns the layer into small GEMMs over low-res pixels (mma.sync: K is 9 or 32 and one operand is a
// gather -- building UMMA shared-memory operand tiles would cost more than the MMA work) plus 1-channel bilinear
// index maps.  See the two section headers below for the algebra.
// Shared-memory tiles use an 80-byte pixel pitch (64 B of bf16 + 16): conflict-free for ldmatrix rows and for the
// C-fragment stores.
#include "common.cuh"

namespace livae {

namespace {
constexpr int kT = 16;                 // low-res tile edge
constexpr int kG = 2 * kT + 4;         // 36: edge of the gradient window a 16x16 low-res tile depends on
constexpr int kPitch = 80;             // bytes per pixel in the bf16 tiles (64 B + pad)
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

// =====================================================================================================================
// Whole backward of the layer in one kernel, through S = the adjoint-upsampled SHIFTED gradient.
//
// upsample + reflect-pad acts per channel, so it commutes with the channel contraction of the convolution:
//   y[Y,X]        = sum_tap upP( V[.,.,tap] )[Y+ky, X+kx],   V[i,j,tap] = sum_ci x[i,j,ci] w[ci,tap]        (forward)
//   S[i,j,tap]    = sum_{Yp,Xp} wy(Yp,i) wx(Xp,j) g[Yp-ky, Xp-kx]                 (adjoint of upP on 9 shifted copies of g)
//   gx[i,j,ci]    = sum_tap S[i,j,tap] w[ci,tap]                                            (data gradient, then ReLU mask)
//   gw[ci,tap]    = sum_{b,i,j} x[i,j,ci] S[i,j,tap]                                                  (weight gradient)
// S is a 9-channel LOW-resolution image computed from the 1-channel gradient with closed-form weights (0.25 / 0.75 /
// 1, clamped and reflected borders exact); both gradients are then small GEMMs over low-res pixels: 4x fewer MACs
// than at the up-sampled resolution, nothing of size [2H+2, 2W+2, 32] exists even in shared memory.
// Persistent CTAs (the weight gradient accumulates in registers across tiles; one round of atomics per CTA).
// Per 16x16 tile: g window 36x36 -> S (thread per pixel, separable, fp32) -> tf32 mma.sync for both GEMMs.
// Algorithmic bytes per patch: g 4(2H)(2W) + x 64 HW + gy 64 HW.
// =====================================================================================================================
namespace {
constexpr int kSP = 12;                // words per pixel in the S tile (9 taps + pad: conflict-free A-fragment reads)
}

// 1-D effective weights of the 6 gradient rows 2i-2 .. 2i+3 for tap offset k (0..2): v[k][m]
__device__ __forceinline__ void adj_weights(int i, int n, float (&v)[3][6]) {
  float W[4] = {0.25f, 0.75f, 0.75f, 0.25f};           // padded rows 2i .. 2i+3
  if (i == 0) { W[0] = 0.75f; W[1] = 1.f; }
  if (i == n - 1) { W[2] = 1.f; W[3] = 0.75f; }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      const int r = m + k - 2;                          // padded row 2i + r reads gradient row 2i - 2 + m at tap k
      v[k][m] = (r >= 0 && r < 4) ? W[r] : 0.f;
    }
  if (i == 1) v[0][0] += 0.25f;                         // reflected pad row 0 (copy of up-sampled row 1)
  if (i == n - 2) v[2][5] += 0.25f;                     // reflected pad row 2n+1 (copy of up-sampled row 2n-2)
}

__global__ void __launch_bounds__(256, 3) upconv_c1_bwd_kernel(const float* __restrict__ gpre,
                                                               const float* __restrict__ w,
                                                               const __nv_bfloat16* __restrict__ x, int B, int H, int W,
                                                               __nv_bfloat16* __restrict__ gy, float* __restrict__ gb_low,
                                                               float* __restrict__ gw, float* __restrict__ gb) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* gs = reinterpret_cast<float*>(smem);                              // [36][36] gradient window
  uint32_t* Ss = reinterpret_cast<uint32_t*>(gs + kG * kG);                // [256][kSP] tf32 bits
  unsigned char* xs = reinterpret_cast<unsigned char*>(Ss + 256 * kSP);    // [256][kPitch] bf16 x tile
  unsigned char* Gs = xs + 256 * kPitch;                                   // [256][kPitch] bf16 gx tile (unmasked)
  float* red = reinterpret_cast<float*>(Gs + 256 * kPitch);                // [8][32] + [8]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int tiles_x = (W + kT - 1) / kT, tiles_y = (H + kT - 1) / kT;
  const int64_t n_tiles = (int64_t)B * tiles_x * tiles_y;
  const int H2 = 2 * H, W2 = 2 * W;

  // B fragments of the data-gradient GEMM (m16n8k8 tf32, K = tap, N = ci): b0 = (k = t, n = g), b1 = (k = t + 4, n = g);
  // a second k-step carries tap 8 at k = 0
  uint32_t bA[4], bB[4], bC[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float* wr = w + (nt * 8 + gq) * 9;
    bA[nt] = to_tf32(__ldg(wr + tq));
    bB[nt] = to_tf32(__ldg(wr + tq + 4));
    bC[nt] = tq == 0 ? to_tf32(__ldg(wr + 8)) : 0u;
  }
  float dw[2][2][4];                                    // [m-tile (ci 0-15 / 16-31)][n-tile (taps 0-7 / 8)][frag]
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) dw[a][c][e] = 0.f;
  float cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.f;
  float gsum = 0.f;
  const int li = tid >> 4, lj = tid & 15;               // this thread's pixel of the tile (S stage)
  const int o = tid & 3;                                // channel octet (store stage)

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_x * tiles_y));
    const int trem = (int)(tile - (int64_t)b * tiles_x * tiles_y);
    const int i0 = (trem / tiles_x) * kT, j0 = (trem % tiles_x) * kT;
    // ---- 1. gradient window (origin 2 i0 - 2, 2 j0 - 2) and the x tile
    const float* gimg = gpre + (int64_t)b * H2 * W2;
    for (int e = tid; e < kG * kG; e += 256) {
      const int r = e / kG, c = e - r * kG;
      const int Y = 2 * i0 - 2 + r, X = 2 * j0 - 2 + c;
      float v = 0.f;
      if (Y >= 0 && Y < H2 && X >= 0 && X < W2) {
        v = __ldg(gimg + (int64_t)Y * W2 + X);
        if (r >= 2 && r < kG - 2 && c >= 2 && c < kG - 2) gsum += v;       // each output pixel belongs to one tile
      }
      gs[e] = v;
    }
    for (int e = tid; e < 256 * 4; e += 256) {
      const int p = e >> 2, oc = e & 3;
      const int i = i0 + (p >> 4), j = j0 + (p & 15);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (i < H && j < W) v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + i) * W + j) * kC) + oc);
      *reinterpret_cast<uint4*>(xs + p * kPitch + oc * 16) = v;
    }
    __syncthreads();
    // ---- 2. S[tap] for this thread's pixel: t[kx][m] = sum_n vx[kx][n] g[m][n];  S[ky][kx] = sum_m vy[ky][m] t[kx][m]
    {
      const int i = i0 + li, j = j0 + lj;
      float S[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) S[k] = 0.f;
      if (i < H && j < W) {
        float vy[3][6], vx[3][6];
        adj_weights(i, H, vy);
        adj_weights(j, W, vx);
        const float* gwin = gs + (2 * li) * kG + 2 * lj;
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          float g6[6];
#pragma unroll
          for (int n2 = 0; n2 < 3; ++n2) {
            const float2 t2 = *reinterpret_cast<const float2*>(gwin + m * kG + 2 * n2);
            g6[2 * n2] = t2.x; g6[2 * n2 + 1] = t2.y;
          }
          float t[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int n = 0; n < 6; ++n) t[kx] = fmaf(vx[kx][n], g6[n], t[kx]);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) S[ky * 3 + kx] = fmaf(vy[ky][m], t[kx], S[ky * 3 + kx]);
        }
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) Ss[tid * kSP + k] = to_tf32(S[k]);
    }
    __syncthreads();
    // ---- 3a. data gradient: [256 px x 9 taps] x [9 x 32 ci], two 16-pixel m-tiles per warp -> Gs (bf16)
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
      const int m0 = (wid * 2 + mm) * 16 + gq, m1 = m0 + 8;
      uint32_t a[4], a8[4];
      a[0] = Ss[m0 * kSP + tq]; a[1] = Ss[m1 * kSP + tq]; a[2] = Ss[m0 * kSP + tq + 4]; a[3] = Ss[m1 * kSP + tq + 4];
      a8[0] = tq == 0 ? Ss[m0 * kSP + 8] : 0u; a8[1] = tq == 0 ? Ss[m1 * kSP + 8] : 0u; a8[2] = 0u; a8[3] = 0u;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32_1688(d, a, bA[nt], bB[nt]);
        mma_tf32_1688(d, a8, bC[nt], 0u);
        *reinterpret_cast<uint32_t*>(Gs + m0 * kPitch + nt * 16 + tq * 4) = pack_bf16(d[0], d[1]);
        *reinterpret_cast<uint32_t*>(Gs + m1 * kPitch + nt * 16 + tq * 4) = pack_bf16(d[2], d[3]);
      }
    }
    // ---- 3b. weight gradient: dw[ci][tap] += sum_px x[px][ci] S[px][tap]; this warp owns pixels wid*32 .. +31
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int p0 = wid * 32 + ks * 8;
      // B (K = pixel, N = tap): b0 = (k = t, n = g), b1 = (k = t + 4, n = g)
      const uint32_t b00 = Ss[(p0 + tq) * kSP + gq], b01 = Ss[(p0 + tq + 4) * kSP + gq];
      const uint32_t b10 = gq == 0 ? Ss[(p0 + tq) * kSP + 8] : 0u, b11 = gq == 0 ? Ss[(p0 + tq + 4) * kSP + 8] : 0u;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        // A (M = ci, K = pixel): a0 = (row g, k = t), a1 = (row g+8, k = t), a2 = (row g, k = t+4), a3 = (row g+8, k = t+4)
        const unsigned char* xa = xs + (p0 + tq) * kPitch + (mt * 16 + gq) * 2;
        uint32_t a[4];
        a[0] = (uint32_t)(*reinterpret_cast<const uint16_t*>(xa)) << 16;
        a[1] = (uint32_t)(*reinterpret_cast<const uint16_t*>(xa + 16)) << 16;
        a[2] = (uint32_t)(*reinterpret_cast<const uint16_t*>(xa + 4 * kPitch)) << 16;
        a[3] = (uint32_t)(*reinterpret_cast<const uint16_t*>(xa + 4 * kPitch + 16)) << 16;
        mma_tf32_1688(dw[mt][0], a, b00, b01);
        mma_tf32_1688(dw[mt][1], a, b10, b11);
      }
    }
    __syncthreads();
    // ---- 4. ReLU mask of the layer below (x > 0), coalesced bf16 store, bias-gradient partial sums
    for (int e = tid; e < 256 * 4; e += 256) {
      const int p = e >> 2;
      const int i = i0 + (p >> 4), j = j0 + (p & 15);
      if (i >= H || j >= W) continue;
      const uint4 gv = *reinterpret_cast<const uint4*>(Gs + p * kPitch + o * 16);
      const uint4 mk = *reinterpret_cast<const uint4*>(xs + p * kPitch + o * 16);
      const uint32_t gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm4[4] = {mk.x, mk.y, mk.z, mk.w};
      uint4 outv;
      uint32_t* ov = reinterpret_cast<uint32_t*>(&outv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool k0 = (mm4[q] & 0x8000u) == 0 && (mm4[q] & 0x7fffu) != 0;
        const bool k1 = (mm4[q] & 0x80000000u) == 0 && (mm4[q] & 0x7fff0000u) != 0;
        ov[q] = (k0 ? (gg[q] & 0xffffu) : 0u) | (k1 ? (gg[q] & 0xffff0000u) : 0u);
        cs[2 * q] += __uint_as_float(ov[q] << 16);
        cs[2 * q + 1] += __uint_as_float(ov[q] & 0xffff0000u);
      }
      *(reinterpret_cast<uint4*>(gy + (((int64_t)b * H + i) * W + j) * kC) + o) = outv;
    }
    __syncthreads();
  }

  // ---- CTA epilogue: bias gradient of the layer below (column sums), then the weight / bias gradient of this layer
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = cs[e];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 4) red[wid * 32 + lane * 8 + e] = v;
  }
  gsum = warp_sum(gsum);
  if (lane == 0) red[256 + wid] = gsum;
  // dw fragments: c0,c1 = (ci = mt*16 + g, tap = nt*8 + 2t, +1), c2,c3 = (ci + 8, same taps); reuse the S tile
  float* dws = reinterpret_cast<float*>(Ss);                               // [8 warps][32 ci][9 taps]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ci = mt * 16 + gq + (e >> 1) * 8, tap = nt * 8 + 2 * tq + (e & 1);
        if (tap < 9) dws[(wid * 32 + ci) * 9 + tap] = dw[mt][nt][e];
      }
  __syncthreads();
  if (tid < 32 && gb_low) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k * 32 + tid];
    atomicAdd(gb_low + tid, t);
  }
  if (tid == 32 && gb) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[256 + k];
    atomicAdd(gb, t);
  }
  for (int e = tid; e < 32 * 9; e += 256) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += dws[k * 288 + e];
    atomicAdd(gw + e, t);
  }
}


// =====================================================================================================================
// Forward of the layer without the up-sampled tensor: V[i,j,tap] = sum_ci x[i,j,ci] w[ci,tap] (a 1x1 convolution to 9
// channels at LOW resolution, bf16 mma.sync with hi/lo-split weights, fp32 accumulate), then
// y[Y,X] = act(bias + sum_tap upP(V[.,.,tap])[Y+ky, X+kx]) with the same up_src / reflect index maps as the
// materialising kernels (aux.cu), in fp32.  Algorithmic bytes per patch: x 64 HW + y 4 (2H)(2W).
// =====================================================================================================================
namespace {
constexpr int kXT = kT + 2;            // 18: low-res rows / cols a 16x16 tile of outputs (x2) depends on
constexpr int kXP = kXT * kXT;         // 324 positions
constexpr int kVP = 9;                 // words per position in the V tile (odd: conflict-free gathers)

struct Lerp { int a, b; float f; };
// padded index p of the [2n+2] up-sampled + reflect-padded axis -> source taps, local to a tile whose first
// source index is org; same numbers as up_src(unpad(p)) in aux.cu
__device__ __forceinline__ Lerp pad_src(int p, int n, int org) {
  int u = p - 1;
  if (u < 0) u = -u;
  if (u >= 2 * n) u = 4 * n - 2 - u;
  Lerp r;
  const int a = u > 0 ? (u - 1) >> 1 : 0;
  r.f = u == 0 ? 0.f : ((u & 1) ? 0.25f : 0.75f);
  r.a = a - org;
  r.b = (a < n - 1 ? a + 1 : a) - org;
  return r;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace

__global__ void __launch_bounds__(256, 3) upconv_c1_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ w,
                                                               const float* __restrict__ bias, int H, int W, int act,
                                                               float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* xs = smem;                                                // [kXP][kPitch] bf16, clamped x tile
  float* Vs = reinterpret_cast<float*>(smem + kXP * kPitch);               // [kXP][kVP]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int b = blockIdx.z, i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
  // ---- 1. x tile, source rows i0-1 .. i0+16 clamped to the image (the clamp IS the up-sampling's edge rule)
  for (int e = tid; e < kXP * 4; e += 256) {
    const int p = e >> 2, oc = e & 3;
    const int i = min(max(i0 - 1 + p / kXT, 0), H - 1), j = min(max(j0 - 1 + p % kXT, 0), W - 1);
    *reinterpret_cast<uint4*>(xs + p * kPitch + oc * 16) =
        __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + i) * W + j) * kC) + oc);
  }
  // ---- B fragments (K = ci, N = tap): b0 = (k = 2t, 2t+1; n = g), b1 = (k = 2t+8, 2t+9; n = g); hi + lo halves
  uint32_t bh[2][2][2], bl[2][2][2];                    // [k-step][n-tile][b0/b1]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ci = ks * 16 + 2 * tq + 8 * h, tap = nt * 8 + gq;
        const float w0 = tap < 9 ? __ldg(w + ci * 9 + tap) : 0.f, w1 = tap < 9 ? __ldg(w + (ci + 1) * 9 + tap) : 0.f;
        const float w0h = __bfloat162float(__float2bfloat16_rn(w0)), w1h = __bfloat162float(__float2bfloat16_rn(w1));
        bh[ks][nt][h] = pack_bf16(w0h, w1h);
        bl[ks][nt][h] = pack_bf16(w0 - w0h, w1 - w1h);
      }
  __syncthreads();
  // ---- 2. V = x w  ([324 x 32] x [32 x 9]) by warp-level MMA
  for (int mt = wid; mt * 16 < kXP; mt += 8) {
    const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int pr = min(mt * 16 + row, kXP - 1);
    float d[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) d[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(a, xs + pr * kPitch + ks * 32 + (lane >> 4) * 16);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        mma_bf16_16816(d[nt], a, bl[ks][nt][0], bl[ks][nt][1]);
        mma_bf16_16816(d[nt], a, bh[ks][nt][0], bh[ks][nt][1]);
      }
    }
    const int m0 = mt * 16 + gq, m1 = m0 + 8;
    if (m0 < kXP) {
      Vs[m0 * kVP + 2 * tq] = d[0][0]; Vs[m0 * kVP + 2 * tq + 1] = d[0][1];
      if (tq == 0) Vs[m0 * kVP + 8] = d[1][0];
    }
    if (m1 < kXP) {
      Vs[m1 * kVP + 2 * tq] = d[0][2]; Vs[m1 * kVP + 2 * tq + 1] = d[0][3];
      if (tq == 0) Vs[m1 * kVP + 8] = d[1][2];
    }
  }
  __syncthreads();
  // ---- 3. y = act(bias + sum_tap upP(V[tap])[Y + ky, X + kx]); thread = one column X, four rows
  const int X = tid & 31, Yq = (tid >> 5) * 4;
  const int Xg = 2 * j0 + X;
  if (Xg >= 2 * W) return;
  Lerp cx[3];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) cx[kx] = pad_src(Xg + kx, W, j0 - 1);
  Lerp ry[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) ry[r] = pad_src(2 * i0 + Yq + r, H, i0 - 1);
  const float bv = __ldg(bias);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int Yg = 2 * i0 + Yq + k;
    if (Yg >= 2 * H) break;
    float acc = bv;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const Lerp r = ry[k + ky];
      const float* v0 = Vs + (r.a * kXT) * kVP + ky * 3;
      const float* v1 = Vs + (r.b * kXT) * kVP + ky * 3;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const Lerp c = cx[kx];
        // ATen's association: h0 * (w0 v00 + w1 v01) + h1 * (w0 v10 + w1 v11)
        const float top = (1.f - c.f) * v0[c.a * kVP + kx] + c.f * v0[c.b * kVP + kx];
        const float bot = (1.f - c.f) * v1[c.a * kVP + kx] + c.f * v1[c.b * kVP + kx];
        acc += (1.f - r.f) * top + r.f * bot;
      }
    }
    if (act == LIVAE_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    else if (act == LIVAE_ACT_RELU) acc = fmaxf(acc, 0.f);
    out[((int64_t)b * 2 * H + Yg) * (2 * W) + Xg] = acc;
  }
}

}  // namespace livae

extern "C" int livae_upconv_c1_bwd(const float* gpre, const float* w, const void* x_bf16, int B, int H, int W,
                                   void* gx_bf16, float* gb_low, float* gw, float* gb, livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && H >= 4 && W >= 4, "upconv_c1_bwd: bad sizes (H, W >= 4)");
  LIVAE_CHECK_ARG(gw, "upconv_c1_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce;
  if ((ce = cudaMemsetAsync(gw, 0, kC * 9 * sizeof(float), st)) != cudaSuccess ||
      (gb && (ce = cudaMemsetAsync(gb, 0, sizeof(float), st)) != cudaSuccess) ||
      (gb_low && (ce = cudaMemsetAsync(gb_low, 0, kC * sizeof(float), st)) != cudaSuccess)) {
    set_error("upconv_c1_bwd: memset failed");
    return (int)ce;
  }
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(gpre && w && x_bf16 && gx_bf16, "upconv_c1_bwd: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x_bf16 | (uintptr_t)gx_bf16) & 15) == 0, "upconv_c1_bwd: 16-byte alignment");
  if (int e = require_sm100()) return e;
  const size_t smem = (size_t)kG * kG * 4 + 256 * kSP * 4 + 2 * 256 * kPitch + (8 * 32 + 8) * 4;
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    ce = cudaFuncSetAttribute(upconv_c1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) { set_error("upconv_c1_bwd: cannot set %zu B of shared memory", smem); return (int)ce; }
  }
  const int64_t n_tiles = (int64_t)B * ((W + kT - 1) / kT) * ((H + kT - 1) / kT);
  const int grid = (int)(n_tiles < 3 * kNumSMs ? n_tiles : 3 * kNumSMs);
  upconv_c1_bwd_kernel<<<grid, 256, smem, st>>>(gpre, w, (const __nv_bfloat16*)x_bf16, B, H, W,
                                                (__nv_bfloat16*)gx_bf16, gb_low, gw, gb);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_upconv_c1_fwd(const void* x_bf16, const float* w, const float* bias, int B, int H, int W, int act,
                                   float* out, livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && H >= 2 && W >= 2, "upconv_c1_fwd: bad sizes");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(x_bf16 && w && bias && out, "upconv_c1_fwd: null pointer");
  LIVAE_CHECK_ARG((((uintptr_t)x_bf16) & 15) == 0 && B <= 65535, "upconv_c1_fwd: 16-byte alignment, B <= 65535");
  if (int e = require_sm100()) return e;
  const size_t smem = (size_t)kXP * kPitch + (size_t)kXP * kVP * sizeof(float);
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaError_t ce = cudaFuncSetAttribute(upconv_c1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) { set_error("upconv_c1_fwd: cannot set %zu B of shared memory", smem); return (int)ce; }
  }
  dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, B);
  upconv_c1_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_bf16, w, bias, H, W, act, out);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

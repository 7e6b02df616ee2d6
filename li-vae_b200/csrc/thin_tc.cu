// Thin 1-channel layers of the rVAE on tcgen05 tensor cores.
//
// The three layers that touch a 1-channel image (STN conv1 1->16 5x5 +ReLU +pool, model.py:204-206;
// encoder c1 1->32 4x4 s2 +ReLU, model.py:290; decoder d4 32->1 3x3 +sigmoid, model.py:371-372) have
// GEMM shapes with K <= 25 or N = 1.  As SIMT kernels they are bound by FMA/LDS issue (one shared load
// per FMA), 5-10x above their HBM time.  Here every one of their forward / data-gradient /
// weight-gradient passes is a 128-row UMMA whose *thin* operand is assembled in shared memory by the
// CTA's threads (written through the hardware swizzle map, fence.proxy.async, then tcgen05.mma) while
// the *wide* bf16 NHWC tensor streams through TMA untouched:
//
//   A. conv1c_tc_kernel   1 -> C forward (STN conv1 + pool, encoder c1, d4 data gradient):
//        D[128 px, C] = im2col(img)[128 px, K=taps] * W[C, taps]^T, A rows gathered from a bf16 copy of
//        the zero-padded image held in shared memory (32-bit loads + funnel shifts, no per-tap loop).
//   B. col2im_tc_kernel   C -> 1 (d4 forward, encoder c1 data gradient):
//        P[128 px, taps] = X[128 px, C] * W[taps, C]^T with X = 128 consecutive NHWC pixels (one TMA
//        box), then out[o] = sum_taps P[o + shift(tap)][tap] from a ring buffer of P rows in shared
//        memory (col2im).  Weights are split hi + lo bf16 (two column groups), so results are fp32-exact
//        in the weights.
//   C. tap_wgrad_tc_kernel  weight gradients of all three layers:
//        D[tap, C] += A[tap, 64 px] * G[64 px, C] with A = tap-shifted copies of the 1-channel image
//        (hi and lo bf16 rows -> fp32-exact), G = the wide tensor as an MN-major TMA box; the
//        accumulator stays in TMEM for the CTA's whole life, one atomicAdd per weight per CTA at the end.
//        A row of ones gives the bias gradient for free.
#include "tc_common.cuh"
#include "thin_tc.cuh"
#include <string.h>

namespace livae {
namespace tc {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }
// {hi | lo << 16}: v ~= bf16(hi) + bf16(lo), relative error ~2^-17
__device__ __forceinline__ uint32_t split_hi_lo(float v) {
  const uint32_t h = bf16_bits(v);
  const uint32_t l = bf16_bits(v - bf16_bits_to_float(h));
  return h | (l << 16);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 variant of split_hi_lo
__device__ __forceinline__ uint32_t split_hi_lo_f16(float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
}
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u);
}

// =============================================================================================
// A. 1 -> C forward
// =============================================================================================
template <int KIND> struct FwdCfg;
// STN conv1: 5x5 p2, ReLU, 2x2 max-pool.  One A row per POOLED pixel = its 6x6 input window (K = 36 -> 48, three
// K steps of a 128-byte row); the N = 64 columns are (pool position, channel) with the 5x5 filter placed at the
// position's offset inside the window, so the A row is the window exactly as it lies in the staged image (18 words,
// no shifts) and one tile is 6 MMAs instead of 16.
// F16: operands in fp16 instead of bf16 -- the image is in [0, 1] (min-max normalised patches, data.py:553-558)
// and the filter weights are O(1), so fp16's 11-bit mantissa costs nothing in range and rounds the image 8x finer.
// WLO: a second weight tile holding the rounding residual (w ~= hi + lo, exact to ~2^-17).  Off for the two layers
// whose output is stored as bf16 (8-bit mantissa): fp16 weights (2^-11) sit below that rounding and below the fp16
// image they multiply, and the residual tile doubled the MMA count and cost the fourth A slot its shared memory.
template <> struct FwdCfg<0> { enum { C = 16, KS = 5, S = 1, PAD = 2, KW2 = 6, KPAD = 64, POOL = 1, FLIP = 0, RELU = 1, F16 = 1, WLO = 0 }; };
// encoder c1: 4x4 s2 p1, ReLU.  K = 16
template <> struct FwdCfg<1> { enum { C = 32, KS = 4, S = 2, PAD = 1, KW2 = 4, KPAD = 16, POOL = 0, FLIP = 0, RELU = 1, F16 = 1, WLO = 0 }; };
// data gradient of decoder d4 (3x3 p0): full correlation with the flipped filter, pad 2.  K (ky, kx padded to 4): 12 -> 16
// (bf16: the image here is a gradient of arbitrary scale)
template <> struct FwdCfg<2> { enum { C = 32, KS = 3, S = 1, PAD = 2, KW2 = 4, KPAD = 16, POOL = 0, FLIP = 1, RELU = 0, F16 = 0, WLO = 1 }; };

struct FwdParams {
  const float* img; const float* w; const float* bias;
  int B, H, W;          // image
  int Ho, Wo;           // convolution output grid
  int Hs, pitch;        // staged (zero padded) image: rows, bf16 elements per row (even)
  void* out; uint8_t* idx;
};

// Warp roles (416 threads): warp 0 = MMA issuer, warps 1-8 = builders (one A row per thread) in two groups that take
// alternate tiles, warps 9-12 = epilogue.  Cycle accounting of the round-2 kernel (one builder group): the issuer
// waited for an A tile 53 % of the time and the epilogue for an accumulator 58 %; a builder's chain per tile (shared
// loads -> wait -> swizzled stores -> proxy fence -> arrive) is ~1200 cycles of latency for ~100 instructions, and
// staging the next image (scattered 4-byte loads with a division per word) took a third of the kernel.  Hence two
// builder groups, and staging by all eight builder warps as coalesced row loads with 12 requests in flight per
// thread from an image the issuer warp has already pulled into L2.
static constexpr int kFwdThreads = 416;
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(kFwdThreads, 2) conv1c_tc_kernel(const FwdParams p) {
  using Cfg = FwdCfg<KIND>;
  constexpr int C = Cfg::C, KS = Cfg::KS, PAD = Cfg::PAD, KW2 = Cfg::KW2, KPAD = Cfg::KPAD;
  constexpr bool POOL = Cfg::POOL != 0;
  constexpr uint32_t RB = KPAD * 2;                 // bytes per A / W row (128, 64 or 32)
  constexpr int NCOL = POOL ? 4 * C : C;            // TMEM columns per accumulator = UMMA N
  constexpr int KSTEPS = POOL ? 3 : KPAD / 16;      // pooled layer: K = 36 of the 64-element row
  constexpr uint32_t A_BUF = 128u * RB;
  constexpr uint32_t W_HALF = (uint32_t)NCOL * RB;  // bytes of one weight tile (hi or lo part)
  constexpr int NS = 4;                             // A-tile slots (builders -> MMA); even: a slot keeps its builder group
  constexpr int NT = 4;                             // TMEM accumulators (MMA -> epilogue)
  constexpr int NHL = Cfg::WLO ? 2 : 1;
  constexpr uint32_t TMEM_COLS = NT * NCOL;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[NS], a_empty[NS], t_full[NT], t_empty[NT];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[32];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + NS * A_BUF;                    // weight tile(s): rounded weights [, rounding residual]
  uint32_t* simg32 = reinterpret_cast<uint32_t*>(sW + NHL * W_HALF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pitch2 = p.pitch >> 1;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NT; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_s, TMEM_COLS); tmem_relinquish(); }
  // weight tile W[n][k], k = ky * KW2 + kx, zero in the padding slots; pooled layer: n = (pool position, c) and the
  // filter sits at the position's offset (dy, dx) inside the 6x6 window
  for (int i = tid; i < NCOL * KPAD; i += kFwdThreads) {
    const int n = i / KPAD, k = i % KPAD;
    const int c = POOL ? (n & (C - 1)) : n;
    const int ky = k / KW2 - (POOL ? (n / C) >> 1 : 0), kx = k % KW2 - (POOL ? (n / C) & 1 : 0);
    float v = 0.f;
    if (k < KW2 * KW2 && ky >= 0 && ky < KS && kx >= 0 && kx < KS) {
      const int t = ky * KS + kx;
      v = p.w[c * KS * KS + (Cfg::FLIP ? KS * KS - 1 - t : t)];
    }
    const uint32_t hl = Cfg::F16 ? split_hi_lo_f16(v) : split_hi_lo(v);
    const uint32_t o = swz_off((uint32_t)n, (uint32_t)k >> 3, RB) + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(sW + o) = (uint16_t)(hl & 0xffffu);
    if (Cfg::WLO) *reinterpret_cast<uint16_t*>(sW + W_HALF + o) = (uint16_t)(hl >> 16);
  }
  if (POOL && warp >= 1 && warp <= 4) {      // K elements 40..47 of every A row: read by the third K step, never rewritten
    for (int s = 0; s < NS; ++s)
      *reinterpret_cast<uint4*>(sA + s * A_BUF + swz_off((uint32_t)(tid - 32), 5u, RB)) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 32) sbias[tid] = (p.bias && tid < C) ? p.bias[tid] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int Wg = POOL ? (p.Wo >> 1) : p.Wo;            // grid the 128-row tiles walk over
  const int npx = POOL ? (p.Ho >> 1) * Wg : p.Ho * Wg;
  const int ntiles = (npx + 127) >> 7;

  if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues)
    {
      const bool leader = elect_one();
      const uint32_t idesc = Cfg::F16 ? make_idesc_f16(128, NCOL, 0, 0) : make_idesc_bf16(128, NCOL, 0, 0);
      const uint32_t lt = RB == 128 ? 2u : RB == 64 ? 4u : 6u;
      const uint64_t ad0 = make_smem_desc(smem_u32(sA), 16u, 8u * RB, lt);
      const uint64_t bd0 = make_smem_desc(smem_u32(sW), 16u, 8u * RB, lt);
      uint32_t it = 0;
      for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
        if (img + (int)gridDim.x < p.B) {      // the builders stage the next image from L2 instead of HBM
          const char* nx = reinterpret_cast<const char*>(p.img + (int64_t)(img + (int)gridDim.x) * p.H * p.W);
          for (int l = lane; l < p.H * p.W * 4 / 128; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + l * 128));
        }
        for (int tile = 0; tile < ntiles; ++tile, ++it) {
          const uint32_t s = it % NS, st = it % NT;
          mbar_wait(&t_empty[st], ((it / NT) & 1u) ^ 1u);
          mbar_wait(&a_full[s], (it / NS) & 1u);
          tc_fence_after();
#pragma unroll
          for (int hl = 0; hl < NHL; ++hl)                 // WLO: D = A * Whi^T + A * Wlo^T
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
              if (leader) umma_f16(tmem_base + st * NCOL, ad0 + ((s * A_BUF + ks * 32u) >> 4),
                                   bd0 + ((hl * W_HALF + ks * 32u) >> 4), idesc, (ks | hl) > 0 ? 1u : 0u);
          if (leader) { umma_commit(&a_empty[s]); umma_commit(&t_full[st]); }
        }
      }
    }
  } else if (warp <= 8) {
    // ------------------------------------------------------------------ builders
    const int bw = warp - 1;
    const uint32_t grp = (uint32_t)bw >> 2;               // tiles with it % 2 == grp
    const int bt = (bw & 3) * 32 + lane;                  // A row of this thread
    uint32_t it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
      asm volatile("bar.sync 1, 256;" ::: "memory");     // every builder is done reading the previous image
      const float* im = p.img + (int64_t)img * p.H * p.W;
      // stage the zero-padded image as fp16 / bf16 pairs: one warp per row, consecutive lanes = consecutive words, the
      // loads of four rows (12 per thread) are issued before the first use
      for (int r0 = bw; r0 < p.Hs; r0 += 8 * 4) {
        float va[4][3], vb[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + 8 * u, iy = r - PAD;
          const bool rowok = r < p.Hs && iy >= 0 && iy < p.H;
          const float* rp = im + (rowok ? iy : 0) * p.W;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int c = lane + 32 * k, ix = 2 * c - PAD;
            float a = 0.f, b = 0.f;
            if (rowok && c < pitch2) {
              if ((PAD & 1) == 0) {                       // even padding: the pair is one aligned 8-byte load, in or out as a whole
                if (ix >= 0 && ix < p.W) { const float2 t = __ldg(reinterpret_cast<const float2*>(rp + ix)); a = t.x; b = t.y; }
              } else {
                if (ix >= 0 && ix < p.W) a = __ldg(rp + ix);
                if (ix + 1 >= 0 && ix + 1 < p.W) b = __ldg(rp + ix + 1);
              }
            }
            va[u][k] = a; vb[u][k] = b;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + 8 * u;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int c = lane + 32 * k;
            if (r < p.Hs && c < pitch2)
              simg32[r * pitch2 + c] = Cfg::F16 ? pack_f16x2(va[u][k], vb[u][k]) : pack_bf16x2(va[u][k], vb[u][k]);
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int pp = bt;                                        // this thread's pixel of the current tile
      for (int tile = 0; tile < ntiles; ++tile, ++it, pp += 128) {
        if ((it & 1u) != grp) continue;
        const uint32_t s = it % NS;
        uint8_t* a_buf = sA + s * A_BUF;
        const int pc = pp < npx ? pp : npx - 1;           // rows past the end: any finite data
        const int gy = pc / Wg, gx = pc - gy * Wg;
        if (KIND == 0) {
          uint32_t wd[6][3];
          const uint32_t* base = simg32 + (2 * gy) * pitch2 + gx;
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) wd[i][j] = base[i * pitch2 + j];
          mbar_wait(&a_empty[s], ((it / NS) & 1u) ^ 1u);
          const uint32_t* k = &wd[0][0];                     // the window IS the A row: 18 words, K index u * 6 + v
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(a_buf + swz_off((uint32_t)bt, (uint32_t)j, RB)) =
                make_uint4(k[4 * j], k[4 * j + 1], k[4 * j + 2], k[4 * j + 3]);
          *reinterpret_cast<uint4*>(a_buf + swz_off((uint32_t)bt, 4u, RB)) = make_uint4(k[16], k[17], 0u, 0u);
        } else if (KIND == 1) {
          const uint32_t* base = simg32 + (2 * gy) * pitch2 + gx;
          uint32_t k[8];
#pragma unroll
          for (int ky = 0; ky < 4; ++ky) { k[2 * ky] = base[ky * pitch2]; k[2 * ky + 1] = base[ky * pitch2 + 1]; }
          mbar_wait(&a_empty[s], ((it / NS) & 1u) ^ 1u);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(a_buf + swz_off((uint32_t)bt, (uint32_t)j, RB)) =
                make_uint4(k[4 * j], k[4 * j + 1], k[4 * j + 2], k[4 * j + 3]);
        } else {
          const uint32_t* base = simg32 + gy * pitch2 + (gx >> 1);
          const uint32_t sh = (uint32_t)(gx & 1) * 16u;
          uint32_t k[8];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t l0 = base[ky * pitch2], l1 = base[ky * pitch2 + 1], l2 = base[ky * pitch2 + 2];
            k[2 * ky] = __funnelshift_r(l0, l1, sh);
            k[2 * ky + 1] = __funnelshift_r(l1, l2, sh);
          }
          k[6] = 0u; k[7] = 0u;
          mbar_wait(&a_empty[s], ((it / NS) & 1u) ^ 1u);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            *reinterpret_cast<uint4*>(a_buf + swz_off((uint32_t)bt, (uint32_t)j, RB)) =
                make_uint4(k[4 * j], k[4 * j + 1], k[4 * j + 2], k[4 * j + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                               // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    uint32_t it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x)
      for (int tile = 0; tile < ntiles; ++tile, ++it) {
        const uint32_t s = it % NT;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * NCOL;
        const int pp = tile * 128 + row;
        const bool valid = pp < npx;
        mbar_wait(&t_full[s], (it / NT) & 1u);
        tc_fence_after();
        if (POOL) {
          // two passes of 8 channels (4 pool positions x 8 columns each): 32 live accumulator registers
          uint32_t ow[8]; uint32_t iw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int c0 = 0; c0 < 16; c0 += 8) {
            uint32_t v[4][8];
#pragma unroll
            for (int sb = 0; sb < 4; ++sb) tmem_ld8(taddr + (uint32_t)(sb * 16 + c0), v[sb]);
            tmem_ld_wait();
            if (c0 == 8) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[s]);
            }
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              float best[2]; uint32_t bi[2];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                // max-pool commutes with the shared bias and the ReLU: pool the raw sums (first maximum in scan
                // order), then bias + ReLU once.  Where all four ReLU outputs tie at zero the reference reports
                // position 0; the gradient there is removed by the ReLU mask either way.
                const float v0 = __uint_as_float(v[0][c + j]), v1 = __uint_as_float(v[1][c + j]);
                const float v2 = __uint_as_float(v[2][c + j]), v3 = __uint_as_float(v[3][c + j]);
                const float m = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                bi[j] = v0 == m ? 0u : v1 == m ? 1u : v2 == m ? 2u : 3u;
                best[j] = fmaxf(m + sbias[c0 + c + j], 0.f);
              }
              ow[(c0 + c) >> 1] = pack_bf16x2(best[0], best[1]);
              iw[(c0 + c) >> 2] |= (bi[0] << ((c & 3) * 8)) | (bi[1] << (((c & 3) + 1) * 8));
            }
          }
          if (valid) {
            const int64_t o = (int64_t)img * npx + pp;
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o * 16);
            op[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            op[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            *reinterpret_cast<uint4*>(p.idx + o * 16) = make_uint4(iw[0], iw[1], iw[2], iw[3]);
          }
        } else {
          uint32_t v[2][16];
          tmem_ld16(taddr, v[0]);
          tmem_ld16(taddr + 16u, v[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[s]);
          if (valid) {
            uint32_t ow[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              float a0 = __uint_as_float(v[c >> 4][c & 15]) + sbias[c];
              float a1 = __uint_as_float(v[c >> 4][(c & 15) + 1]) + sbias[c + 1];
              if (Cfg::RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
              ow[c >> 1] = pack_bf16x2(a0, a1);
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ((int64_t)img * npx + pp) * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_uint4(ow[4 * j], ow[4 * j + 1], ow[4 * j + 2], ow[4 * j + 3]);
          }
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int KIND>
static int launch_fwd(const float* img, const float* w, const float* bias, int B, int H, int W, int Ho, int Wo,
                      void* out, uint8_t* idx, cudaStream_t st) {
  using Cfg = FwdCfg<KIND>;
  FwdParams p;
  p.img = img; p.w = w; p.bias = bias; p.B = B; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo;
  p.Hs = (Ho - 1) * Cfg::S + Cfg::KS;
  p.pitch = ((Wo - 1) * Cfg::S + Cfg::KW2 + 2 + 1) & ~1;
  p.out = out; p.idx = idx;
  const int ns = 4;
  const size_t a_buf = (size_t)128 * Cfg::KPAD * 2;
  const size_t wbytes = (Cfg::WLO ? 2 : 1) * (size_t)(Cfg::POOL ? 4 : 1) * Cfg::C * Cfg::KPAD * 2;
  const size_t smem = 1024 + ns * a_buf + wbytes + (size_t)p.Hs * p.pitch * 2;
  if (smem > 110 * 1024) return 1;
  static OncePerDevice attr;
  if (attr.first()) { cudaFuncSetAttribute(conv1c_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);}
  int occ = (int)((227 * 1024) / (smem + 1024));
  const int occ_tmem = 512 / (4 * (Cfg::POOL ? 4 : 1) * Cfg::C);
  if (occ > occ_tmem) occ = occ_tmem;
  if (occ > 2) occ = 2;
  if (occ < 1) occ = 1;
  int grid = kNumSMs * occ;
  if (grid > B) grid = B;
  conv1c_tc_kernel<KIND><<<grid, kFwdThreads, smem, st>>>(p);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

int thin_tc_conv1c_fwd(int kind, const float* img, const float* w, const float* bias, int B, int H, int W,
                       void* out_bf16, uint8_t* pool_idx, cudaStream_t st) {
  if ((H & 1) || (W & 1) || H < 4 || W < 4) return 1;
  if ((((uintptr_t)out_bf16 | (uintptr_t)pool_idx) & 15) != 0) return 1;
  if (((uintptr_t)img & 7) != 0) return 1;       // the image rows are staged with 8-byte loads
  if (kind == 0) return launch_fwd<0>(img, w, bias, B, H, W, H, W, out_bf16, pool_idx, st);
  if (kind == 1) return launch_fwd<1>(img, w, bias, B, H, W, H / 2, W / 2, out_bf16, nullptr, st);
  return launch_fwd<2>(img, w, nullptr, B, H, W, H + 2, W + 2, out_bf16, nullptr, st);
}

// =============================================================================================
// B. C -> 1 through GEMM + col2im
// =============================================================================================
struct C2iParams {
  int B, Hin, Win, npx, ntiles;
  int Ho, Wo;
  int ring;             // power of two
  int act;
  const float* w; const float* bias; float* out;
};

static constexpr int kC2iStages = 4;

// KIND 0: decoder d4 forward, out[oy,ox] = act(bias + sum_{ky,kx,c} x[oy+ky, ox+kx, c] w[c][ky][kx]), 3x3
// KIND 1: encoder c1 data gradient, gimg[iy,ix] = sum_{ky,kx,c} g[(iy+1-ky)/2, (ix+1-kx)/2, c] w[c][ky][kx], 4x4 s2 p1
template <int KIND>
__global__ void __launch_bounds__(192) col2im_tc_kernel(const __grid_constant__ CUtensorMap tmX, const C2iParams p) {
  constexpr int T = KIND == 0 ? 9 : 16;
  constexpr uint32_t A_BYTES = 128u * 64u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t fullA[kC2iStages], emptyA[kC2iStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + kC2iStages * A_BYTES;            // [32 rows: 16 hi taps, 16 lo taps][32 ch] bf16, 64-byte rows
  float* ring = reinterpret_cast<float*>(sW + 2048);    // [T][ring]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RING = p.ring, RM = p.ring - 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    for (int s = 0; s < kC2iStages; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 32 * 32; i += 192) {
    const int n = i >> 5, c = i & 31, t = n & 15;
    const float wv = t < T ? p.w[c * T + t] : 0.f;
    const uint32_t hl = split_hi_lo(wv);
    *reinterpret_cast<uint16_t*>(sW + swz_off((uint32_t)n, (uint32_t)c >> 3, 64u) + (c & 7) * 2) =
        (uint16_t)(n < 16 ? (hl & 0xffffu) : (hl >> 16));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int img = blockIdx.x; img < p.B; img += gridDim.x)
        for (int tile = 0; tile < p.ntiles; ++tile, ++it) {
          const uint32_t s = it % kC2iStages;
          mbar_wait(&emptyA[s], ((it / kC2iStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&fullA[s], A_BYTES);
          tma_load_2d(sA + s * A_BYTES, &tmX, &fullA[s], 0, img * p.npx + tile * 128);
        }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();      // whole-warp loop, elected lane issues
      const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
      const uint32_t w_addr = smem_u32(sW);
      uint32_t it = 0;
      for (int img = blockIdx.x; img < p.B; img += gridDim.x)
        for (int tile = 0; tile < p.ntiles; ++tile, ++it) {
          const uint32_t s = it % kC2iStages, acc = it & 1u;
          mbar_wait(&tempty[acc], ((it >> 1) & 1u) ^ 1u);
          mbar_wait(&fullA[s], (it / kC2iStages) & 1u);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * A_BYTES);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = make_smem_desc(a_addr + ks * 32u, 16u, 512u, 4u);
            const uint64_t bd = make_smem_desc(w_addr + ks * 32u, 16u, 512u, 4u);
            if (leader) umma_f16(tmem_base + acc * 32u, ad, bd, idesc, ks > 0 ? 1u : 0u);
          }
          if (leader) { umma_commit(&emptyA[s]); umma_commit(&tfull[acc]); }
        }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const float b0 = p.bias ? p.bias[0] : 0.f;
    uint32_t it = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
      int lim_prev = 0;
      for (int tile = 0; tile < p.ntiles; ++tile, ++it) {
        const uint32_t acc = it & 1u;
        mbar_wait(&tfull[acc], (it >> 1) & 1u);
        tc_fence_after();
        uint32_t v0[16], v1[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32u;
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 16u, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const int pin = tile * 128 + r;
        if (pin < p.npx) {
#pragma unroll
          for (int t = 0; t < T; ++t) ring[t * RING + (pin & RM)] = __uint_as_float(v0[t]) + __uint_as_float(v1[t]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (KIND == 0) {
          const int D = 2 * p.Win + 2;
          const int qo = tile * 128 - D + r;
          if (qo >= 0) {
            const int oy = qo / p.Win, ox = qo - oy * p.Win;
            if (oy < p.Ho && ox < p.Wo) {
              float a = b0;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) a += ring[(ky * 3 + kx) * RING + ((qo + ky * p.Win + kx) & RM)];
              if (p.act == LIVAE_ACT_SIGMOID) a = 1.f / (1.f + __expf(-a));
              else if (p.act == LIVAE_ACT_RELU) a = fmaxf(a, 0.f);
              p.out[((int64_t)img * p.Ho + oy) * p.Wo + ox] = a;
            }
          }
        } else {
          // input rows complete after this tile -> output rows that can be finished
          int rc = (tile + 1) * 128 / p.Win;
          if (tile == p.ntiles - 1 || rc > p.Hin) rc = p.Hin;
          int lim = rc >= p.Hin ? p.Ho : (2 * rc - 1 > 0 ? 2 * rc - 1 : 0);
          const int nout = (lim - lim_prev) * p.Wo;
          for (int j = r; j < nout; j += 128) {
            const int dy = j / p.Wo;
            const int iy = lim_prev + dy, ix = j - dy * p.Wo;
            float a = 0.f;
#pragma unroll
            for (int ay = 0; ay < 2; ++ay) {
              const int ky = ((iy + 1) & 1) + 2 * ay;
              const int ty = iy + 1 - ky;
              if (ty < 0 || (ty >> 1) >= p.Hin) continue;
#pragma unroll
              for (int ax = 0; ax < 2; ++ax) {
                const int kx = ((ix + 1) & 1) + 2 * ax;
                const int tx = ix + 1 - kx;
                if (tx < 0 || (tx >> 1) >= p.Win) continue;
                a += ring[(ky * 4 + kx) * RING + (((ty >> 1) * p.Win + (tx >> 1)) & RM)];
              }
            }
            p.out[((int64_t)img * p.Ho + iy) * p.Wo + ix] = a;
          }
          lim_prev = lim;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");    // ring reads of this image done before the next one writes
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

int thin_tc_col2im(int kind, const void* x, const float* w, const float* bias, int B, int Hin, int Win, int act,
                   float* out, cudaStream_t st) {
  if (((uintptr_t)x & 15) != 0) return 1;
  C2iParams p;
  p.B = B; p.Hin = Hin; p.Win = Win; p.npx = Hin * Win; p.ntiles = (p.npx + 127) / 128;
  if (kind == 0) { p.Ho = Hin - 2; p.Wo = Win - 2; } else { p.Ho = 2 * Hin; p.Wo = 2 * Win; }
  if ((int64_t)B * p.npx >= (1ll << 31)) return 1;
  int ring = 256;
  while (ring < 256 + 3 * Win + 2) ring <<= 1;
  p.ring = ring; p.act = act; p.w = w; p.bias = bias; p.out = out;
  const int T = kind == 0 ? 9 : 16;
  const size_t smem = 1024 + (size_t)kC2iStages * 128 * 64 + 2048 + (size_t)T * ring * 4;
  if (smem > 110 * 1024) return 1;
  CUtensorMap tmX;
  {
    uint64_t dims[2] = {32, (uint64_t)B * p.npx};
    uint64_t str[1] = {64};
    uint32_t box[2] = {32, 128};
    if (int e = make_tmap_bf16(&tmX, x, 2, dims, str, box, nullptr, 64)) return e;
  }
  static OncePerDevice attr;
  if (attr.first()) {
    cudaFuncSetAttribute(col2im_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaFuncSetAttribute(col2im_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  }
  int occ = (int)((227 * 1024) / (smem + 1024));
  if (occ > 4) occ = 4;
  if (occ < 1) occ = 1;
  int grid = kNumSMs * occ;
  if (grid > B) grid = B;
  if (kind == 0) col2im_tc_kernel<0><<<grid, 192, smem, st>>>(tmX, p);
  else col2im_tc_kernel<1><<<grid, 192, smem, st>>>(tmX, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// C. weight gradients
// =============================================================================================
template <int KIND> struct WgCfg;
// STN conv1: gw[c][ky][kx] = sum img[y+ky-2, x+kx-2] * gfull[y,x,c]; gfull = pooled gradient routed by idx
template <> struct WgCfg<0> { enum { C = 16, KS = 5, S = 1, PAD = 2, FLIP = 0, BTMA = 0, ONES = 1 }; };
// encoder c1: gw[c][ky][kx] = sum img[2oy-1+ky, 2ox-1+kx] * g[oy,ox,c]
template <> struct WgCfg<1> { enum { C = 32, KS = 4, S = 2, PAD = 1, FLIP = 0, BTMA = 1, ONES = 1 }; };
// decoder d4: gw[c][ky][kx] = sum_{iy,ix} x[iy,ix,c] * gpre[iy-ky, ix-kx]   (src = gpre, wide = x)
template <> struct WgCfg<2> { enum { C = 32, KS = 3, S = 1, PAD = 2, FLIP = 1, BTMA = 1, ONES = 0 }; };

struct WgParams {
  const float* src;     // 1-channel fp32 [B][H][W]
  int B, H, W;
  int Hg, Wg, npx, nst; // reduction grid, pixels and 64-pixel stages per image
  int Hs, pitch;        // staged source (zero padded): rows, 32-bit words per row
  const void* gp; const uint8_t* idx;   // KIND 0: pooled gradient bf16 [B][H/2][W/2][16] and argmax
  float* gw; float* gb;
};

// Ring geometry.  Every mbarrier is waited on by ONE owner that visits its phases in order (a parity
// wait two phases ahead of the barrier would pass spuriously), so the owners' strides divide the ring
// sizes: issuer = g % 3 owns A slots {i, i+3} and B slots {i, i+3, i+6, i+9}; team = g % NTEAMS (2 or 3)
// owns A slots of the same residue.
static constexpr int kWgA = 6, kWgB = 12, kWgIssuers = 3;
template <int KIND> struct WgLayout {
  using Cfg = WgCfg<KIND>;
  static constexpr int TASKS = (Cfg::KS * Cfg::KS + Cfg::ONES) * 8;
  static constexpr int NAW = (TASKS + 31) / 32;           // A-builder warps per team
  static constexpr int NBWT = Cfg::BTMA ? 0 : 4;          // un-pool warps per team
  static constexpr int TEAMW = NAW + NBWT;
  static constexpr int NTEAMS = KIND == 0 ? 2 : 3;
  static constexpr int BUILDW = NTEAMS * TEAMW;           // 22 / 15 / 9 builder warps
  static constexpr int FIRSTB = 1 + kWgIssuers;           // first builder warp
  static constexpr int THREADS = 32 * (FIRSTB + BUILDW);
};

// 1-D bulk copy global -> shared, completion on an mbarrier (size and addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One CTA per SM: warp 0 = producer of the wide operand (TMA boxes / bulk copies, kWgB slots = 40 KB in
// flight: with one CTA per SM the bytes in flight decide the HBM throughput), warps 1-3 = MMA issuers
// (stage g belongs to issuer g % 3, each with its own TMEM accumulator: waiting on two mbarriers, issuing 4
// UMMAs and committing costs one thread ~800 cycles of pure latency per 64-pixel stage, three of them keep
// the tensor pipe busy), the remaining warps = builders.  A stage's tap
// rows are (T + 1) * 8 independent 16-byte chunk tasks, i.e. only 3-7 warps of work and a long latency
// chain (shared loads -> byte permutes -> swizzled stores -> proxy fence -> mbarrier), so the builder
// warps form TEAMS that work on different stages concurrently; every thread owns one fixed (tap, chunk)
// task and walks its pixel coordinates incrementally (no divisions in the stage loop).
template <int KIND>
__global__ void __launch_bounds__(WgLayout<KIND>::THREADS, 1) tap_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmB, const WgParams p) {
  using Cfg = WgCfg<KIND>;
  constexpr int C = Cfg::C, KS = Cfg::KS, S = Cfg::S, PAD = Cfg::PAD, T = KS * KS;
  constexpr bool BTMA = Cfg::BTMA != 0;
  constexpr uint32_t RBB = C * 2;                       // bytes per pixel row of the wide operand
  constexpr uint32_t A_BYTES = 128u * 128u;             // [128 rows][64 px] bf16, 128-byte swizzled rows
  constexpr uint32_t B_BYTES = 64u * RBB;               // MN-major tile the MMA reads
  constexpr uint32_t R_SLOT = BTMA ? 4096u : 2048u;     // ring slot: the TMA box itself, or pooled gradient (1 KB) + argmax (512 B)
  constexpr uint32_t U_SLOT = 2048u;                    // KIND 0: un-pooled tile, one per A slot
  using Lay = WgLayout<KIND>;
  constexpr int TASKS = Lay::TASKS, NAW = Lay::NAW, NBWT = Lay::NBWT, TEAMW = Lay::TEAMW, NTEAMS = Lay::NTEAMS;
  constexpr int NBUILD = 32 * Lay::BUILDW, FIRSTB = Lay::FIRSTB;
  static_assert(kWgA % NTEAMS == 0 && kWgA % kWgIssuers == 0 && kWgB % kWgIssuers == 0, "ring ownership");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kWgA], a_empty[kWgA], b_full[kWgB], b_empty[kWgB], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sR = smem + kWgA * A_BYTES;
  uint8_t* sU = sR + kWgB * R_SLOT;
  uint32_t* simg = reinterpret_cast<uint32_t*>(sU + (BTMA ? 0u : kWgA * U_SLOT));   // {hi | lo << 16} per pixel
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if (BTMA) prefetch_tmap(&tmB);
    for (int s = 0; s < kWgA; ++s) { mbar_init(&a_full[s], TEAMW); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kWgB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], BTMA ? 1 : NBWT); }
    mbar_init(&accum_bar, kWgIssuers);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int Wp = p.Wg >> 1;
  const int pool_cnt = p.Wg >= 64 ? 32 : 16;              // KIND 0: pooled pixels under one 64-pixel stage

  int nimg = 0;
  for (int img = blockIdx.x; img < p.B; img += gridDim.x) ++nimg;
  const uint32_t total = (uint32_t)nimg * (uint32_t)p.nst;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer of the wide operand
    if (lane == 0) {
      uint32_t g = 0;
      for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
        int y0 = 0, x0 = 0;                               // KIND 0: first pixel of the stage
        for (int stg = 0; stg < p.nst; ++stg, ++g) {
          const uint32_t b = g % kWgB;
          mbar_wait(&b_empty[b], ((g / kWgB) & 1u) ^ 1u);
          if (BTMA) {
            mbar_arrive_expect_tx(&b_full[b], B_BYTES);
            tma_load_2d(sR + b * R_SLOT, &tmB, &b_full[b], 0, img * p.npx + stg * 64);
          } else {
            const int64_t ps = (int64_t)img * (p.Hg >> 1) * Wp + (y0 >> 1) * Wp + (x0 >> 1);
            mbar_arrive_expect_tx(&b_full[b], (uint32_t)pool_cnt * 48u);
            bulk_load_1d(sR + b * R_SLOT, reinterpret_cast<const __nv_bfloat16*>(p.gp) + ps * 16, (uint32_t)pool_cnt * 32u, &b_full[b]);
            bulk_load_1d(sR + b * R_SLOT + 1024, p.idx + ps * 16, (uint32_t)pool_cnt * 16u, &b_full[b]);
            x0 += 64;
            while (x0 >= p.Wg) { x0 -= p.Wg; ++y0; }
          }
        }
      }
    }
  } else if (warp < FIRSTB) {
    // ------------------------------------------------------------------ MMA issuers (whole warps, elected lanes issue)
    {
      const bool leader = elect_one();
      const uint32_t wi = (uint32_t)(warp - 1);
      const uint32_t idesc = make_idesc_bf16(128, C, 0, 1);   // A K-major, B MN-major
      const uint32_t ltb = RBB == 64 ? 4u : 6u;
      const uint64_t ad0 = make_smem_desc(smem_u32(sA), 16u, 1024u, 2u);
      const uint64_t bd0 = make_smem_desc(smem_u32(BTMA ? sR : sU), B_BYTES, 8u * RBB, ltb);
      const uint32_t d_addr = tmem_base + wi * 32u;
      uint32_t a = wi % kWgA, b = wi % kWgB, pa = 0u, pb = 0u, first = 1u;   // slots / phase parities of stage g
      for (uint32_t g = wi; g < total; g += kWgIssuers) {
        mbar_wait(&a_full[a], pa);
        if (BTMA) mbar_wait(&b_full[b], pb);
        tc_fence_after();
        const uint64_t ad = ad0 + ((a * A_BYTES) >> 4);
        const uint64_t bd = bd0 + ((BTMA ? b * R_SLOT : a * U_SLOT) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (leader) umma_f16(d_addr, ad + ((ks * 32u) >> 4), bd + ((ks * 16u * RBB) >> 4), idesc, (ks > 0 || !first) ? 1u : 0u);
        first = 0u;
        if (leader) umma_commit(&a_empty[a]);
        if (BTMA && leader) umma_commit(&b_empty[b]);
        a += kWgIssuers; if (a >= kWgA) { a -= kWgA; pa ^= 1u; }
        b += kWgIssuers; if (b >= kWgB) { b -= kWgB; pb ^= 1u; }
      }
      if (leader) umma_commit(&accum_bar);
    }
  } else {
    // ------------------------------------------------------------------ builders
    const int bw = warp - FIRSTB, bt = tid - 32 * FIRSTB;
    const int team = bw / TEAMW, rw = bw - team * TEAMW;
    const bool in_team = team < NTEAMS;
    const bool a_role = in_team && rw < NAW;
    const int task = rw * 32 + lane;                      // a_role: fixed (tap, chunk) of this thread
    const bool a_task = a_role && task < TASKS;
    const int t = task >> 3, ch = task & 7;
    const int ky = t / KS, kx = t - ky * KS;
    const int dy = Cfg::FLIP ? (KS - 1 - ky) : ky, dx = Cfg::FLIP ? (KS - 1 - kx) : kx;
    const uint32_t off_hi = swz_off((uint32_t)(t < T ? t : 64), (uint32_t)ch, 128u);
    const uint32_t off_lo = swz_off((uint32_t)(32 + (t < T ? t : 0)), (uint32_t)ch, 128u);
    const int ub = (rw - NAW) * 32 + lane;                // un-pool role: (pixel j, channel half h)
    const int uj = ub >> 1, uh = ub & 1;
    const uint32_t off_u = swz_off((uint32_t)uj, (uint32_t)uh, 32u);
    uint32_t gbase = 0;
    float bias_acc = 0.f;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, gbase += (uint32_t)p.nst) {
      asm volatile("bar.sync 1, %0;" ::"n"(NBUILD) : "memory");     // every builder is done reading the previous image
      const float* im = p.src + (int64_t)img * p.H * p.W;
      {
        const int n = p.Hs * p.pitch;
        for (int i0 = bt; i0 < n; i0 += NBUILD * 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = min(i0 + u * NBUILD, n - 1);
            const int r = i / p.pitch, c = i - r * p.pitch;
            const int iy = r - PAD, ix = c - PAD;
            const float a = __ldg(im + min(max(iy, 0), p.H - 1) * p.W + min(max(ix, 0), p.W - 1));
            v[u] = (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) ? a : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (i0 + u * NBUILD < n) {
              if (KIND == 2) bias_acc += v[u];
              simg[i0 + u * NBUILD] = split_hi_lo(v[u]);
            }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NBUILD) : "memory");
      if (!in_team) continue;
      // this team's stages of the image: those with global index g = gbase + stg congruent to team
      const int stg0 = (team + NTEAMS - (int)(gbase % NTEAMS)) % NTEAMS;
      // pixel coordinates of this thread's chunk (a_role) or pixel (un-pool role) and of the stage start
      int pbeg = stg0 * 64 + (a_role ? ch * 8 : uj);
      int y = pbeg / p.Wg, x = pbeg - y * p.Wg;
      int y0 = (stg0 * 64) / p.Wg, x0 = stg0 * 64 - y0 * p.Wg;
      for (int stg = stg0; stg < p.nst; stg += NTEAMS) {
        const uint32_t g = gbase + (uint32_t)stg;
        const uint32_t a = g % kWgA, b = g % kWgB;
        uint8_t* a_buf = sA + a * A_BYTES;
        if (a_role) {
          // ---- A: rows t (hi), 32 + t (lo), 64 (ones)
          uint32_t e8[8];
          if (a_task && t < T) {
            if (KIND != 2) {                                // Wg % 8 == 0: the chunk stays inside one row
              const uint32_t* src = simg + (y * S + dy) * p.pitch + x * S + dx;
              const bool ok = pbeg < p.npx;
#pragma unroll
              for (int e = 0; e < 8; ++e) e8[e] = ok ? src[e * S] : 0u;
            } else {                                        // the chunk may wrap once into the next row
              const uint32_t* src = simg + (y + dy) * p.pitch + x + dx;
              const int wrap = p.Wg - x;                    // elements e >= wrap belong to row y + 1
#pragma unroll
              for (int e = 0; e < 8; ++e)
                e8[e] = (pbeg + e < p.npx) ? src[e >= wrap ? e - p.Wg + p.pitch : e] : 0u;
            }
          }
          mbar_wait(&a_empty[a], ((g / kWgA) & 1u) ^ 1u);   // MMAs of stage g - kWgA have finished reading this slot
          if (a_task) {
            if (t < T) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hi[e] = __byte_perm(e8[2 * e], e8[2 * e + 1], 0x5410);
                lo[e] = __byte_perm(e8[2 * e], e8[2 * e + 1], 0x7632);
              }
              *reinterpret_cast<uint4*>(a_buf + off_hi) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(a_buf + off_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            } else {                                        // ones row: bias gradient
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = (pbeg + 2 * e < p.npx ? 0x3F80u : 0u) | (pbeg + 2 * e + 1 < p.npx ? 0x3F800000u : 0u);
              *reinterpret_cast<uint4*>(a_buf + off_hi) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        } else {
          // ---- B (KIND 0): un-pool the gradient into the MN-major tile [64 px][16 ch], 32-byte rows
          mbar_wait(&b_full[b], (g / kWgB) & 1u);
          uint4 o = make_uint4(0, 0, 0, 0);
          if (pbeg < p.npx) {
            const int local = (((y >> 1) - (y0 >> 1)) * Wp + ((x - x0) >> 1)) * 2 + uh;   // 16-byte units of the pooled slot
            const uint4 pg = *reinterpret_cast<const uint4*>(sR + b * R_SLOT + local * 16);
            const uint2 pi = *reinterpret_cast<const uint2*>(sR + b * R_SLOT + 1024 + local * 8);
            const uint32_t pos = (uint32_t)(((y & 1) << 1) | (x & 1));
            const uint32_t gv[4] = {pg.x, pg.y, pg.z, pg.w};
            uint32_t ov[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t iw = e < 2 ? pi.x : pi.y;
              const uint32_t i0 = (iw >> ((e & 1) * 16)) & 0xffu, i1 = (iw >> ((e & 1) * 16 + 8)) & 0xffu;
              ov[e] = (i0 == pos ? (gv[e] & 0xffffu) : 0u) | (i1 == pos ? (gv[e] & 0xffff0000u) : 0u);
            }
            o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          }
          mbar_wait(&a_empty[a], ((g / kWgA) & 1u) ^ 1u);   // the un-pooled tile shares the A slot's lifetime
          *reinterpret_cast<uint4*>(sU + a * U_SLOT + off_u) = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[a]);
          if (!a_role) mbar_arrive(&b_empty[b]);            // pooled slot consumed
        }
        // advance this thread's pixel by NTEAMS stages
        pbeg += 64 * NTEAMS;
        x += 64 * NTEAMS; while (x >= p.Wg) { x -= p.Wg; ++y; }
        x0 += 64 * NTEAMS; while (x0 >= p.Wg) { x0 -= p.Wg; ++y0; }
      }
    }
    if (KIND == 2 && p.gb) {
      const float sacc = warp_sum(bias_acc);
      if (lane == 0) atomicAdd(p.gb, sacc);
    }
    if (bw < 4) {   // four builder warps cover the four TMEM lane quadrants
      const int q = warp & 3;
      const int r = q * 32 + lane;
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
      float acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0.f;
      for (uint32_t wi = 0; wi < (uint32_t)kWgIssuers && wi < total; ++wi) {   // issuers that never ran left garbage
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + wi * 32u;
        uint32_t v[2][16];
        tmem_ld16(taddr, v[0]);
        if (C == 32) tmem_ld16(taddr + 16u, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += __uint_as_float(v[c >> 4][c & 15]);
      }
      const int tt = r < 32 ? r : r - 32;
      if (r < 64 && tt < T) {
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(p.gw + c * T + tt, acc[c]);
      } else if (Cfg::ONES && r == 64 && p.gb) {
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(p.gb + c, acc[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

template <int KIND>
static int launch_wg(const float* src, const void* big, const uint8_t* idx, int B, int H, int W, int Hg, int Wg,
                     float* gw, float* gb, cudaStream_t st) {
  using Cfg = WgCfg<KIND>;
  WgParams p;
  p.src = src; p.B = B; p.H = H; p.W = W; p.Hg = Hg; p.Wg = Wg; p.npx = Hg * Wg; p.nst = (p.npx + 63) / 64;
  p.Hs = (Hg - 1) * Cfg::S + Cfg::KS; p.pitch = (Wg - 1) * Cfg::S + Cfg::KS;
  p.gp = big; p.idx = idx; p.gw = gw; p.gb = gb;
  if ((int64_t)B * p.npx >= (1ll << 31)) return 1;
  const size_t ring = Cfg::BTMA ? (size_t)kWgB * 4096 : (size_t)kWgB * 2048 + (size_t)kWgA * 2048;
  const size_t smem = 1024 + (size_t)kWgA * 128 * 128 + ring + (size_t)p.Hs * p.pitch * 4;
  if (smem > 220 * 1024) return 1;
  CUtensorMap tmB;
  if (Cfg::BTMA) {
    uint64_t dims[2] = {(uint64_t)Cfg::C, (uint64_t)B * p.npx};
    uint64_t str[1] = {(uint64_t)Cfg::C * 2};
    uint32_t box[2] = {(uint32_t)Cfg::C, 64};
    if (int e = make_tmap_bf16(&tmB, big, 2, dims, str, box, nullptr, Cfg::C * 2)) return e;
  } else {
    memset(&tmB, 0, sizeof(tmB));
  }
  static OncePerDevice attr;
  if (attr.first()) { cudaFuncSetAttribute(tap_wgrad_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);}
  int grid = kNumSMs;
  if (grid > B) grid = B;
  tap_wgrad_tc_kernel<KIND><<<grid, WgLayout<KIND>::THREADS, smem, st>>>(tmB, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// C'. STN conv1 weight gradient, pixel-phase folded (W = 128)
// =============================================================================================
// gw[c][ky][kx] = sum_{b,y,x} imgpad[y+ky][x+kx] * G[y][x][c]  (G = un-pooled gradient, 16 channels).
// Writing x = 8*xg + ph turns one image ROW into a single K = 16 UMMA:
//     D[(ph, c)][(ky, j)] += sum_xg  G[y][8*xg + ph][c] * imgpad[y + ky][8*xg + j],      j = ph + kx in [0, 12)
//   A (MN-major, M = 8 phases x 16 channels = 128): the un-pooled gradient row exactly as it lies in memory,
//     assembled by the builder warps from the pooled gradient + argmax (bulk-copied ring);
//   B (K-major, 32-byte rows): R[r*12 + j][xg] = imgpad[r][8*xg + j], built ONCE per image for all 132 rows;
//     the operand of image row y is the 64-row window starting at row 12*y (rows 60..63 = garbage columns);
//     hi and lo bf16 copies (two MMAs into the same accumulator) keep the image exact to ~2^-17.
// 256 MMAs per image instead of 1024, every one with all 128 accumulator rows useful, and an eighth of the
// thread-side operand assembly of the tap-row formulation above.  gw[c][ky][kx] = sum_ph D[(ph,c)][(ky, ph+kx)].
static int g_fold = 1;
// Stage = kFoldRows image rows (two pooled rows): a builder team's latency chain per stage (mbarrier wait -> shared
// loads -> selects -> swizzled stores -> fence.proxy.async -> arrive, ~1000 cycles) is independent of how much each
// thread stores, so with one row per stage (round 1) the kernel ran at 5 % of its issue bound.
static constexpr int kFoldRows = 4, kFoldA = 4, kFoldP = 10, kFoldIssuers = 2, kFoldTeams = 2, kFoldTeamW = 8;
static constexpr int kFoldFirstB = 1 + kFoldIssuers;
static constexpr int kFoldThreads = 32 * (kFoldFirstB + kFoldTeams * kFoldTeamW);
static constexpr uint32_t kFoldBRows = 1600;          // 132 * 12 = 1584 rows + the tail the last window reads
static constexpr uint32_t kFoldRowBytes = 4096, kFoldABytes = kFoldRows * kFoldRowBytes, kFoldPSlot = 6144;

struct FoldParams {
  const float* img; const void* gp; const uint8_t* idx;
  int B;
  float* gw; float* gb;
  float* part;          // [gridDim.x][400] weight partials then [gridDim.x][16] bias partials (summed in CTA order); null: atomics
};

__global__ void __launch_bounds__(kFoldThreads, 1) conv1_wgrad_fold_kernel(const FoldParams p) {
  constexpr int H = 128, W = 128, Hp = 64, Wp = 64;
  constexpr uint32_t A_BYTES = kFoldABytes, P_SLOT = kFoldPSlot, B_ARR = kFoldBRows * 32u;
  constexpr int NBUILD = 32 * kFoldTeams * kFoldTeamW;
  constexpr uint32_t SPI = H / kFoldRows;               // stages per image
  static_assert(kFoldA % kFoldIssuers == 0 && kFoldA % kFoldTeams == 0, "ring owners");
  static_assert(kFoldP % kFoldTeams == 0 && kFoldRows == 4, "a stage is two pooled rows");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kFoldA], a_empty[kFoldA], p_full[kFoldP], p_empty[kFoldP], b_free, accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_gb[16];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sBhi = smem;
  uint8_t* sBlo = sBhi + B_ARR;
  uint8_t* sA = sBlo + B_ARR;
  uint8_t* sP = sA + kFoldA * A_BYTES;
  float* sD = reinterpret_cast<float*>(sA);            // epilogue scratch [128][65], after the pipeline drained (33 KB of the A ring)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kFoldA; ++s) { mbar_init(&a_full[s], kFoldTeamW); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kFoldP; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], kFoldTeamW); }
    mbar_init(&b_free, kFoldIssuers);
    mbar_init(&accum_bar, kFoldIssuers);
    fence_barrier_init();
  }
  if (tid < 16) s_gb[tid] = 0.f;
  if (warp == 1) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  int nimg = 0;
  for (int img = blockIdx.x; img < p.B; img += gridDim.x) ++nimg;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: pooled gradient rows + argmax
    uint32_t g = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
      if (lane == 0) {
        // one slot = one stage = two pooled rows (contiguous in both tensors): two bulk copies per 6 KB -- a copy
        // per pooled row kept this single lane, which shares its scheduler with busy builder warps, behind the builders
        const char* gsrc = reinterpret_cast<const char*>(p.gp) + (int64_t)img * Hp * Wp * 32;
        const uint8_t* isrc = p.idx + (int64_t)img * Hp * Wp * 16;
        for (uint32_t st = 0; st < SPI; ++st, ++g) {
          const uint32_t s = g % kFoldP;
          mbar_wait(&p_empty[s], ((g / kFoldP) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&p_full[s], 2u * Wp * 48u);
          bulk_load_1d(sP + s * P_SLOT, gsrc + st * (2u * Wp * 32u), 2u * Wp * 32u, &p_full[s]);
          bulk_load_1d(sP + s * P_SLOT + 4096, isrc + st * (2u * Wp * 16u), 2u * Wp * 16u, &p_full[s]);
        }
      } else if (img + (int)gridDim.x < p.B) {
        // the other lanes pull the NEXT image of this CTA into L2 while this one is processed: the B-array build is
        // the one phase in which the MMA pipeline is empty, and its first-touch misses were most of its cost
        const char* nx = reinterpret_cast<const char*>(p.img + (int64_t)(img + (int)gridDim.x) * H * W);
        for (int l = lane - 1; l < H * W * 4 / 128; l += 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + l * 128));
      }
      __syncwarp();
    }
  } else if (warp < kFoldFirstB) {
    // ------------------------------------------------------------------ MMA issuers (stage sg -> issuer sg % 2; whole warps)
    {
      const bool leader = elect_one();
      const uint32_t wi = (uint32_t)(warp - 1);
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 0);    // A MN-major, B K-major
      // A: two M atoms of 64 (LBO = 2048 B), K atoms of 8 rows 1024 B apart; B: 32-byte rows, 8-row groups 256 B apart
      const uint64_t ad0 = make_smem_desc(smem_u32(sA), 2048u, 1024u, 2u);
      const uint64_t bh0 = make_smem_desc(smem_u32(sBhi), 16u, 256u, 6u);
      const uint64_t bl0 = make_smem_desc(smem_u32(sBlo), 16u, 256u, 6u);
      const uint32_t d_addr = tmem_base + wi * 64u;
      const uint32_t total = (uint32_t)nimg * SPI;
      uint32_t a = wi % kFoldA, pa = 0u, first = 1u;
      for (uint32_t sg = wi; sg < total; sg += kFoldIssuers) {
        const uint32_t ys = sg % SPI;
        mbar_wait(&a_full[a], pa);
        tc_fence_after();
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)kFoldRows; ++r) {
          const uint64_t ad = ad0 + ((a * A_BYTES + r * kFoldRowBytes) >> 4);
          const uint32_t boff = ((ys * kFoldRows + r) * 12u * 32u) >> 4;
          if (leader) { umma_f16(d_addr, ad, bh0 + boff, idesc, first ? 0u : 1u); umma_f16(d_addr, ad, bl0 + boff, idesc, 1u); }
          first = 0u;
        }
        if (leader) umma_commit(&a_empty[a]);
        // last stage of an image handled by this issuer: the B arrays may be rebuilt once these MMAs are done
        if (ys + kFoldIssuers >= SPI && leader) umma_commit(&b_free);
        a += kFoldIssuers; if (a >= kFoldA) { a -= kFoldA; pa ^= 1u; }
      }
      if (leader) umma_commit(&accum_bar);
    }
  } else {
    // ------------------------------------------------------------------ builders
    const int bw = warp - kFoldFirstB, bt = tid - 32 * kFoldFirstB;
    const int team = bw / kFoldTeamW;
    const int task = (bw - team * kFoldTeamW) * 32 + lane;        // (pixel x, channel half h) of the image row
    const int x = task >> 1, h = task & 1;
    const int xg = x >> 3, ph = x & 7;
    // MN-major, 128-byte swizzled: M atom (ph / 4) * 2048 + K atom (xg / 8) * 1024 + K row (xg % 8) * 128 + 16-byte chunk
    const uint32_t a_off = (uint32_t)(ph >> 2) * 2048u + (uint32_t)(xg >> 3) * 1024u + (uint32_t)(xg & 7) * 128u +
                           ((((uint32_t)(ph & 3) * 2u + (uint32_t)h) ^ (uint32_t)(xg & 7)) << 4);
    const uint32_t xbit = (uint32_t)(x & 1), xm = xbit * 0x01010101u;
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};                          // channels h * 8 + xbit * 4 + {0..3}
    uint32_t ii = 0;
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++ii) {
      // ---- B arrays of this image: R[r*12 + j][xg] = imgpad[r][8*xg + j] (hi and lo)
      if (ii > 0) mbar_wait(&b_free, (ii - 1) & 1u);             // the previous image's MMAs are done reading them
      const float* im = p.img + (int64_t)img * H * W;
#pragma unroll 2
      for (int c = bt; c < 132 * 12 * 2; c += NBUILD) {           // chunk = (row n, 8 consecutive xg)
        const int n = c >> 1, kh = c & 1;
        const int r = n / 12, j = n - r * 12;
        const int iy = r - 2;
        const bool rowok = iy >= 0 && iy < H;
        // elements ix = 64*kh + 8*e + j - 2: only the first of the left chunk and the last of the right one can
        // fall outside the row
        const float* rp = im + (rowok ? iy : 0) * W + 64 * kh + j - 2;
        float v[8];
#pragma unroll
        for (int e = 1; e < 7; ++e) v[e] = __ldg(rp + 8 * e);
        v[0] = (kh || j >= 2) ? __ldg(rp) : 0.f;
        v[7] = (!kh || j < 10) ? __ldg(rp + 56) : 0.f;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {                               // two elements per conversion instruction
          const uint32_t h2 = pack_bf16x2(v[2 * e], v[2 * e + 1]);
          const float r0 = v[2 * e] - __uint_as_float(h2 << 16), r1 = v[2 * e + 1] - __uint_as_float(h2 & 0xffff0000u);
          hi[e] = rowok ? h2 : 0u;
          lo[e] = rowok ? pack_bf16x2(r0, r1) : 0u;
        }
        const uint32_t o = swz_off((uint32_t)n, (uint32_t)kh, 32u);
        *reinterpret_cast<uint4*>(sBhi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sBlo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(NBUILD) : "memory");
      // ---- one A stage per two pooled rows: un-pool them into the four image rows they came from
      for (uint32_t ys = (uint32_t)team; ys < SPI; ys += kFoldTeams) {
        const uint32_t sg = ii * SPI + ys;                         // global stage counter
        const uint32_t a = sg % kFoldA;
        const int local = (x >> 1) * 2 + h;                        // 16-byte units of the pooled slot
        uint4 pg[2]; uint2 pi[2];
        const uint32_t ps = sg % kFoldP;
        mbar_wait(&p_full[ps], (sg / kFoldP) & 1u);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          pg[pr] = *reinterpret_cast<const uint4*>(sP + ps * P_SLOT + pr * 2048 + local * 16);
          pi[pr] = *reinterpret_cast<const uint2*>(sP + ps * P_SLOT + 4096 + pr * 1024 + local * 8);
        }
        uint32_t ov[4][4];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t gv[4] = {pg[pr].x, pg[pr].y, pg[pr].z, pg[pr].w};
          // argmax bytes (0..3) -> per-byte flags in bit 7: position (yy, x & 1) matches iff byte ^ xbit is 0 (yy = 0)
          // or 2 (yy = 1); PRMT with sign replication then widens a flag byte to the 16-bit mask of its channel
          uint32_t f[2][2];
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const uint32_t t = (w ? pi[pr].y : pi[pr].x) ^ xm;
            const uint32_t a7 = t << 7, a6 = t << 6;
            f[w][0] = ~(a7 | a6) & 0x80808080u;
            f[w][1] = (a6 & ~a7) & 0x80808080u;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int yy = 0; yy < 2; ++yy) {
              uint32_t m;
              if (e & 1) asm("prmt.b32 %0, %1, %1, 0xbbaa;" : "=r"(m) : "r"(f[e >> 1][yy]));
              else asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m) : "r"(f[e >> 1][yy]));
              ov[pr * 2 + yy][e] = gv[e] & m;
            }
          // bias gradient = sum of the pooled gradient: the two threads of a pooled pixel take four channels each
          const uint32_t wa = xbit ? gv[2] : gv[0], wb = xbit ? gv[3] : gv[1];
          gsum[0] += __uint_as_float(wa << 16); gsum[1] += __uint_as_float(wa & 0xffff0000u);
          gsum[2] += __uint_as_float(wb << 16); gsum[3] += __uint_as_float(wb & 0xffff0000u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_empty[ps]);                  // pooled slot consumed (it is in registers)
        mbar_wait(&a_empty[a], ((sg / kFoldA) & 1u) ^ 1u);
#pragma unroll
        for (int r = 0; r < kFoldRows; ++r)
          *reinterpret_cast<uint4*>(sA + a * A_BYTES + r * kFoldRowBytes + a_off) = make_uint4(ov[r][0], ov[r][1], ov[r][2], ov[r][3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[a]);
      }
    }
    // ---- bias gradient: channel (h*8 + e) summed over this thread's pixels, then over the CTA
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = gsum[e];
      // lanes with equal (h, x & 1) = lane bits 0, 1: xor-shuffle over the other bits
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < 4) atomicAdd(&s_gb[h * 8 + (int)xbit * 4 + e], v);
    }
    // ---- epilogue: D[(ph,c)][(ky,j)] summed over the issuers' accumulators -> gw[c][ky][kx] += D[..][(ky, ph+kx)]
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, %0;" ::"n"(NBUILD) : "memory");       // every builder is past the rings: reuse them as scratch
    if (bw < 4) {
      const int q = warp & 3;
      const int m = q * 32 + lane;
      float acc[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc[c] = 0.f;
      for (uint32_t wi = 0; wi < (uint32_t)kFoldIssuers; ++wi) {
#pragma unroll
        for (int cc = 0; cc < 64; cc += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + wi * 64u + (uint32_t)cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[cc + c] += __uint_as_float(v[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) sD[m * 65 + c] = acc[c];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NBUILD) : "memory");
    // one value per weight and CTA (the eight pixel phases are summed here): round 2's 3200 atomics per CTA on 400
    // addresses cost a third of the kernel
    if (bt < 400) {
      const int c = bt / 25, t = bt - c * 25;
      const int ky = t / 5, kx = t - ky * 5;
      float a = 0.f;
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) a += sD[(ph * 16 + c) * 65 + ky * 12 + ph + kx];
      if (p.part) p.part[(int64_t)blockIdx.x * 400 + bt] = a;
      else atomicAdd(p.gw + bt, a);
    } else if (bt < 416) {
      const float a = s_gb[bt - 400];
      if (p.part) p.part[(int64_t)gridDim.x * 400 + (int64_t)blockIdx.x * 16 + (bt - 400)] = a;
      else if (p.gb) atomicAdd(p.gb + (bt - 400), a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

static int launch_conv1_wgrad_fold(const float* img, const void* gp, const uint8_t* idx, int B, float* gw, float* gb,
                                   cudaStream_t st) {
  FoldParams p;
  p.img = img; p.gp = gp; p.idx = idx; p.B = B; p.gw = gw; p.gb = gb;
  const size_t smem = 1024 + 2 * (size_t)kFoldBRows * 32 + (size_t)kFoldA * kFoldABytes + (size_t)kFoldP * kFoldPSlot;
  static OncePerDevice attr;
  if (attr.first()) { cudaFuncSetAttribute(conv1_wgrad_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);}
  int grid = kNumSMs;
  if (grid > B) grid = B;
  p.part = scratch_floats((int64_t)grid * 416);
  conv1_wgrad_fold_kernel<<<grid, kFoldThreads, smem, st>>>(p);
  LIVAE_CUDA_LAUNCH_CHECK();
  if (p.part) {                      // fixed-order sums over the CTAs: bit-reproducible (gw / gb are written)
    sum_slices(p.part, grid, 400, gw, st);
    if (gb) sum_slices(p.part + (int64_t)grid * 400, grid, 16, gb, st);
  }
  return 0;
}

// kind 0: STN conv1 (src = img [B,H,W], big = pooled gradient, idx); kind 1: encoder c1 (big = g [B,H/2,W/2,32]);
// kind 2: decoder d4 (src = gpre [B,H,W], big = x [B,H+2,W+2,32]).  gw / gb must be zeroed by the caller.
int thin_tc_wgrad(int kind, const float* src, const void* big, const uint8_t* idx, int B, int H, int W, float* gw,
                  float* gb, cudaStream_t st) {
  if (((uintptr_t)big & 15) != 0) return 1;
  if (kind == 0) {
    // a 64-pixel stage must cover whole pooled rows or half of one: W a multiple of 64, or 16 / 32
    if (!((W % 64) == 0 || W == 32 || W == 16) || (H & 3) || ((uintptr_t)idx & 15)) return 1;
    if (H == 128 && W == 128 && g_fold) return launch_conv1_wgrad_fold(src, big, idx, B, gw, gb, st);
    return launch_wg<0>(src, big, idx, B, H, W, H, W, gw, gb, st);
  }
  if (kind == 1) {
    if ((W & 15) || (H & 1)) return 1;
    return launch_wg<1>(src, big, nullptr, B, H, W, H / 2, W / 2, gw, gb, st);
  }
  return launch_wg<2>(src, big, nullptr, B, H, W, H + 2, W + 2, gw, gb, st);
}

}  // namespace tc
}  // namespace livae


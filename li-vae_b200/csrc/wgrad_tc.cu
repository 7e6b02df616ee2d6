// Engine 1, weight gradient: gw[tap][cb][cs] = sum_{b,oy,ox} x[b, oy*s-p+ky, ox*s-p+kx, cb] * gy[b,oy,ox,cs]
// on tcgen05 with BOTH operands MN-major: the reduction (UMMA K) axis is the pixel index, which is
// the slow axis of an NHWC tile, so the very same TMA boxes the forward kernel loads ([pixels x
// channels], 128/64/32-byte swizzled rows) are consumed directly -- no transposes anywhere.
//
//   D_g[128, Cs] += A_g[128 x 16 pixels] * B[Cs x 16 pixels]^T
//   rows of A_g: 128 consecutive (tap, cb) pairs ("row group" g), assembled from 128/kcb TMA boxes of
//   the tap-shifted input; B: the gy tile.  G row groups keep their accumulators in TMEM (G*Cs <= 512
//   columns) so one gy tile is reused for G*128 weight rows.
// A CTA owns (a set of G row groups) x (a contiguous range of 64-pixel tiles); partial sums are
// added to an fp32 [tap][cb][cs] buffer with red.global.add, then unpacked to torch's layout.
#include "tc_common.cuh"

namespace livae {
namespace tc {

static constexpr int kWgThreads = 192;
static constexpr int kPK = 64;   // pixels per pipeline stage (UMMA K = 16 -> 4 MMAs per group per stage)

struct WgradParams {
  int tw, th, nb;             // 64-pixel tile over the small-side (gy) grid
  int tiles_x, tiles_y, tiles_b;
  int stride, pad, kw, ntaps;
  int Cb, Cs, kcb, kcs;       // channel counts and channels per TMA box
  int G;                      // row groups per CTA
  int groups_total;
  int tiles_per_cta;
  float* gw_acc;              // fp32 [ntaps][Cb][Cs], zeroed by the host wrapper
};

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(kWgThreads) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                              const __grid_constant__ CUtensorMap tmG,
                                                              const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t xbox_bytes = (uint32_t)kPK * p.kcb * 2u;        // one input box
  const uint32_t gbox_bytes = (uint32_t)kPK * p.kcs * 2u;        // one gy box
  const int boxes_per_group = 128 / p.kcb;
  const int chunks_b = p.Cb / p.kcb;                             // channel chunks per tap
  const int nboxg = p.Cs / p.kcs;                                // gy boxes
  const uint32_t a_bytes = (uint32_t)p.G * boxes_per_group * xbox_bytes;   // = G * 16 KB
  const uint32_t b_bytes = (uint32_t)nboxg * gbox_bytes;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(p.G * p.Cs)) ncols <<= 1;

  const int gset = blockIdx.x;            // which set of G row groups
  const int g0 = gset * p.G;
  const int ng = min(p.G, p.groups_total - g0);
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int t_beg = blockIdx.y * p.tiles_per_cta;
  const int t_end = min(total_tiles, t_beg + p.tiles_per_cta);
  const int nt = t_end - t_beg;
  const int total_boxes = p.ntaps * chunks_b;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, ncols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (nt > 0) {
    if (warp == 0) {
      if (lane == 0) {
        // bytes that will actually arrive per stage: only boxes of real taps are loaded
        int live_boxes = 0;
        for (int g = 0; g < ng; ++g)
          for (int j = 0; j < boxes_per_group; ++j)
            if ((g0 + g) * boxes_per_group + j < total_boxes) ++live_boxes;
        const uint32_t tx_bytes = (uint32_t)live_boxes * xbox_bytes + b_bytes;
        for (int it = 0; it < nt; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          int tile = t_beg + it;
          const int tx = tile % p.tiles_x; tile /= p.tiles_x;
          const int ty = tile % p.tiles_y; tile /= p.tiles_y;
          const int b0 = tile * p.nb;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          uint8_t* st = smem + (uint32_t)s * stage_bytes;
          for (int c = 0; c < nboxg; ++c)
            tma_load_4d(st + a_bytes + (uint32_t)c * gbox_bytes, &tmG, &full_bar[s], c * p.kcs, tx * p.tw, ty * p.th, b0);
          for (int g = 0; g < ng; ++g)
            for (int j = 0; j < boxes_per_group; ++j) {
              const int f = (g0 + g) * boxes_per_group + j;
              if (f >= total_boxes) continue;
              const int tap = f / chunks_b, ch = f - tap * chunks_b;
              const int ky = tap / p.kw, kx = tap - ky * p.kw;
              tma_load_4d(st + (uint32_t)(g * boxes_per_group + j) * xbox_bytes, &tmX, &full_bar[s], ch * p.kcb,
                          tx * p.tw * p.stride - p.pad + kx, ty * p.th * p.stride - p.pad + ky, b0);
            }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, p.Cs, 1, 1);   // both operands MN-major
        const uint32_t rbx = (uint32_t)p.kcb * 2u, rbg = (uint32_t)p.kcs * 2u;
        const uint32_t ltx = rbx == 128 ? 2u : rbx == 64 ? 4u : 6u;
        const uint32_t ltg = rbg == 128 ? 2u : rbg == 64 ? 4u : 6u;
        const uint32_t sbox = 8u * rbx, sbog = 8u * rbg;           // 8 pixels (one K atom) apart
        for (int it = 0; it < nt; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (uint32_t)s * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
          for (int g = 0; g < ng; ++g) {
            const uint32_t ga = a_addr + (uint32_t)(g * boxes_per_group) * xbox_bytes;
#pragma unroll
            for (int k = 0; k < kPK / 16; ++k) {
              // K step = 16 pixels = two 8-pixel atoms; MN atoms (boxes) are LBO apart
              const uint64_t ad = make_smem_desc(ga + (uint32_t)k * 2u * sbox, xbox_bytes, sbox, ltx);
              const uint64_t bd = make_smem_desc(b_addr + (uint32_t)k * 2u * sbog, gbox_bytes, sbog, ltg);
              umma_f16(tmem_base + (uint32_t)(g * p.Cs), ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&accum_bar);
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
      for (int g = 0; g < ng; ++g) {
        const int f = (g0 + g) * boxes_per_group + row / p.kcb;      // (tap, chunk) of this row
        const bool valid = f < total_boxes;
        const int tap = f / chunks_b, ch = f - tap * chunks_b;
        const int cb = ch * p.kcb + row % p.kcb;
        float* dst = p.gw_acc + ((int64_t)tap * p.Cb + cb) * p.Cs;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.Cs);
        for (int c0 = 0; c0 < p.Cs; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) red_add_f32(dst + c0 + i, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// gw_acc fp32 [tap][cb][cs] -> torch layout gw[cs][cb][tap] (written)
__global__ void unpack_wgrad_kernel(const float* __restrict__ acc, int Cs, int Cb, int taps, float* __restrict__ gw) {
  int n = Cs * Cb * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int tap = i % taps; int t = i / taps; int cb = t % Cb; int cs = t / Cb;
    gw[i] = acc[((int64_t)tap * Cb + cb) * Cs + cs];
  }
}

// column sums of a bf16 [R, C] matrix -> fp32 gb[C] (written via atomics on a zeroed buffer)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ g, int64_t R, int C,
                                                          float* __restrict__ gb, int64_t rows_per_cta) {
  __shared__ float red[256];
  const int cols = C < 256 ? C : 256;
  const int rgs = 256 / cols;
  const int tid = threadIdx.x, cl = tid % cols, rg = tid / cols;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
  for (int c0 = 0; c0 < C; c0 += cols) {
    const int c = c0 + cl;
    float a = 0.f;
    if (rg < rgs && c < C)
      for (int64_t r = r0 + rg; r < r1; r += rgs) a += __bfloat162float(g[r * C + c]);
    red[tid] = a;
    __syncthreads();
    if (rg == 0 && c < C) {
      for (int j = 1; j < rgs; ++j) a += red[j * cols + cl];
      atomicAdd(gb + c, a);
    }
    __syncthreads();
  }
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;

extern "C" int64_t livae_tc_wgrad_ws_bytes(const livae_tc_conv_desc* d) {
  if (!d) return 0;
  return (int64_t)d->kh * d->kw * d->Cin * d->Cout * 4;
}

// gw (fp32, torch layout [Cout][Cin][kh][kw]) and gb (fp32 [Cout], optional) of the convolution d.
// x: bf16 [B,Hin,Win,Cin]; gy: bf16 [B,Ho,Wo,Cout] PRE-activation gradient; ws: livae_tc_wgrad_ws_bytes.
extern "C" int livae_tc_conv_wgrad(const livae_tc_conv_desc* d, const void* x, const void* gy, float* gw, float* gb,
                                   void* ws, livae_stream_t stream) {
  LIVAE_CHECK_ARG(d, "tc_conv_wgrad: null descriptor");
  if (d->B == 0) return 0;
  LIVAE_CHECK_ARG(x && gy && gw && ws, "tc_conv_wgrad: null pointer");
  LIVAE_CHECK_ARG(livae_tc_conv_supported(d) && (d->Cout == 16 || d->Cout == 32 || d->Cout % 64 == 0),
                  "tc_conv_wgrad: shape not supported by the tensor-core engine");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)gy | (uintptr_t)ws) & 15) == 0, "tc_conv_wgrad: pointers must be 16-byte aligned");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int s = d->stride;
  const int Ho = (d->Hin + 2 * d->pad - d->kh) / s + 1, Wo = (d->Win + 2 * d->pad - d->kw) / s + 1;
  WgradParams p;
  // 64-pixel tile over the gy grid
  if (Ho * Wo <= kPK) {
    p.tw = Wo; p.th = Ho; p.nb = kPK / (Ho * Wo);
    LIVAE_CHECK_ARG(p.nb * Ho * Wo == kPK, "tc_conv_wgrad: output map must divide 64 pixels");
  } else {
    p.nb = 1;
    p.tw = Wo < 16 ? Wo : 16;
    while (kPK % p.tw != 0) --p.tw;
    p.th = kPK / p.tw;
  }
  p.tiles_x = (Wo + p.tw - 1) / p.tw; p.tiles_y = (Ho + p.th - 1) / p.th; p.tiles_b = (d->B + p.nb - 1) / p.nb;
  p.stride = s; p.pad = d->pad; p.kw = d->kw; p.ntaps = d->kh * d->kw;
  p.Cb = d->Cin; p.Cs = d->Cout;
  p.kcb = p.Cb >= 64 ? 64 : p.Cb;
  p.kcs = p.Cs >= 64 ? 64 : p.Cs;
  LIVAE_CHECK_ARG(p.tw * s <= 256 && p.th * s <= 256, "tc_conv_wgrad: TMA box too large");
  const int rows_total = p.ntaps * p.Cb;
  p.groups_total = (rows_total + 127) / 128;
  int G = 512 / p.Cs;
  if (G > 4) G = 4;
  if (G > p.groups_total) G = p.groups_total;
  p.G = G;
  const int gsets = (p.groups_total + G - 1) / G;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  int splits = (kNumSMs + gsets - 1) / gsets;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.gw_acc = (float*)ws;

  CUtensorMap tmX, tmG;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Win, (uint64_t)d->Hin, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->Win * d->Cin * 2, (uint64_t)d->Hin * d->Win * d->Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kcb, (uint32_t)(p.tw * s), (uint32_t)(p.th * s), (uint32_t)p.nb};
    uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
    if (int e = make_tmap_bf16(&tmX, x, 4, dims, str, box, es, p.kcb * 2)) return e;
  }
  {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)Wo * d->Cout * 2, (uint64_t)Ho * Wo * d->Cout * 2};
    uint32_t box[4] = {(uint32_t)p.kcs, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.nb};
    if (int e = make_tmap_bf16(&tmG, gy, 4, dims, str, box, nullptr, p.kcs * 2)) return e;
  }
  cudaError_t ce = cudaMemsetAsync(ws, 0, (size_t)livae_tc_wgrad_ws_bytes(d), st);
  if (ce != cudaSuccess) { set_error("tc_conv_wgrad memset: %s", cudaGetErrorString(ce)); return (int)ce; }

  const uint32_t a_bytes = (uint32_t)G * 128u * kPK * 2u;
  const uint32_t b_bytes = (uint32_t)kPK * p.Cs * 2u;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  dim3 grid(gsets, splits);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (3u * stage_bytes + 1024u <= 200u * 1024u)
    wgrad_tc_kernel<3><<<grid, kWgThreads, 3 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  else
    wgrad_tc_kernel<2><<<grid, kWgThreads, 2 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  const int n = p.Cs * p.Cb * p.ntaps;
  unpack_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(p.gw_acc, p.Cs, p.Cb, p.ntaps, gw);
  LIVAE_CUDA_LAUNCH_CHECK();
  if (gb) {
    ce = cudaMemsetAsync(gb, 0, (size_t)p.Cs * sizeof(float), st);
    if (ce != cudaSuccess) { set_error("tc_conv_wgrad memset gb: %s", cudaGetErrorString(ce)); return (int)ce; }
    const int64_t R = (int64_t)d->B * Ho * Wo;
    int64_t rows = (R + kNumSMs * 4 - 1) / (kNumSMs * 4);
    if (rows < 64) rows = 64;
    colsum_bf16_kernel<<<(int)((R + rows - 1) / rows), 256, 0, st>>>((const __nv_bfloat16*)gy, R, p.Cs, gb, rows);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  return 0;
}

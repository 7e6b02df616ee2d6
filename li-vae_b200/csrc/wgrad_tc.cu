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
static constexpr int kWgHaloThreads = 192;
static constexpr int kWgIssuers = 5;         // halo kernel: lane 0 of warps 1-5 each issue the MMAs of their own row groups
static constexpr int kPK = 64;   // pixels per pipeline stage (UMMA K = 16 -> 4 MMAs per group per stage)

struct WgradParams {
  int tw, th, nb;             // 64-pixel tile over the small-side (gy) grid
  int tiles_x, tiles_y, tiles_b;
  int stride, pad, kw, ntaps;
  int Cb, Cs, kcb, kcs;       // channel counts and channels per TMA box
  int G;                      // row groups per CTA
  int groups_total;
  int tiles_per_cta;
  float* gw_acc;              // fp32 [ntaps][Cb][Cs], zeroed by the host wrapper (slice_elems == 0: red.global.add) or
  int64_t slice_elems;        // fp32 [splits][ntaps][Cb][Cs] per-split partial sums (plain stores), summed in split order
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
        float* dst = p.gw_acc + (int64_t)blockIdx.y * p.slice_elems + ((int64_t)tap * p.Cb + cb) * p.Cs;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.Cs);
        for (int c0 = 0; c0 < p.Cs; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (valid) {
            if (p.slice_elems) {       // this split's own slice: plain stores, summed later in split order
              float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) red_add_f32(dst + c0 + i, __uint_as_float(v[i]));
            }
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

// ---------------------------------------------------------------------------------------------
// Halo variant of the weight gradient.  Per pipeline stage ONE haloed input box per (stride-parity
// class, channel chunk) and one gy tile [TH rows x TW cols] are fetched; every filter tap is then a
// ROW-SHIFTED view of the same input box (MN-major descriptors take an arbitrary 16-byte-aligned
// start, and the swizzle is a function of the absolute shared address), so the input tile is read
// once instead of once per tap (9x / 25x less L2->shared traffic).
//   K step j = gy row j (16 pixels; TW = 8: two 8-pixel rows, K-atom stride = one input row)
//   A' start  = box + ((j + sy) * Wx + sx) * row_bytes          B' start = gy + j * 16 * row_bytes
// Row groups: 128 accumulator rows = 128/kcb "atoms" of kcb channels, LBO apart.  kcb = 64: the two
// atoms are two channel chunks of one tap, or two taps (LBO = difference of their shifts).
// kcb = 32/16: atoms are consecutive-column taps of one filter row (LBO = one pixel row); atoms
// past the end of the filter row compute garbage rows that the epilogue skips.
static constexpr int kMaxGroups = 32;
static constexpr int kMaxTapsW = 32;
struct WgGroup { uint32_t base_off, lbo; int16_t atom[8]; uint32_t boxmask; };   // atom: tap * 8 + chunk, or -1; boxmask: input boxes read
struct WgradHaloParams {
  int TH, TW, Wx;              // gy tile, input box row width (pixels)
  int tiles_x, tiles_y, B, tiles_per_cta;
  int stride;
  int nbox, nbox_cta; int16_t box_dy[16], box_dx[16], box_c[16];   // input boxes: origin offset and channel chunk
  uint32_t box_slot;           // bytes per input box slot (1024-aligned)
  uint32_t xbox_bytes;         // bytes one input box delivers (Hbox * Wx * kcb * 2)
  int Cb, Cs, kcb, kcs, ntaps;
  int ngroups, G;
  WgGroup grp[kMaxGroups];
  float* gw_acc;
  int64_t slice_elems;        // see WgradParams
  long long* probe;
  int x_s2d;                   // input operand read in space-to-depth form through a strided tensor map (tc_common.cuh)
  int g_cpr;                   // > 0: gy is a plain NHWC tensor [B,2Ho,2Wo,Cs/4] read block-wise (channels = (ey,ex,c));
                               //      g_cpr = channel chunks per pixel row of the 2x2 block
};

template <int STAGES>
__global__ void __launch_bounds__(kWgHaloThreads) wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                const __grid_constant__ CUtensorMap tmG,
                                                                const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rbx = (uint32_t)p.kcb * 2u, rbg = (uint32_t)p.kcs * 2u;
  const int nboxg = p.Cs / p.kcs;
  const uint32_t gbox_bytes = (uint32_t)(p.TH * p.TW) * rbg;
  const uint32_t gbox_slot = (gbox_bytes + 1023u) & ~1023u;
  const uint32_t a_region = (uint32_t)p.nbox_cta * p.box_slot;
  const uint32_t stage_bytes = a_region + (uint32_t)nboxg * gbox_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(p.G * p.Cs)) ncols <<= 1;

  const int g0 = blockIdx.x * p.G;
  const int ng = min(p.G, p.ngroups - g0);
  const int total_tiles = p.tiles_x * p.tiles_y * p.B;
  const int t_beg = blockIdx.y * p.tiles_per_cta;
  const int t_end = min(total_tiles, t_beg + p.tiles_per_cta);
  const int nt = t_end - t_beg;
  // only the input boxes (parity class x channel chunk) that this CTA's row groups read; slot = box id - first
  uint32_t need = 0u;
  for (int g = 0; g < ng; ++g) need |= p.grp[g0 + g].boxmask;
  const int box_lo = need ? __ffs(need) - 1 : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kWgIssuers); }
    mbar_init(&accum_bar, kWgIssuers);
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
        const uint32_t tx_bytes = (uint32_t)__popc(need) * p.xbox_bytes + (uint32_t)nboxg * gbox_bytes;
        int pn = 0;
        for (int it = 0; it < nt; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          int tile = t_beg + it;
          const int tx = tile % p.tiles_x; tile /= p.tiles_x;
          const int ty = tile % p.tiles_y; tile /= p.tiles_y;
          const int b = tile;
          probe_rec(p.probe, 0, 0, pn);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          probe_rec(p.probe, 0, 1, pn);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          uint8_t* st = smem + (uint32_t)s * stage_bytes;
          for (int c = 0; c < nboxg; ++c) {
            if (p.g_cpr)
              tma_load_4d(st + a_region + (uint32_t)c * gbox_slot, &tmG, &full_bar[s], (c % p.g_cpr) * p.kcs, tx * p.TW,
                          2 * ty * p.TH + c / p.g_cpr, b);
            else
              tma_load_4d(st + a_region + (uint32_t)c * gbox_slot, &tmG, &full_bar[s], c * p.kcs, tx * p.TW, ty * p.TH, b);
          }
          for (int i = 0; i < p.nbox; ++i) {
            if (!((need >> i) & 1u)) continue;
            if (p.x_s2d)   // chunk = pixel row inside the 2x2 block (HaloOpts in tc_common.cuh)
              tma_load_4d(st + (uint32_t)(i - box_lo) * p.box_slot, &tmX, &full_bar[s], 0, tx * p.TW + p.box_dx[i],
                          2 * (ty * p.TH + p.box_dy[i]) + p.box_c[i], b);
            else
              tma_load_4d(st + (uint32_t)(i - box_lo) * p.box_slot, &tmX, &full_bar[s], p.box_c[i] * p.kcb,
                          tx * p.TW * p.stride + p.box_dx[i], ty * p.TH * p.stride + p.box_dy[i], b);
          }
        }
      }
    } else {
      {
        const bool leader = elect_one();      // whole-warp loop, one lane issues (tc_common.cuh)
        // FIVE issuing warps (1-5; warps 2-5 are idle until the epilogue anyway), row groups
        // dealt round-robin: one thread sustains one tcgen05.mma per ~50-85 cycles (tools/umma_rate.cu and the
        // probe of this loop), but an M=128 MMA with N <= 64 only occupies the pipe for (128+N)/4 = 36..48
        // cycles (shared-memory operand read), so a single issuer left the tensor pipe idle half the time on
        // the narrow layers.  All wait on the same full barrier and all commit to the stage's empty barrier.
        const int iw = warp - 1;
        // per-thread issue loop: every descriptor is "precomputed constant + stage base + K-step
        // increment" (two 64-bit adds per MMA; building descriptors from kernel parameters inside the
        // loop cost several hundred cycles of dependent latency per MMA)
        const uint32_t idesc = make_idesc_bf16(128, p.Cs, 1, 1);
        const uint32_t ltx = rbx == 128 ? 2u : rbx == 64 ? 4u : 6u;
        const uint32_t ltg = rbg == 128 ? 2u : rbg == 64 ? 4u : 6u;
        // K atom = 8 pixels: TW = 16 -> consecutive rows; TW = 8 -> the next gy row, i.e. one input row down
        const uint32_t sbox = (p.TW == 16 ? 8u : (uint32_t)p.Wx) * rbx;
        const uint32_t sbog = 8u * rbg;
        const int ksteps = p.TH * p.TW / 16;
        const int rows_per_kstep = 16 / p.TW;   // gy rows covered by one K step
        const uint32_t a_kstep16 = ((uint32_t)(rows_per_kstep * p.Wx) * rbx) >> 4;
        const uint32_t b_kstep16 = (16u * rbg) >> 4;
        const uint64_t bd0 = make_smem_desc(smem_u32(smem) + a_region, gbox_slot, sbog, ltg);
        const uint32_t stage16 = stage_bytes >> 4;
        // this warp's row groups (G <= 8, five issuers: at most two), descriptors in registers
        const int gA = iw, gB = iw + kWgIssuers;
        const bool hasA = gA < ng, hasB = gB < ng;
        const uint32_t tile_base = smem_u32(smem) - (uint32_t)box_lo * p.box_slot;
        const uint64_t dA = hasA ? make_smem_desc(tile_base + p.grp[g0 + gA].base_off, p.grp[g0 + gA].lbo, sbox, ltx) : 0ull;
        const uint64_t dB = hasB ? make_smem_desc(tile_base + p.grp[g0 + gB].base_off, p.grp[g0 + gB].lbo, sbox, ltx) : 0ull;
        const uint32_t tA = tmem_base + (uint32_t)(gA * p.Cs), tB = tmem_base + (uint32_t)(gB * p.Cs);
        uint32_t accum = 0u;
        int pn = 0;
        for (int it = 0; it < nt; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          if (iw == 0 && leader) probe_rec(p.probe, 1, 0, pn);
          mbar_wait(&full_bar[s], ph);
          if (iw == 0 && leader) probe_rec(p.probe, 1, 1, pn);
          tc_fence_after();
          const uint32_t so = (uint32_t)s * stage16;
          if (hasA) {
            uint64_t ad = dA + so, bd = bd0 + so;
            for (int k = 0; k < ksteps; ++k, ad += a_kstep16, bd += b_kstep16)
              if (leader) umma_f16(tA, ad, bd, idesc, accum | (uint32_t)(k > 0));
          }
          if (hasB) {
            uint64_t ad = dB + so, bd = bd0 + so;
            for (int k = 0; k < ksteps; ++k, ad += a_kstep16, bd += b_kstep16)
              if (leader) umma_f16(tB, ad, bd, idesc, accum | (uint32_t)(k > 0));
          }
          accum = 1u;
          if (iw == 0 && leader) probe_rec(p.probe, 1, 2, pn);
          if (leader) umma_commit(&empty_bar[s]);
        }
        if (leader) umma_commit(&accum_bar);
      }
      __syncwarp();
      if (warp >= 2) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
      for (int g = 0; g < ng; ++g) {
        const int a = p.grp[g0 + g].atom[row / p.kcb];
        const bool valid = a >= 0;
        const int tap = a >> 3, ch = a & 7;
        const int cb = ch * p.kcb + row % p.kcb;
        float* dst = p.gw_acc + (int64_t)blockIdx.y * p.slice_elems + ((int64_t)tap * p.Cb + cb) * p.Cs;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.Cs);
        for (int c0 = 0; c0 < p.Cs; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (valid) {
            if (p.slice_elems) {       // this split's own slice: plain stores, summed later in split order
              float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) red_add_f32(dst + c0 + i, __uint_as_float(v[i]));
            }
          }
        }
      }
      tc_fence_before();
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// out[i] = sum over slices z of part[z][i]: the fixed-order second pass of a reduction that was split over CTAs
// (bit-reproducible, unlike red.global.add whose commit order follows CTA retirement).  `lanes` (a power of two
// <= 32) threads share one element: lane l adds slices l, l + lanes, ... with four independent accumulators (the
// loads of a serial `a += part[z]` chain would each wait a full L2 round trip), then the lanes are combined by a
// shuffle tree -- the association is a function of (nslices, lanes) only.
__global__ void sum_slices_kernel(const float* __restrict__ part, int nslices, int64_t n, float* __restrict__ out, int lanes) {
  const int64_t total = n * lanes;           // every thread of a block runs the same number of iterations (full-mask shuffles)
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gtid = base + threadIdx.x;
    const int64_t i = gtid / lanes;
    const int l = (int)(gtid - i * lanes);
    const bool live = i < n;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (live) {
      int z = l;
      for (; z + 3 * lanes < nslices; z += 4 * lanes) {
        a0 += part[(int64_t)z * n + i];
        a1 += part[(int64_t)(z + lanes) * n + i];
        a2 += part[(int64_t)(z + 2 * lanes) * n + i];
        a3 += part[(int64_t)(z + 3 * lanes) * n + i];
      }
      for (; z < nslices; z += lanes) a0 += part[(int64_t)z * n + i];
    }
    float a = (a0 + a1) + (a2 + a3);
    for (int o = lanes >> 1; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o, lanes);
    if (live && l == 0) out[i] = a;
  }
}
void sum_slices(const float* part, int nslices, int64_t n, float* out, cudaStream_t st) {
  int lanes = 1;
  while (lanes < 32 && n * lanes < 16384 && lanes * 8 <= nslices) lanes <<= 1;    // few elements, many slices: share them
  int64_t blocks = (n * lanes + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  sum_slices_kernel<<<(int)blocks, 256, 0, st>>>(part, nslices, n, out, lanes);
  count_launch(1);
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
// part_mode: gb is [gridDim.x][C] and every CTA stores its own row of partial sums (summed in CTA order afterwards)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ g, int64_t R, int C,
                                                          float* __restrict__ gb, int64_t rows_per_cta, int part_mode) {
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
      if (part_mode) gb[(int64_t)blockIdx.x * C + c] = a;
      else atomicAdd(gb + c, a);
    }
    __syncthreads();
  }
}

// The same for C a multiple of 8 and at most 256 (every layer of the step): a thread owns 8 channels and reads them as
// one 16-byte load, C/8 threads cover a row, the other thread groups of the CTA walk further rows -- 2-byte loads per
// thread ran the kernel at 1.7 TB/s.  Same fixed association per CTA (row groups summed in group order).
__global__ void __launch_bounds__(256) colsum_bf16_v8_kernel(const uint4* __restrict__ g, int64_t R, int C8,
                                                             float* __restrict__ gb, int64_t rows_per_cta, int part_mode) {
  __shared__ float red[256 * 8];
  const int rgs = 256 / C8;
  const int tid = threadIdx.x, cl = tid % C8, rg = tid / C8;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rg < rgs) {
    int64_t r = r0 + rg;
    for (; r + 3 * rgs < r1; r += 4 * rgs) {        // four loads in flight
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(g + (r + (int64_t)u * rgs) * C8 + cl);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[2 * k] += __uint_as_float(w[k] << 16); a[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u); }
      }
    }
    for (; r < r1; r += rgs) {
      const uint4 v = __ldg(g + r * C8 + cl);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[2 * k] += __uint_as_float(w[k] << 16); a[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u); }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[k * 256 + tid] = a[k];
  __syncthreads();
  if (tid < 8 * C8) {                                // one thread per channel: channel c = 8 * (c / 8) + c % 8
    const int c = tid, lane8 = c >> 3, k = c & 7;
    float s = 0.f;
    for (int j = 0; j < rgs; ++j) s += red[k * 256 + j * C8 + lane8];
    if (part_mode) gb[(int64_t)blockIdx.x * (8 * C8) + c] = s;
    else atomicAdd(gb + c, s);
  }
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;


static int g_wgrad_halo = 1;

// returns 0 = launched, 1 = shape not eligible (caller falls back to the per-tap kernel)
namespace livae { namespace tc {
int launch_wgrad_halo(const livae_tc_conv_desc* d, const void* x, const void* gy, float* gw_acc, int Ho,
                      int Wo, cudaStream_t st, int x_s2d) {
  const int s = d->stride, kh = d->kh, kw = d->kw, pad = d->pad;
  if (Wo < 8 || Ho < 2 || kh * kw < 2) return 1;
  WgradHaloParams p;
  p.Cb = d->Cin; p.Cs = d->Cout; p.ntaps = kh * kw; p.stride = s; p.B = d->B;
  p.kcb = p.Cb >= 64 ? 64 : p.Cb;
  if (x_s2d & 1) p.kcb = p.Cb / 2;
  const bool g_s2d = (x_s2d & 2) != 0;       // bit 1: gy read block-wise (WgradHaloParams::g_cpr)
  p.kcs = p.Cs >= 64 ? 64 : p.Cs;
  const int nchunk = p.Cb / p.kcb;
  if (nchunk > 8) return 1;
  const int apg = 128 / p.kcb;
  const uint32_t rbx = (uint32_t)p.kcb * 2u;
  p.TW = Wo >= 16 ? 16 : 8;
  // parity classes of the tap offsets (dy = ky - pad, in input pixels)
  struct Tap { int par, sy, sx, idx; };
  Tap taps[kMaxTapsW];
  int npar = 0, par_key[4][2], par_min[4][2];
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) {
      int dy = ky - pad, dx = kx - pad;
      int ry = ((dy % s) + s) % s, rx = ((dx % s) + s) % s;
      int g = -1;
      for (int i = 0; i < npar; ++i) if (par_key[i][0] == ry && par_key[i][1] == rx) g = i;
      if (g < 0) { if (npar == 4) return 1; g = npar++; par_key[g][0] = ry; par_key[g][1] = rx; par_min[g][0] = dy; par_min[g][1] = dx; }
      if (dy < par_min[g][0]) par_min[g][0] = dy;
      if (dx < par_min[g][1]) par_min[g][1] = dx;
      taps[ky * kw + kx] = Tap{g, dy, dx, ky * kw + kx};
    }
  int max_sy = 0, max_sx = 0;
  for (int t = 0; t < p.ntaps; ++t) {
    taps[t].sy = (taps[t].sy - par_min[taps[t].par][0]) / s;
    taps[t].sx = (taps[t].sx - par_min[taps[t].par][1]) / s;
    if (taps[t].sy > max_sy) max_sy = taps[t].sy;
    if (taps[t].sx > max_sx) max_sx = taps[t].sx;
  }
  p.Wx = p.TW + max_sx;
  // plan(TH): row groups, their input boxes and the stage size for a gy tile of TH rows; 0 = fits (>= 2 stages)
  int gsets = 1, splits = 1, total_tiles = 0;
  uint32_t stage_bytes = 0;
  auto plan = [&](int TH) -> int {
  if (TH * p.TW < 16) return 1;
  if (TH > Ho) { TH = Ho; if ((TH * p.TW) % 16 != 0) return 1; }
  p.TH = TH;
  const int Hbox = TH + max_sy;
  if (Hbox * s > 256 || p.Wx * s > 256) return 1;
  p.xbox_bytes = (uint32_t)(Hbox * p.Wx) * rbx;
  p.box_slot = (p.xbox_bytes + 2048u + 1023u) & ~1023u;   // slack: garbage atoms read past the box
  p.nbox = npar * nchunk;
  if (p.nbox > 16) return 1;
  for (int g = 0; g < npar; ++g)
    for (int c = 0; c < nchunk; ++c) {
      p.box_dy[g * nchunk + c] = (int16_t)par_min[g][0];
      p.box_dx[g * nchunk + c] = (int16_t)par_min[g][1];
      p.box_c[g * nchunk + c] = (int16_t)c;
    }
  auto tap_off = [&](const Tap& t, int c) -> uint32_t {
    return (uint32_t)(t.par * nchunk + c) * p.box_slot + (uint32_t)(t.sy * p.Wx + t.sx) * rbx;
  };
  // row groups
  int ng = 0;
  auto new_group = [&]() -> WgGroup* {
    if (ng == kMaxGroups) return nullptr;
    WgGroup* g = &p.grp[ng++];
    for (int i = 0; i < 8; ++i) g->atom[i] = -1;
    g->lbo = rbx;
    g->boxmask = 0u;
    return g;
  };
  auto box_of = [&](const Tap& t, int c) -> uint32_t { return 1u << (t.par * nchunk + c); };
  if (apg == 2 && nchunk >= 2) {
    // parity class outermost: the groups of one CTA then share their input boxes
    for (int par = 0; par < npar; ++par)
      for (int c = 0; c < nchunk; c += 2)
        for (int t = 0; t < p.ntaps; ++t) {
          if (taps[t].par != par) continue;
          WgGroup* g = new_group(); if (!g) return 1;
          g->base_off = tap_off(taps[t], c); g->lbo = p.box_slot;
          g->atom[0] = (int16_t)(taps[t].idx * 8 + c); g->atom[1] = (int16_t)(taps[t].idx * 8 + c + 1);
          g->boxmask = box_of(taps[t], c) | box_of(taps[t], c + 1);
        }
  } else if (apg == 2) {   // one chunk: pair taps of the same parity class (ascending shift)
    for (int par = 0; par < npar; ++par) {
      int prev = -1;
      for (int t = 0; t < p.ntaps; ++t) {
        if (taps[t].par != par) continue;
        if (prev < 0) { prev = t; continue; }
        WgGroup* g = new_group(); if (!g) return 1;
        g->base_off = tap_off(taps[prev], 0); g->lbo = tap_off(taps[t], 0) - tap_off(taps[prev], 0);
        g->atom[0] = (int16_t)(taps[prev].idx * 8); g->atom[1] = (int16_t)(taps[t].idx * 8);
        g->boxmask = box_of(taps[prev], 0);
        prev = -1;
      }
      if (prev >= 0) {
        WgGroup* g = new_group(); if (!g) return 1;
        g->base_off = tap_off(taps[prev], 0); g->lbo = rbx;
        g->atom[0] = (int16_t)(taps[prev].idx * 8);
        g->boxmask = box_of(taps[prev], 0);
      }
    }
  } else {   // kcb = 32 / 16: atoms = consecutive-column taps of one (parity, row shift, channel chunk)
    for (int c = 0; c < nchunk; ++c)
      for (int par = 0; par < npar; ++par)
        for (int sy = 0; sy <= max_sy; ++sy) {
          int cnt = 0; WgGroup* g = nullptr;
          for (int sx = 0; sx <= max_sx; ++sx) {
            int found = -1;
            for (int t = 0; t < p.ntaps; ++t) if (taps[t].par == par && taps[t].sy == sy && taps[t].sx == sx) found = t;
            if (found < 0) continue;
            if (!g || cnt == apg) {
              g = new_group(); if (!g) return 1;
              g->base_off = tap_off(taps[found], c); g->lbo = rbx; cnt = 0;
            }
            // atoms are LBO = one pixel row apart: column shifts must be consecutive inside a group
            int a = (int)((tap_off(taps[found], c) - g->base_off) / rbx);
            if (a >= apg) { g = new_group(); if (!g) return 1; g->base_off = tap_off(taps[found], c); g->lbo = rbx; a = 0; }
            g->atom[a] = (int16_t)(taps[found].idx * 8 + c);
            g->boxmask |= box_of(taps[found], c);
            cnt = a + 1;
          }
        }
  }
  p.ngroups = ng;
  int G = 512 / p.Cs;
  if (G > 8) G = 8;
  if (G > ng) G = ng;
  gsets = (ng + G - 1) / G;
  G = (ng + gsets - 1) / gsets;          // balance the row groups over the sets
  gsets = (ng + G - 1) / G;
  p.G = G;
  p.tiles_x = (Wo + p.TW - 1) / p.TW; p.tiles_y = (Ho + TH - 1) / TH;
  total_tiles = p.tiles_x * p.tiles_y * d->B;
  splits = (kNumSMs + gsets - 1) / gsets;
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (total_tiles + splits - 1) / splits;
  splits = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.gw_acc = gw_acc;
  p.slice_elems = 0;
  p.probe = g_probe;
  p.x_s2d = x_s2d & 1;
  p.g_cpr = 0;
  if (p.x_s2d && (d->Cin != 64 || s != 1)) return 1;
  if (g_s2d) {
    if (s != 1 || p.Cs % 128 != 0 || p.Cs > 256) return 1;      // run of a pixel row = Cs/2 >= 64 channels; UMMA N <= 256
    p.g_cpr = (p.Cs / 2) / p.kcs;
  }
  const int nboxg = p.Cs / p.kcs;
  const uint32_t gbox_slot = ((uint32_t)(TH * p.TW * p.kcs * 2) + 1023u) & ~1023u;
  // a CTA keeps only the boxes its row groups read (a contiguous range of box ids: groups are ordered
  // parity class / channel chunk outermost), in slots numbered from the first one
  p.nbox_cta = 1;
  for (int gs = 0; gs < gsets; ++gs) {
    uint32_t need = 0u;
    for (int g = gs * G; g < ng && g < (gs + 1) * G; ++g) need |= p.grp[g].boxmask;
    int lo = 0, hi = 0;
    while (!((need >> lo) & 1u)) ++lo;
    for (int i = 0; i < 16; ++i) if ((need >> i) & 1u) hi = i;
    if (hi - lo + 1 > p.nbox_cta) p.nbox_cta = hi - lo + 1;
  }
  stage_bytes = (uint32_t)p.nbox_cta * p.box_slot + (uint32_t)nboxg * gbox_slot;
  return (2u * stage_bytes + 1024u > 200u * 1024u) ? 2 : 0;
  };
  {
    int rc = 1;
    for (int TH = 8; TH >= 2; TH >>= 1) {
      rc = plan(TH);
      if (rc == 0) break;          // the largest tile that leaves room for two stages
      if (rc == 1) return 1;
    }
    if (rc != 0) { rc = plan(2); if (rc != 0) return 1; }
  }
  const int TH = p.TH, Hbox = TH + max_sy;

  CUtensorMap tmX, tmG;
  if (p.x_s2d) {
    const uint64_t C = (uint64_t)d->Cin / 4, Wf = 2 * (uint64_t)d->Win, Hf = 2 * (uint64_t)d->Hin;
    uint64_t dims[4] = {2 * C, (uint64_t)d->Win, Hf, (uint64_t)d->B};
    uint64_t str[3] = {2 * C * 2, Wf * C * 2, Hf * Wf * C * 2};
    uint32_t box[4] = {(uint32_t)(2 * C), (uint32_t)p.Wx, (uint32_t)(2 * Hbox), 1u};
    uint32_t es[4] = {1, 1, 2, 1};
    if (int e = make_tmap_bf16(&tmX, x, 4, dims, str, box, es, p.kcb * 2)) return e;
  } else {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Win, (uint64_t)d->Hin, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->Win * d->Cin * 2, (uint64_t)d->Hin * d->Win * d->Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kcb, (uint32_t)(p.Wx * s), (uint32_t)(Hbox * s), 1u};
    uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
    if (int e = make_tmap_bf16(&tmX, x, 4, dims, str, box, es, p.kcb * 2)) return e;
  }
  if (g_s2d) {       // see launch_conv_tc_halo: {(ex, c): 2C contiguous, X: stride 2C, pixel row 2*Y + ey (element stride 2), B}
    const uint64_t C = (uint64_t)d->Cout / 4, Wf = 2 * (uint64_t)Wo, Hf = 2 * (uint64_t)Ho;
    uint64_t dims[4] = {2 * C, (uint64_t)Wo, Hf, (uint64_t)d->B};
    uint64_t str[3] = {2 * C * 2, Wf * C * 2, Hf * Wf * C * 2};
    uint32_t box[4] = {(uint32_t)p.kcs, (uint32_t)p.TW, (uint32_t)(2 * TH), 1u};
    uint32_t es[4] = {1, 1, 2, 1};
    if (int e = make_tmap_bf16(&tmG, gy, 4, dims, str, box, es, p.kcs * 2)) return e;
  } else {
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)Wo * d->Cout * 2, (uint64_t)Ho * Wo * d->Cout * 2};
    uint32_t box[4] = {(uint32_t)p.kcs, (uint32_t)p.TW, (uint32_t)TH, 1u};
    if (int e = make_tmap_bf16(&tmG, gy, 4, dims, str, box, nullptr, p.kcs * 2)) return e;
  }
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(wgrad_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(wgrad_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 grid(gsets, splits);
  const int64_t elems = (int64_t)p.ntaps * p.Cb * p.Cs;
  float* part = scratch_floats((int64_t)splits * elems);
  if (part) {                                               // per-split slices, summed in split order below
    cudaError_t ce = cudaMemsetAsync(part, 0, (size_t)splits * elems * sizeof(float), st);   // rows no group owns stay 0
    if (ce != cudaSuccess) { set_error("wgrad scratch memset: %s", cudaGetErrorString(ce)); return (int)ce; }
    p.gw_acc = part; p.slice_elems = elems;
  }
  if (4u * stage_bytes + 1024u <= 200u * 1024u)
    wgrad_halo_kernel<4><<<grid, kWgHaloThreads, 4 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  else
    wgrad_halo_kernel<2><<<grid, kWgHaloThreads, 2 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  if (part) sum_slices(part, splits, elems, gw_acc, st);
  return 0;
}

void colsum_bf16(const void* g, int64_t R, int C, float* gb, cudaStream_t st) {
  int64_t rows = (R + kNumSMs * 4 - 1) / (kNumSMs * 4);
  if (rows < 64) rows = 64;
  const int blocks = (int)((R + rows - 1) / rows);
  const bool v8 = (C & 7) == 0 && C <= 256 && ((uintptr_t)g & 15) == 0;
  if (float* part = scratch_floats((int64_t)blocks * C)) {       // two fixed-order passes (gb need not be zeroed)
    if (v8) colsum_bf16_v8_kernel<<<blocks, 256, 0, st>>>((const uint4*)g, R, C / 8, part, rows, 1);
    else colsum_bf16_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g, R, C, part, rows, 1);
    count_launch(1);
    sum_slices(part, blocks, C, gb, st);
    return;
  }
  if (v8) colsum_bf16_v8_kernel<<<blocks, 256, 0, st>>>((const uint4*)g, R, C / 8, gb, rows, 0);
  else colsum_bf16_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g, R, C, gb, rows, 0);
  count_launch(1);
}
}}  // namespace livae::tc

// gb[C] (fp32, written) = column sums of the bf16 matrix g[R, C]: a bias gradient from a pre-activation gradient
extern "C" int livae_colsum_bf16(const void* g, int64_t R, int C, float* gb, livae_stream_t stream) {
  LIVAE_CHECK_ARG(R >= 0 && C > 0, "colsum_bf16: bad sizes");
  LIVAE_CHECK_ARG(gb && (g || R == 0), "colsum_bf16: null pointer");
  if (int e = require_sm100()) return e;
  cudaError_t ce = cudaMemsetAsync(gb, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
  if (ce != cudaSuccess) { set_error("colsum_bf16 memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  if (R == 0) return 0;
  livae::tc::colsum_bf16(g, R, C, gb, (cudaStream_t)stream);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

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
  p.slice_elems = 0;

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

  int halo_rc = g_wgrad_halo ? launch_wgrad_halo(d, x, gy, (float*)ws, Ho, Wo, st, 0) : 1;
  if (halo_rc != 0 && halo_rc != 1) return halo_rc;
  if (halo_rc == 1) {

  const uint32_t a_bytes = (uint32_t)G * 128u * kPK * 2u;
  const uint32_t b_bytes = (uint32_t)kPK * p.Cs * 2u;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  dim3 grid(gsets, splits);
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int64_t elems = (int64_t)p.ntaps * p.Cb * p.Cs;
  float* part = scratch_floats((int64_t)splits * elems);
  p.slice_elems = 0;
  if (part) {
    ce = cudaMemsetAsync(part, 0, (size_t)splits * elems * sizeof(float), st);
    if (ce != cudaSuccess) { set_error("wgrad scratch memset: %s", cudaGetErrorString(ce)); return (int)ce; }
    p.gw_acc = part; p.slice_elems = elems;
  }
  if (3u * stage_bytes + 1024u <= 200u * 1024u)
    wgrad_tc_kernel<3><<<grid, kWgThreads, 3 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  else
    wgrad_tc_kernel<2><<<grid, kWgThreads, 2 * stage_bytes + 1024, st>>>(tmX, tmG, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  if (part) { sum_slices(part, splits, elems, (float*)ws, st); p.gw_acc = (float*)ws; }
  }
  const int n = p.Cs * p.Cb * p.ntaps;
  unpack_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(p.gw_acc, p.Cs, p.Cb, p.ntaps, gw);
  LIVAE_CUDA_LAUNCH_CHECK();
  if (gb) {
    ce = cudaMemsetAsync(gb, 0, (size_t)p.Cs * sizeof(float), st);
    if (ce != cudaSuccess) { set_error("tc_conv_wgrad memset gb: %s", cudaGetErrorString(ce)); return (int)ce; }
    colsum_bf16(gy, (int64_t)d->B * Ho * Wo, p.Cs, gb, st);
    LIVAE_CUDA_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" void livae_tc_set_wgrad_halo(int mode) { g_wgrad_halo = mode ? 1 : 0; }

// Engine 1: tcgen05 / TMA implicit-GEMM convolution for the dense layers of the step (encoder
// c2-c4, decoder d1-d3, STN conv2; reference model.py:207, 292-296, 359-367), bf16 operands,
// fp32 accumulation in TMEM.  Parity class: 1e-2 relative (bf16 GEMM inputs, north_star).
//
// Formulation ("tap decomposition"): for a 128-pixel output tile (nb images x th x tw pixels),
//     D[128, Cout] = sum over taps (ky,kx) and Cin-chunks c of  A_tap,c[128, kc] * W_tap,c[Cout, kc]^T
// where A_tap,c is the input tile shifted by the tap -- ONE TMA box load from the NHWC tensor
// (zero padding = TMA out-of-bounds fill; stride 2 = tensor-map elementStrides), landing in shared
// memory already in the swizzled K-major layout tcgen05.mma consumes.  No im2col, no thread gathers.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2-5 = epilogue (TMEM -> registers -> bias/activation/ReLU-mask -> bf16/fp32 NHWC).
// A ring of STAGES {A,B} buffers with full/empty mbarriers decouples TMA from the MMA issue.
#include "tc_common.cuh"

namespace livae {
namespace tc {

long long* g_probe = nullptr;

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, int inner_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
  CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : inner_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                              : CU_TENSOR_MAP_SWIZZLE_NONE;
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return -3; }
  return 0;
}

static constexpr int kMaxTaps = 32;
static constexpr int kActSplitK = 99;     // internal epilogue mode of conv_tc_kernel (never part of the ABI)
struct ConvTcParams {
  int tiles_x, tiles_y;   // tiles per image in x / y (ragged last tiles are predicated)
  int tw, th, nb;         // tile = nb images x th rows x tw cols <= 128 pixels
  int Hq, Wq;             // tile-grid extent in output-grid units (Ho/os, Wo/os)
  int Ho, Wo;             // full output spatial size
  int os, oy0, ox0;       // output pixel = (q * os + o?0): os = 2 for the stride-2 phase kernels
  int in_stride;          // input coordinate of grid point q is q * in_stride + tap offset
  int ntaps;
  int8_t tap_dy[kMaxTaps], tap_dx[kMaxTaps];  // input offset of each tap (pad already folded in)
  int8_t tap_w[kMaxTaps];                     // index of each tap in the packed weight tensor
  int kc, nkc;            // channels per k-block, k-blocks per tap
  int N;                  // output channels per CTA (UMMA N); blockIdx.y selects the chunk
  int Ntot;               // total output channels (row pitch of out / relu_mask / bias)
  int B;
  void* out;              // [B,Ho,Wo,N] bf16 or fp32
  int out_f32;
  const float* bias;      // may be null
  int act;
  const __nv_bfloat16* relu_mask;  // may be null: out *= (relu_mask > 0), same shape as out
  int ksplit;             // > 1: blockIdx.z owns a slice of the K blocks and hands over its raw partial sums
  float* part;            // ksplit > 1: fp32 [ksplit][pixels][Ntot] partial sums (plain stores, summed in slice order by
  int64_t part_stride;    //   the finishing kernel); null: red.global.add into the zeroed fp32 out
};

// Epilogue of one accumulator row per thread: TMEM -> registers -> bias / activation / ReLU mask of the
// consumer -> bf16 or fp32 NHWC row.  tcgen05.ld is warp-collective, so invalid rows still issue it.
__device__ __forceinline__ void epilogue_rows(uint32_t taddr, int nbase, int N, int Ntot, bool valid, int64_t pix,
                                              void* out, int out_f32, const float* __restrict__ bias, int act,
                                              const __nv_bfloat16* __restrict__ relu_mask, float* part = nullptr) {
    for (int cc = 0; cc < N; cc += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + (uint32_t)cc, v);   // warp-collective: issued by all lanes, stores predicated
      tmem_ld_wait();
      if (!valid) continue;
      const int c0 = nbase + cc;
      if (act == kActSplitK) {     // split-K partial: raw sums, bias / activation applied by the finishing kernel
        if (part) {                // this slice's own copy: summed later in slice order (deterministic)
          float4* o = reinterpret_cast<float4*>(part + pix * Ntot + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
          continue;
        }
        float* o = reinterpret_cast<float*>(out) + pix * Ntot + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o + i), "f"(__uint_as_float(v[i])) : "memory");
        continue;
      }
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(v[i]);
        if (bias) a += __ldg(bias + c0 + i);
        if (act == LIVAE_ACT_RELU) a = fmaxf(a, 0.f);
        else if (act == LIVAE_ACT_SIGMOID) a = 1.f / (1.f + __expf(-a));
        f[i] = a;
      }
      if (relu_mask) {
        const uint4* mp = reinterpret_cast<const uint4*>(relu_mask + pix * Ntot + c0);
        uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
        const __nv_bfloat16* mb0 = reinterpret_cast<const __nv_bfloat16*>(&m0);
        const __nv_bfloat16* mb1 = reinterpret_cast<const __nv_bfloat16*>(&m1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!(__bfloat162float(mb0[i]) > 0.f)) f[i] = 0.f;
          if (!(__bfloat162float(mb1[i]) > 0.f)) f[8 + i] = 0.f;
        }
      }
      if (out_f32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * Ntot + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      } else {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * Ntot + c0);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
}

static constexpr int kThreads = 192;
static constexpr int kThreadsEpi2 = 320;   // un-blocking epilogue (EPI = 2): 8 epilogue warps, two per TMEM lane quarter

template <int STAGES>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)p.kc * 2u;
  const uint32_t a_bytes = (uint32_t)(p.nb * p.th * p.tw) * row_bytes;   // bytes one A box delivers
  const uint32_t b_bytes = (uint32_t)p.N * row_bytes;
  const uint32_t a_slot = (128u * row_bytes + 1023u) & ~1023u;
  const uint32_t b_slot = (b_bytes + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_slot + b_slot;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t ncols = 32;
  while (ncols < (uint32_t)p.N) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
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

  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y; tile /= p.tiles_y;
  const int b0 = tile * p.nb;
  const int nk_all = p.ntaps * p.nkc;
  // split-K (Linear layers: 16 pixel tiles but 256-512 K blocks): blockIdx.z owns K blocks [kb0, kb0 + nk)
  const int kper = (nk_all + p.ksplit - 1) / p.ksplit;
  const int kb0 = (int)blockIdx.z * kper;
  const int nk = min(kper, nk_all - kb0);

  if (warp == 0) {
    if (lane == 0) {
      const int x0 = tx * p.tw * p.in_stride, y0 = ty * p.th * p.in_stride;
      for (int ki = 0; ki < nk; ++ki) {
        const int kb = kb0 + ki;
        const int s = ki % STAGES;
        const uint32_t ph = (uint32_t)(ki / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes);
        const int tap = kb / p.nkc, c = kb - tap * p.nkc;
        uint8_t* a_s = smem + (uint32_t)s * stage_bytes;
        tma_load_4d(a_s, &tmA, &full_bar[s], c * p.kc, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b0);
        tma_load_3d(a_s + a_slot, &tmB, &full_bar[s], c * p.kc, (int)blockIdx.y * p.N, (int)p.tap_w[tap]);
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();      // whole-warp loop, one lane issues (tc_common.cuh)
      const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
      const uint32_t lt = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
      const uint32_t sbo = 8u * row_bytes;
      const int ksteps = p.kc / 16;
      for (int kb = 0; kb < nk; ++kb) {          // kb: index inside this CTA's K slice
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (uint32_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + a_slot;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = make_smem_desc(a_addr + (uint32_t)k * 32u, 16u, sbo, lt);
          const uint64_t bd = make_smem_desc(b_addr + (uint32_t)k * 32u, 16u, sbo, lt);
          if (leader) umma_f16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (leader) umma_commit(&empty_bar[s]);   // frees the smem slot once these MMAs have read it
      }
      if (leader) umma_commit(&accum_bar);        // accumulator complete
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int px = row % p.tw;
    const int t2 = row / p.tw;
    const int py = t2 % p.th;
    const int bi = t2 / p.th;
    const int qy = ty * p.th + py, qx = tx * p.tw + px;
    const bool valid = bi < p.nb && (b0 + bi) < p.B && qy < p.Hq && qx < p.Wq;
    const int64_t pix = ((int64_t)(b0 + bi) * p.Ho + (qy * p.os + p.oy0)) * p.Wo + (qx * p.os + p.ox0);
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    float* part = (p.ksplit > 1 && p.part) ? p.part + (int64_t)blockIdx.z * p.part_stride : nullptr;
    if (part && nk <= 0) {        // an empty K slice still owns its partial rows: they are summed unconditionally
      if (valid)
        for (int c = 0; c < p.N; ++c) part[pix * p.Ntot + (int)blockIdx.y * p.N + c] = 0.f;
    } else {
      epilogue_rows(tmem_base + ((uint32_t)(q * 32) << 16), (int)blockIdx.y * p.N, p.N, p.Ntot, valid && nk > 0, pix, p.out,
                    p.out_f32, p.bias, p.ksplit > 1 ? kActSplitK : p.act, p.relu_mask, part);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}


// ---------------------------------------------------------------------------------------------
// "Halo" variant: the taps of a convolution are row shifts of ONE input tile.  The CTA's output
// tile is 8 rows x 16 columns of the output grid, stored as 128 consecutive shared-memory rows
// (pixel-major); the input box [8 + hy rows] x [16 cols] x kc channels is fetched ONCE per
// (tap group, channel chunk) and each tap's A operand is the same buffer addressed from row
// (dy*16 + dx).  Columns x >= tw = 16 - hx read pixels of the next row and are discarded, so the
// MMA runs at tw/16 efficiency but L2->shared traffic drops by the number of taps per group
// (9x for 3x3, 25x for 5x5, 4x for the stride-2 4x4 layers, whose taps split into 4 parity groups).
// Two independent mbarrier rings: A boxes (2 slots) and per-tap weight tiles (4 slots).
struct HaloGroup { int16_t dy, dx, tap_begin, tap_end; };
struct ConvTcHaloParams {
  int tiles_x, tiles_y;
  int tw;                 // valid output columns per tile (bw - max column shift)
  int bw, th;             // box width in grid columns (16..20) and output rows per tile (128 / bw)
  int box_rows;           // rows of the input box (th + max row shift)
  int Hq, Wq, Ho, Wo, os, oy0, ox0, in_stride;
  int ngroups;
  HaloGroup grp[4];
  int ntaps;
  uint8_t tap_shift[kMaxTaps];   // shared-memory row offset of each tap inside its group's box
  int8_t tap_w[kMaxTaps];
  int kc, nkc, N, Ntot, B;
  int nsa;                // slots of the A ring
  int b_resident;         // 1: all ntaps*nkc weight tiles stay in shared memory for the CTA's lifetime
  void* out; int out_f32; const float* bias; int act; const __nv_bfloat16* relu_mask;
  long long* probe;
  int a_s2d, epi_mode; uint8_t* pool_idx;     // HaloOpts (tc_common.cuh)
  int s2d_cpr;                                // a_s2d: K chunks per pixel row of the 2x2 block (2C / kc)
  int up_cout;                                // epi_mode 3 (epilogue_upfold)
};

static constexpr int kHaloSmemMax = 225 * 1024;     // dynamic shared memory opt-in of the halo kernels (SM: 227 KB = 232448 B incl. ~1.3 KB static)
static constexpr int kSA = 2, kSAmax = 4, kSB = 4;   // A ring: kSA..kSAmax slots (p.nsa), as many as fit without costing a resident CTA

// Epilogue of one accumulator row per thread, 32 columns per pass: both tcgen05.ld and the ReLU-mask loads
// of the pass are in flight before the first use; bias comes from shared memory.
__device__ __forceinline__ void epilogue_rows32(uint32_t taddr, int nbase, int N, int Ntot, bool valid, int64_t pix,
                                                void* out, int out_f32, const float* __restrict__ sbias, int act,
                                                const __nv_bfloat16* __restrict__ relu_mask) {
  for (int cc = 0; cc < N; cc += 32) {
    const bool two = cc + 16 < N;                 // N is a multiple of 16: the last pass may be half
    uint32_t v[32];
    tmem_ld16(taddr + (uint32_t)cc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    if (two) tmem_ld16(taddr + (uint32_t)cc + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
    uint4 m[4];
    const int c0 = nbase + cc;
    if (relu_mask && valid) {
      const uint4* mp = reinterpret_cast<const uint4*>(relu_mask + pix * Ntot + c0);
      m[0] = __ldg(mp); m[1] = __ldg(mp + 1);
      if (two) { m[2] = __ldg(mp + 2); m[3] = __ldg(mp + 3); }
    }
    tmem_ld_wait();
    if (!valid) continue;
    const int nh = two ? 2 : 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h >= nh) break;
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(v[h * 16 + i]) + sbias[cc + h * 16 + i];
        if (act == LIVAE_ACT_RELU) a = fmaxf(a, 0.f);
        else if (act == LIVAE_ACT_SIGMOID) a = 1.f / (1.f + __expf(-a));
        f[i] = a;
      }
      if (relu_mask) {
        const __nv_bfloat16* mb0 = reinterpret_cast<const __nv_bfloat16*>(&m[2 * h]);
        const __nv_bfloat16* mb1 = reinterpret_cast<const __nv_bfloat16*>(&m[2 * h + 1]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!(__bfloat162float(mb0[i]) > 0.f)) f[i] = 0.f;
          if (!(__bfloat162float(mb1[i]) > 0.f)) f[8 + i] = 0.f;
        }
      }
      if (out_f32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + pix * Ntot + c0 + h * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
      } else {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * Ntot + c0 + h * 16);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
}

// The same for a bf16 output whose ReLU-mask vectors (N <= 128 columns: 16 x 8 bf16) are already in registers
__device__ __forceinline__ void epilogue_rows32_premask(uint32_t taddr, int nbase, int N, int Ntot, bool valid, int64_t pix,
                                                        void* out, const float* __restrict__ sbias, int act,
                                                        const uint4 (&m)[16]) {
#pragma unroll
  for (int p32 = 0; p32 < 4; ++p32) {
    const int cc = p32 * 32;
    if (cc >= N) break;
    const bool two = cc + 16 < N;
    uint32_t v[32];
    tmem_ld16(taddr + (uint32_t)cc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    if (two) tmem_ld16(taddr + (uint32_t)cc + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
    tmem_ld_wait();
    if (!valid) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m[p32 * 4 + h * 2 + j]);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float a0 = __uint_as_float(v[h * 16 + j * 8 + i]) + sbias[cc + h * 16 + j * 8 + i];
          float a1 = __uint_as_float(v[h * 16 + j * 8 + i + 1]) + sbias[cc + h * 16 + j * 8 + i + 1];
          if (act == LIVAE_ACT_RELU) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
          else if (act == LIVAE_ACT_SIGMOID) { a0 = 1.f / (1.f + __expf(-a0)); a1 = 1.f / (1.f + __expf(-a1)); }
          if (!(__bfloat162float(mb[i]) > 0.f)) a0 = 0.f;
          if (!(__bfloat162float(mb[i + 1]) > 0.f)) a1 = 0.f;
          __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
          w[j * 4 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&hh);
        }
      }
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * Ntot + nbase + cc + h * 16);
      o[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

// epi_mode 1: the N = 4*Co accumulator columns of a row are the 2x2 output pixels (phase = dy*2+dx, torch's
// max-pool scan order) of one block: bias + ReLU + max / argmax over the four phases.
__device__ __forceinline__ void epilogue_pool4(uint32_t taddr, int N, bool valid, int64_t pix, void* out,
                                               uint8_t* __restrict__ idx, const float* __restrict__ sbias) {
  const int Co = N >> 2;
  for (int cc = 0; cc < Co; cc += 16) {
    uint32_t v[4][16];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) tmem_ld16(taddr + (uint32_t)(ph * Co + cc), v[ph]);
    tmem_ld_wait();
    if (!valid) continue;
    uint32_t ow[8], iw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      float best[2]; uint32_t bi[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float b = sbias[cc + c + j];
        best[j] = fmaxf(__uint_as_float(v[0][c + j]) + b, 0.f); bi[j] = 0u;
#pragma unroll
        for (int ph = 1; ph < 4; ++ph) {
          const float a = fmaxf(__uint_as_float(v[ph][c + j]) + b, 0.f);
          if (a > best[j]) { best[j] = a; bi[j] = (uint32_t)ph; }
        }
      }
      __nv_bfloat162 hh = __floats2bfloat162_rn(best[0], best[1]);
      ow[c >> 1] = *reinterpret_cast<uint32_t*>(&hh);
      iw[c >> 2] |= (bi[0] << ((c & 3) * 8)) | (bi[1] << (((c & 3) + 1) * 8));
    }
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix * Co + cc);
    op[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    op[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
    *reinterpret_cast<uint4*>(idx + pix * Co + cc) = make_uint4(iw[0], iw[1], iw[2], iw[3]);
  }
}

// epi_mode 2: the N = 4*Ci columns of block (qy,qx) go back to the four pixels of a plain NHWC tensor
// [B, 2*Hq, 2*Wq, Ci] (bf16), each masked by relu_mask (same layout) > 0.  The mask vectors of phase p+1 are
// requested before phase p is processed (a global-load latency per 16 columns made this epilogue slower than
// the tile's MMAs).  CI16 = Ci / 16 (1, 2 or 4).
// The four phases are split between TWO warps per TMEM lane quarter (ph0 = 0 or 2): with one warp per quarter
// this epilogue took 4600-5200 cycles per tile against 2350 for the tile's MMAs (probe), i.e. the un-blocking
// data-gradient kernels were epilogue-bound.
template <int CI16, typename Wait>
__device__ __forceinline__ void epilogue_unblock(uint32_t taddr, bool valid, int b, int qy, int qx, int Hq, int Wq,
                                                 void* out, const __nv_bfloat16* __restrict__ relu_mask, int ph0,
                                                 Wait wait_accumulator) {
  constexpr int Ci = CI16 * 16;
  constexpr int NPRE = CI16 <= 2 ? 2 : 1;      // phases whose mask vectors are requested up front
  const bool use_mask = relu_mask != nullptr && valid;
  auto pix_of = [&](int ph) { return ((int64_t)b * 2 * Hq + 2 * qy + (ph >> 1)) * (2 * Wq) + 2 * qx + (ph & 1); };
  // The ReLU-mask vectors of this thread's pixels are requested BEFORE waiting for the accumulator: issued after
  // it, their global-load latency was exposed once per tile and doubled the kernel (0.37 -> 0.74 ms on the c2
  // data gradient with / without a mask).
  uint4 m[2][2 * CI16];
  if (use_mask) {
#pragma unroll
    for (int pi = 0; pi < NPRE; ++pi) {
      const uint4* mp = reinterpret_cast<const uint4*>(relu_mask + pix_of(ph0 + pi) * Ci);
#pragma unroll
      for (int i = 0; i < 2 * CI16; ++i) m[pi][i] = __ldg(mp + i);
    }
  }
  wait_accumulator();
#pragma unroll
  for (int pi = 0; pi < 2; ++pi) {
    const int ph = ph0 + pi;
    uint32_t v[CI16][16];
#pragma unroll
    for (int c = 0; c < CI16; ++c) tmem_ld16(taddr + (uint32_t)(ph * Ci + c * 16), v[c]);
    if (NPRE == 1 && pi == 0 && use_mask) {
      const uint4* mp = reinterpret_cast<const uint4*>(relu_mask + pix_of(ph0 + 1) * Ci);
#pragma unroll
      for (int i = 0; i < 2 * CI16; ++i) m[1][i] = __ldg(mp + i);
    }
    tmem_ld_wait();
    if (!valid) continue;
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + pix_of(ph) * Ci);
#pragma unroll
    for (int c = 0; c < CI16; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m[pi][2 * c + h]);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f0 = __uint_as_float(v[c][h * 8 + 2 * i]), f1 = __uint_as_float(v[c][h * 8 + 2 * i + 1]);
          if (relu_mask) {
            if (!(__bfloat162float(mb[2 * i]) > 0.f)) f0 = 0.f;
            if (!(__bfloat162float(mb[2 * i + 1]) > 0.f)) f1 = 0.f;
          }
          __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
          w[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        o[2 * c + h] = make_uint4(w[0], w[1], w[2], w[3]);
      }
  }
}

// epi_mode 3 (phase-folded Upsample(x2) -> ReflectionPad(1) -> Conv3x3 -> ReLU, csrc/upfold.cu): the N = 4*Cout
// accumulator columns of low-resolution pixel (qy,qx) are its four output pixels (2qy+py, 2qx+px), phase-major;
// out = ReLU(acc + bias) goes to the plain NHWC tensor [B, 2Hq, 2Wq, Cout].  The two outermost output rows / columns
// still lack their border correction: they are stored WITHOUT the ReLU and finished by upfold_ring_kernel (adding the
// correction here cost a dependent global-load round trip per 32 columns: 7900 instead of 2900 cycles per tile).
__device__ __forceinline__ void epilogue_upfold(uint32_t taddr, int nbase, int N, int Cout, bool valid, int b, int qy,
                                                int qx, int Hq, int Wq, __nv_bfloat16* __restrict__ out,
                                                const float* __restrict__ sbias) {
  const int Ho = 2 * Hq, Wo = 2 * Wq;
  for (int cc = 0; cc < N; cc += 32) {
    uint32_t v[32];
    tmem_ld16(taddr + (uint32_t)cc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    tmem_ld16(taddr + (uint32_t)cc + 16u, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
    const int n0 = nbase + cc;
    const int ph = n0 / Cout, c0 = n0 - ph * Cout;       // Cout is a multiple of 32: the 32 columns share a phase
    const int Y = 2 * qy + (ph >> 1), X = 2 * qx + (ph & 1);
    const float lo = (Y < 2 || Y >= Ho - 2 || X < 2 || X >= Wo - 2) ? -3.0e38f : 0.f;
    tmem_ld_wait();
    if (!valid) continue;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a0 = fmaxf(__uint_as_float(v[2 * i]) + sbias[cc + 2 * i], lo);
      const float a1 = fmaxf(__uint_as_float(v[2 * i + 1]) + sbias[cc + 2 * i + 1], lo);
      __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
      w[i] = *reinterpret_cast<uint32_t*>(&hh);
    }
    uint4* o = reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + Y) * Wo + X) * Cout + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

// Persistent: grid.x CTAs walk the tile list round-robin.  The accumulator is double-buffered in
// TMEM (2 x N columns), so the epilogue of tile i (tcgen05.ld, activation, global stores) overlaps the
// TMA + MMA main loop of tile i+1, and barrier/TMEM/tensor-map setup is paid once per CTA.
// The MMA issuer is ONE thread and every instruction of its loop is a dependent-latency step (measured
// with livae_set_probe: descriptor construction from kernel parameters cost ~450 cycles per MMA, 10x the
// MMA itself), so the loop is reduced to: one shared-memory load (tap offset, precomputed in descriptor
// units) + two 64-bit adds per tap, K steps unrolled at compile time.
// EPI = epilogue mode (HaloOpts): a template parameter so that the register-hungry pooled / un-blocking
// epilogues do not cost the plain convolutions their occupancy.
template <int KSTEPS, int EPI>
__global__ void __launch_bounds__(EPI == 2 ? kThreadsEpi2 : kThreads, EPI == 2 ? 2 : 1) conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const ConvTcHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t fullA[kSAmax], emptyA[kSAmax], fullB[kSB], emptyB[kSB], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_tapoff[kMaxTaps];        // row shift of each tap in descriptor address units (16 B)
  __shared__ int s_grp[4][2];                    // tap_begin, tap_end of each group
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t row_bytes = KSTEPS * 32u;
  const uint32_t a_bytes = (uint32_t)(p.box_rows * p.bw) * row_bytes;
  const uint32_t a_slot = ((uint32_t)(p.box_rows * p.bw + p.bw) * row_bytes + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.N * row_bytes;
  const uint32_t b_slot = (b_bytes + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smemB = smem + (uint32_t)p.nsa * a_slot;
  uint32_t acc_cols = 32;
  while (acc_cols < (uint32_t)p.N) acc_cols <<= 1;
  // Two TMEM accumulators (double buffer).  Up to four of them, a second epilogue warp group on alternate tiles and a
  // deeper A ring were measured on the folded-decoder kernels and on every kernel that has its SM to itself: no gain
  // (15.68 -> 15.79 ms per step) -- those epilogues are bound by LSU wavefronts (every thread stores its own 32..128-byte
  // run: 79 % of the LSU data pipe in the ncu capture of epilogue_upfold), which more buffering cannot hide.
  constexpr uint32_t nacc = 2u;
  const uint32_t ncols = nacc * acc_cols;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < kSAmax; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < kSB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI == 2 ? 8 : 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, ncols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.ntaps; i += blockDim.x) s_tapoff[i] = ((uint32_t)p.tap_shift[i] * row_bytes) >> 4;
  for (int i = threadIdx.x; i < p.ngroups; i += blockDim.x) { s_grp[i][0] = p.grp[i].tap_begin; s_grp[i][1] = p.grp[i].tap_end; }
  if (EPI == 1) {
    for (int i = threadIdx.x; i < (p.N >> 2); i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  } else if (EPI == 3) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[((int)blockIdx.y * p.N + i) % p.up_cout] : 0.f;
  } else {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[(int)blockIdx.y * p.N + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int total_tiles = p.tiles_x * p.tiles_y * p.B;
  const long long cta_t0 = p.probe ? clock64() : 0ll;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, ib = 0, pn = 0;
      uint32_t pha = 0u;
      if (p.b_resident) {   // small filters: fetch every weight tile once
        mbar_arrive_expect_tx(&fullB[0], b_bytes * (uint32_t)(p.ntaps * p.nkc));
        for (int t = 0; t < p.ntaps; ++t)
          for (int c = 0; c < p.nkc; ++c)
            tma_load_3d(smemB + (uint32_t)(t * p.nkc + c) * b_slot, &tmB, &fullB[0], c * p.kc, (int)blockIdx.y * p.N,
                        (int)p.tap_w[t]);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t3 = tile;
        const int tx = t3 % p.tiles_x; t3 /= p.tiles_x;
        const int ty = t3 % p.tiles_y; t3 /= p.tiles_y;
        const int b = t3;
        const int x0 = tx * p.tw * p.in_stride, y0 = ty * p.th * p.in_stride;
        for (int g = 0; g < p.ngroups; ++g) {
          const HaloGroup G = p.grp[g];
          for (int c = 0; c < p.nkc; ++c) {
            probe_rec(p.probe, 0, 0, pn);
            mbar_wait(&emptyA[sa], pha ^ 1u);
            probe_rec(p.probe, 0, 1, pn);
            mbar_arrive_expect_tx(&fullA[sa], a_bytes);
            if (p.a_s2d) tma_load_4d(smem + (uint32_t)sa * a_slot, &tmA, &fullA[sa], (c % p.s2d_cpr) * p.kc, x0 + G.dx,
                                     2 * (y0 + G.dy) + c / p.s2d_cpr, b);
            else tma_load_4d(smem + (uint32_t)sa * a_slot, &tmA, &fullA[sa], c * p.kc, x0 + G.dx, y0 + G.dy, b);
            if (++sa == p.nsa) { sa = 0; pha ^= 1u; }
            if (p.b_resident) continue;
            for (int t = G.tap_begin; t < G.tap_end; ++t) {
              const int sb = ib % kSB;
              mbar_wait(&emptyB[sb], ((uint32_t)(ib / kSB) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&fullB[sb], b_bytes);
              tma_load_3d(smemB + (uint32_t)sb * b_slot, &tmB, &fullB[sb], c * p.kc, (int)blockIdx.y * p.N,
                          (int)p.tap_w[t]);
              ++ib;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();      // whole-warp loop, one lane issues (tc_common.cuh)
      const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
      constexpr uint32_t lt = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
      constexpr uint32_t sbo = 8u * row_bytes;
      // The swizzle is a function of the absolute shared-memory address (measured: a start shifted by
      // whole rows needs NO descriptor base offset), so TMA's write pattern and the MMA's read pattern
      // agree for any row shift: a tap is "the A descriptor + its row shift".
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem), 16u, sbo, lt);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smemB), 16u, sbo, lt);
      const uint32_t a_slot16 = a_slot >> 4, b_slot16 = b_slot >> 4;
      const int ngroups = p.ngroups, nkc = p.nkc;
      const bool resident = p.b_resident != 0;
      uint32_t sa = 0, pha = 0, ib = 0, it = 0;
      const uint32_t nsa = (uint32_t)p.nsa;
      int pn = 0;
      if (resident) { mbar_wait(&fullB[0], 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it % nacc;
        if (leader) probe_rec(p.probe, 1, 0, pn);
        mbar_wait(&tempty[acc], ((it / nacc) & 1u) ^ 1u);   // epilogue has drained this buffer
        if (leader) probe_rec(p.probe, 1, 1, pn);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * acc_cols;
        uint32_t accum = 0u;
        for (int g = 0; g < ngroups; ++g) {
          const int tb = s_grp[g][0], te = s_grp[g][1];
          for (int c = 0; c < nkc; ++c) {
            mbar_wait(&fullA[sa], pha);
            if (leader) probe_rec(p.probe, 1, 2, pn);
            tc_fence_after();
            const uint64_t ad_s = adesc0 + sa * a_slot16;
            if (resident) {
              uint64_t bd = bdesc0 + (uint32_t)(tb * nkc + c) * b_slot16;
              const uint32_t bstep = (uint32_t)nkc * b_slot16;
              for (int t = tb; t < te; ++t, bd += bstep) {
                const uint64_t ad = ad_s + (((uint32_t)p.tap_shift[t] * row_bytes) >> 4);   // kernel parameter: stays uniform
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) { if (leader) umma_f16(d_addr, ad + 2u * k, bd + 2u * k, idesc, accum); accum = 1u; }
              }
            } else {
              for (int t = tb; t < te; ++t) {
                const uint32_t sb = ib & (kSB - 1);
                mbar_wait(&fullB[sb], (ib / kSB) & 1u);
                tc_fence_after();
                const uint64_t ad = ad_s + (((uint32_t)p.tap_shift[t] * row_bytes) >> 4);   // kernel parameter: stays uniform
                const uint64_t bd = bdesc0 + sb * b_slot16;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) { if (leader) umma_f16(d_addr, ad + 2u * k, bd + 2u * k, idesc, accum); accum = 1u; }
                if (leader) umma_commit(&emptyB[sb]);
                ++ib;
              }
            }
            if (leader) probe_rec(p.probe, 1, 3, pn);
            if (leader) umma_commit(&emptyA[sa]);
            if (++sa == nsa) { sa = 0; pha ^= 1u; }
          }
        }
        if (leader) umma_commit(&tfull[acc]);
        if (leader) probe_rec(p.probe, 1, 4, pn);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int py = row / p.bw, px = row - py * p.bw;
    int it = 0, pn = 0;
    long long* pb = (warp == 2 && lane == 0) ? p.probe : nullptr;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int t3 = tile;
      const int tx = t3 % p.tiles_x; t3 /= p.tiles_x;
      const int ty = t3 % p.tiles_y; t3 /= p.tiles_y;
      const int b = t3;
      const int acc = (int)((uint32_t)it % nacc);
      const int qy = ty * p.th + py, qx = tx * p.tw + px;
      const bool valid = px < p.tw && py < p.th && qy < p.Hq && qx < p.Wq;
      const int64_t pix = ((int64_t)b * p.Ho + (qy * p.os + p.oy0)) * p.Wo + (qx * p.os + p.ox0);
      probe_rec(pb, 2, 0, pn);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols;
      auto wait_acc = [&]() {
        mbar_wait(&tfull[acc], ((uint32_t)it / nacc) & 1u);
        probe_rec(pb, 2, 1, pn);
        tc_fence_after();
      };
      if (EPI == 0) {
        wait_acc();
        epilogue_rows32(taddr, (int)blockIdx.y * p.N, p.N, p.Ntot, valid, pix, p.out, p.out_f32, s_bias, p.act, p.relu_mask);
      } else if (EPI == 4) {
        {   // EPI 4 = EPI 0 for a bf16 output with a ReLU mask and N <= 128 (its own instantiation: 64 more registers)
          // all ReLU-mask vectors of the row are requested BEFORE waiting for the accumulator (as epilogue_unblock
          // does): issued per 32-column pass after it, each pass exposed a global-load round trip (~2500 cycles per
          // pass on the folded decoder's data gradients, more than the tile's MMAs)
          uint4 m[16];
          const int nv = p.N >> 3;
          if (valid) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.relu_mask + pix * p.Ntot + (int)blockIdx.y * p.N);
#pragma unroll
            for (int i = 0; i < 16; ++i) if (i < nv) m[i] = __ldg(mp + i);
          }
          wait_acc();
          epilogue_rows32_premask(taddr, (int)blockIdx.y * p.N, p.N, p.Ntot, valid, pix, p.out, s_bias, p.act, m);
        }
      } else if (EPI == 1) {
        wait_acc();
        epilogue_pool4(taddr, p.N, valid, ((int64_t)b * p.Hq + qy) * p.Wq + qx, p.out, p.pool_idx, s_bias);
      } else if (EPI == 3) {
        wait_acc();
        epilogue_upfold(taddr, (int)blockIdx.y * p.N, p.N, p.up_cout, valid, b, qy, qx, p.Hq, p.Wq,
                        reinterpret_cast<__nv_bfloat16*>(p.out), s_bias);
      } else if (p.N == 64)
        epilogue_unblock<1>(taddr, valid, b, qy, qx, p.Hq, p.Wq, p.out, p.relu_mask, warp >= 6 ? 2 : 0, wait_acc);
      else      // N = 128 (the launcher admits 64 and 128 only: at N = 256 the mask registers spill)
        epilogue_unblock<2>(taddr, valid, b, qy, qx, p.Hq, p.Wq, p.out, p.relu_mask, warp >= 6 ? 2 : 0, wait_acc);
      probe_rec(pb, 2, 2, pn);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  __syncthreads();
  if (p.probe && threadIdx.x == 0 && blockIdx.x < 1024 && blockIdx.y == 0)     // role 3: whole-CTA span of every CTA
    p.probe[3 * 1024 + blockIdx.x] = clock64() - cta_t0;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// w fp32 [Cs][Cb][kh][kw] (torch) -> bf16
//   mode 0: out[tap][cs][cb]                      (P1: big -> small, K = cb)
//   mode 1: out[tap'][cb][cs], tap' = flipped tap (P2 at stride 1 run as a P1 over the small side)
//   mode 2: out[tap][cb][cs]                      (data gradient / ConvTranspose2d forward, K = cs)
//   mode 3: Linear over an NHWC-flattened map, forward:  out[cs][tap][cb] = w[cs][cb][tap]
//           (one "tap", K = taps*Cb in (h,w,c) order; torch keeps the K axis in (c,h,w) order)
//   mode 4: the same Linear, data gradient:              out[tap][cb][cs] (N = taps*Cb rows of K = cs)
// Rows cs >= Cs_real (padding up to Cs) are written as zeros (modes 3, 4).
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cs, int Cb, int kh, int kw, int mode,
                                    int Cs_real, __nv_bfloat16* __restrict__ out) {
  int n = Cs * Cb * kh * kw;
  int taps = kh * kw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int tap, cs, cb;
    if (mode == 0) { cb = i % Cb; int t = i / Cb; cs = t % Cs; tap = t / Cs; }
    else if (mode == 3) { cb = i % Cb; int t = i / Cb; tap = t % taps; cs = t / taps; }
    else { cs = i % Cs; int t = i / Cs; cb = t % Cb; tap = mode == 1 ? taps - 1 - t / Cb : t / Cb; }
    out[i] = cs < Cs_real ? __float2bfloat16_rn(w[((int64_t)cs * Cb + cb) * taps + tap]) : __float2bfloat16_rn(0.f);
  }
}

// out[r][c] = act(sum + bias[c]): finishes a split-K accumulation.  part == null: the sum is already in out (atomics);
// else sum = part[0][i] + part[1][i] + ... in slice order (bit-reproducible)
__global__ void bias_act_rows_kernel(float* __restrict__ out, const float* __restrict__ bias, int64_t n, int C, int act,
                                     const float* __restrict__ part, int nparts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a;
    if (part) {          // four independent chains (a serial chain waits one L2 round trip per slice), fixed association
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int z = 0;
      for (; z + 3 < nparts; z += 4) {
        a0 += part[(int64_t)z * n + i]; a1 += part[(int64_t)(z + 1) * n + i];
        a2 += part[(int64_t)(z + 2) * n + i]; a3 += part[(int64_t)(z + 3) * n + i];
      }
      for (; z < nparts; ++z) a0 += part[(int64_t)z * n + i];
      a = (a0 + a1) + (a2 + a3);
    } else {
      a = out[i];
    }
    a += bias ? __ldg(bias + (int)(i % C)) : 0.f;
    if (act == LIVAE_ACT_RELU) a = fmaxf(a, 0.f);
    else if (act == LIVAE_ACT_SIGMOID) a = 1.f / (1.f + __expf(-a));
    out[i] = a;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ y) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, float* __restrict__ y) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;

namespace livae {
namespace tc {

// Pick the (nb, th, tw) pixel tile (<= 128 pixels = UMMA M rows) that covers an Hq x Wq grid with
// the fewest tiles; ragged edges are handled by TMA zero fill + predicated stores.
static void choose_tile(int Hq, int Wq, int B, int* tw, int* th, int* nb) {
  if (Hq * Wq <= 64) {
    *tw = Wq; *th = Hq;
    int n = 128 / (Hq * Wq);
    if (n > B) n = B;
    *nb = n < 1 ? 1 : n;
    return;
  }
  *nb = 1;
  int best = 1 << 30, btw = 1, bth = 1;
  for (int w = 1; w <= (Wq < 128 ? Wq : 128); ++w) {
    int h = 128 / w;
    if (h > Hq) h = Hq;
    if (h < 1) continue;
    int tiles = ((Wq + w - 1) / w) * ((Hq + h - 1) / h);
    if (tiles < best || (tiles == best && w > btw)) { best = tiles; btw = w; bth = h; }
  }
  *tw = btw; *th = bth;
}

static bool channels_ok(int Cin, int Cout) {
  if (Cout % 16 != 0 || Cout < 16) return false;
  if (Cin >= 64) return Cin % 64 == 0;
  return Cin == 16 || Cin == 32;
}

// 0 = one TMA box per tap (conv_tc_kernel); 1 (default) = one haloed box per tap group (conv_tc_halo_kernel), box width
// chosen per layer; 2 = the same with 16-column boxes only (A/B switch for the box-width choice)
static int g_halo_mode = 1;
static int g_halo_wide = 1;

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Box width of the halo kernel's tile for an Hq x Wq output grid whose taps shift by up to max_sx columns inside a tap
// group: the width in 16..20 that covers the grid with the fewest 128-position tiles (128 / bw rows of bw - max_sx valid
// columns each); ties go to the narrower box.  Pure host arithmetic (livae_tc_halo_geometry exposes it to the CPU tests).
static int halo_box_width(int Hq, int Wq, int max_sx) {
  int best_bw = 16, best_tiles = 1 << 30;
  for (int bw = 16; bw <= 20; ++bw) {
    const int tw = bw - max_sx, th = 128 / bw;
    if (tw < 1) continue;
    const int tiles = ((Wq + tw - 1) / tw) * ((Hq + th - 1) / th);
    if (tiles < best_tiles) { best_tiles = tiles; best_bw = bw; }
  }
  return best_bw;
}

int halo_box_width_public(int Hq, int Wq, int max_sx) { return halo_box_width(Hq, Wq, max_sx); }

// returns 1 when the shape is not eligible for the halo kernel
int launch_conv_tc_halo(const void* in, int B, int Hin, int Win, int Cin, const void* wpacked, int wtaps, int N,
                        int Hq, int Wq, int Ho, int Wo, int os, int oy0, int ox0, int in_stride, int ntaps,
                        const int* tdy, const int* tdx, const int* tw_idx, void* out, int out_f32,
                        const float* bias, int act, const void* relu_mask, cudaStream_t st, HaloOpts opts) {
  ConvTcHaloParams p;
  p.a_s2d = opts.a_s2d; p.epi_mode = opts.epi_mode; p.pool_idx = opts.pool_idx;
  p.s2d_cpr = 1; p.up_cout = opts.up_cout;
  if (opts.a_s2d && ((Cin != 64 && Cin % 128 != 0) || in_stride != 1)) return 1;
  if ((opts.epi_mode == 1 || opts.epi_mode == 2) && N != 64 && N != 128 && N != 256) return 1;
  if (opts.epi_mode == 2 && N == 256) return 1;
  if (opts.epi_mode == 3 && (opts.up_cout % 32 != 0 || N != 4 * opts.up_cout || Cin % 64 != 0 || out_f32)) return 1;
  const int s = in_stride;
  // group taps by the parity class of their input offset; inside a group taps are whole-row/col shifts
  int gkey[4][2]; int ng = 0;
  int order[kMaxTaps], gof[kMaxTaps], n = 0;
  int gmin_y[4], gmin_x[4];
  for (int t = 0; t < ntaps; ++t) {
    int ry = ((tdy[t] % s) + s) % s, rx = ((tdx[t] % s) + s) % s;
    int g = -1;
    for (int i = 0; i < ng; ++i) if (gkey[i][0] == ry && gkey[i][1] == rx) g = i;
    if (g < 0) { if (ng == 4) return 1; g = ng++; gkey[g][0] = ry; gkey[g][1] = rx; gmin_y[g] = tdy[t]; gmin_x[g] = tdx[t]; }
    if (tdy[t] < gmin_y[g]) gmin_y[g] = tdy[t];
    if (tdx[t] < gmin_x[g]) gmin_x[g] = tdx[t];
    gof[t] = g;
  }
  int max_sy = 0, max_sx = 0;
  int tsy[kMaxTaps], tsx[kMaxTaps];
  for (int g = 0; g < ng; ++g) {
    p.grp[g].tap_begin = (int16_t)n;
    for (int t = 0; t < ntaps; ++t) {
      if (gof[t] != g) continue;
      int sy = (tdy[t] - gmin_y[g]) / s, sx = (tdx[t] - gmin_x[g]) / s;
      if (sy > max_sy) max_sy = sy;
      if (sx > max_sx) max_sx = sx;
      order[n] = t;
      tsy[n] = sy; tsx[n] = sx;
      p.tap_w[n] = (int8_t)tw_idx[t];
      ++n;
    }
    p.grp[g].tap_end = (int16_t)n;
    p.grp[g].dy = (int16_t)gmin_y[g]; p.grp[g].dx = (int16_t)gmin_x[g];
  }
  (void)order; (void)floordiv;
  if (max_sx > 8) return 1;
  p.ngroups = ng; p.ntaps = ntaps;
  // Box width: the tile is 128 consecutive positions of a bw-wide box, i.e. 128 / bw rows of bw - max_sx valid columns.
  // 16 (8 rows) is the default; a wider box (7 or 6 rows) is taken when it covers the grid with fewer tiles -- a 32-wide
  // grid under 3x3 taps needs three 14-column tiles per row at bw = 16 (12 tiles per 32 x 32 map) but two 16-column
  // tiles at bw = 18 (10 tiles).
  p.bw = g_halo_wide ? halo_box_width(Hq, Wq, max_sx) : 16;
  p.th = 128 / p.bw;
  p.tw = p.bw - max_sx;
  p.box_rows = p.th + max_sy;
  if (max_sy * p.bw + max_sx > 200) return 1;
  for (int t = 0; t < ntaps; ++t) p.tap_shift[t] = (uint8_t)(tsy[t] * p.bw + tsx[t]);
  if ((p.box_rows * s) > 256 || p.bw * s > 256) return 1;
  p.tiles_x = (Wq + p.tw - 1) / p.tw; p.tiles_y = (Hq + p.th - 1) / p.th;
  p.Hq = Hq; p.Wq = Wq; p.Ho = Ho; p.Wo = Wo; p.os = os; p.oy0 = oy0; p.ox0 = ox0; p.in_stride = s;
  p.kc = Cin >= 64 ? 64 : Cin;
  if (opts.a_s2d) {                        // K chunks never straddle a pixel row of the 2x2 block (see the tensor map below)
    p.kc = Cin / 2 < 64 ? Cin / 2 : 64;
    p.s2d_cpr = (Cin / 2) / p.kc;
  }
  p.nkc = Cin / p.kc;
  int nchunk = N;
  if (N > 256) { nchunk = 256; while (N % nchunk != 0) nchunk -= 16; }
  p.N = nchunk; p.Ntot = N; p.B = B;
  p.out = out; p.out_f32 = out_f32; p.bias = bias; p.act = act; p.relu_mask = (const __nv_bfloat16*)relu_mask;
  p.probe = g_probe;
  const int row_bytes = p.kc * 2;
  CUtensorMap tmA, tmB;
  if (opts.a_s2d) {
    // Plain NHWC [B, 2*Hin, 2*Win, C] (C = Cin / 4) read block-wise.  K chunk ey (= pixel row inside the
    // block) is a box over {(ex, c): 2C contiguous elements, X: stride 2C, pixel row r = 2*Y + ey walked
    // with element stride 2, B}.  (A single 5-D box {2C, ey, X, Y, B} does NOT work: with a swizzled map
    // TMA lays every innermost-dimension run on its own 128-byte shared-memory row -- measured -- so the two
    // halves of a block would not be contiguous.)
    const uint64_t C = (uint64_t)Cin / 4, Wf = 2 * (uint64_t)Win, Hf = 2 * (uint64_t)Hin;
    uint64_t dims[4] = {2 * C, (uint64_t)Win, Hf, (uint64_t)B};
    uint64_t str[3] = {2 * C * 2, Wf * C * 2, Hf * Wf * C * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.bw, (uint32_t)(2 * p.box_rows), 1u};
    uint32_t es[4] = {1, 1, 2, 1};
    if (int e = make_tmap_bf16(&tmA, in, 4, dims, str, box, es, row_bytes)) return e;
  } else {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)(p.bw * s), (uint32_t)(p.box_rows * s), 1u};
    uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
    if (int e = make_tmap_bf16(&tmA, in, 4, dims, str, box, es, row_bytes)) return e;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)N, (uint64_t)wtaps};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)N * Cin * 2};
    uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)p.N, 1};
    if (int e = make_tmap_bf16(&tmB, wpacked, 3, dims, str, box, nullptr, row_bytes)) return e;
  }
  const uint32_t a_slot = ((uint32_t)(p.box_rows * p.bw + p.bw) * row_bytes + 1023u) & ~1023u;
  const uint32_t b_slot = ((uint32_t)p.N * row_bytes + 1023u) & ~1023u;
  // keep every weight tile in shared memory for the CTA's lifetime whenever they fit next to the two A slots
  // (even at one CTA per SM: streaming them costs an mbarrier wait + a commit per tap, more than the tap's MMAs)
  p.b_resident = ((size_t)ntaps * p.nkc * b_slot + (size_t)kSA * a_slot + 1024 <= 200 * 1024) ? 1 : 0;
  size_t smem = (size_t)kSA * a_slot + (size_t)(p.b_resident ? ntaps * p.nkc : kSB) * b_slot + 1024;
  if (smem > 200 * 1024) return 1;
  p.nsa = kSA;
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(conv_tc_halo_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
    cudaFuncSetAttribute(conv_tc_halo_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemMax);
  }
  // plain epilogue with the ReLU-mask vectors prefetched (see EPI 4 in the kernel)
  const bool premask = opts.epi_mode == 0 && relu_mask != nullptr && !out_f32 && p.N <= 128 && p.kc == 64 && opts.a_s2d;
  const int tiles = p.tiles_x * p.tiles_y * B;
  // persistent grid: as many CTAs per SM as shared memory and the 512 TMEM columns allow
  uint32_t acc_cols = 32;
  while (acc_cols < (uint32_t)p.N) acc_cols <<= 1;
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  int per_sm_tmem = (int)(512 / (2 * acc_cols));
  if (per_sm > per_sm_tmem) per_sm = per_sm_tmem;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  // deepen the A ring while the same number of CTAs still fits (two slots leave the producer one box ahead
  // at most: the MMA warp waited 400-900 cycles per box for TMA latency)
  // (up to the whole 227 KB of the SM when the CTA is alone on it: with resident weights of 144 KB the folded decoder
  // kernels had two slots, i.e. ONE box in flight, and ran at one TMA round trip -- 4000+ cycles -- per tile)
  while (p.nsa < kSAmax && (smem + a_slot + 1024) * per_sm <= 220 * 1024 && smem + a_slot <= (size_t)kHaloSmemMax) { smem += a_slot; ++p.nsa; }
  int gx = kNumSMs * per_sm;
  if (gx > tiles) gx = tiles;
  const dim3 grid(gx, N / p.N);
  if (opts.epi_mode == 1) {
    if (p.kc != 32) return 1;
    conv_tc_halo_kernel<2, 1><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  } else if (opts.epi_mode == 2) {
    if (p.kc != 64) return 1;
    conv_tc_halo_kernel<4, 2><<<grid, kThreadsEpi2, smem, st>>>(tmA, tmB, p);
  } else if (opts.epi_mode == 3) {
    if (p.kc != 64) return 1;
    conv_tc_halo_kernel<4, 3><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  } else if (premask) conv_tc_halo_kernel<4, 4><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  else if (p.kc == 64) conv_tc_halo_kernel<4, 0><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  else if (p.kc == 32) conv_tc_halo_kernel<2, 0><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  else conv_tc_halo_kernel<1, 0><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// One launch: out[b, q*os+o0, :, N] = epilogue( sum_taps in[b, q*in_stride + d(tap), :, Cin] * W[wtap][N][Cin] )
static int launch_conv_tc(const void* in, int B, int Hin, int Win, int Cin, const void* wpacked, int wtaps, int N,
                          int Hq, int Wq, int Ho, int Wo, int os, int oy0, int ox0, int in_stride, int ntaps,
                          const int* tdy, const int* tdx, const int* tw_idx, void* out, int out_f32,
                          const float* bias, int act, const void* relu_mask, cudaStream_t st) {
  LIVAE_CHECK_ARG(ntaps >= 1 && ntaps <= kMaxTaps, "tc_conv: too many taps (%d)", ntaps);
  if (g_halo_mode != 0 && ntaps > 1 && Wq >= 8 && Hq >= 4) {
    int rc = launch_conv_tc_halo(in, B, Hin, Win, Cin, wpacked, wtaps, N, Hq, Wq, Ho, Wo, os, oy0, ox0, in_stride,
                                 ntaps, tdy, tdx, tw_idx, out, out_f32, bias, act, relu_mask, st, HaloOpts{0, 0, nullptr, 0});
    if (rc != 1) return rc;   // 1 = shape not eligible, fall through to the per-tap kernel
  }
  ConvTcParams p;
  choose_tile(Hq, Wq, B, &p.tw, &p.th, &p.nb);
  LIVAE_CHECK_ARG(p.tw * in_stride <= 256 && p.th * in_stride <= 256, "tc_conv: TMA box too large");
  p.tiles_x = (Wq + p.tw - 1) / p.tw; p.tiles_y = (Hq + p.th - 1) / p.th;
  p.Hq = Hq; p.Wq = Wq; p.Ho = Ho; p.Wo = Wo; p.os = os; p.oy0 = oy0; p.ox0 = ox0; p.in_stride = in_stride;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { p.tap_dy[t] = (int8_t)tdy[t]; p.tap_dx[t] = (int8_t)tdx[t]; p.tap_w[t] = (int8_t)tw_idx[t]; }
  p.kc = Cin >= 64 ? 64 : Cin;
  p.nkc = Cin / p.kc;
  int nchunk = N;
  if (N > 256) {
    nchunk = 256;
    while (N % nchunk != 0) nchunk -= 16;
  }
  p.N = nchunk; p.Ntot = N; p.B = B;
  p.out = out; p.out_f32 = out_f32; p.bias = bias; p.act = act;
  p.relu_mask = (const __nv_bfloat16*)relu_mask;
  const int row_bytes = p.kc * 2;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)(p.tw * in_stride), (uint32_t)(p.th * in_stride), (uint32_t)p.nb};
    uint32_t es[4] = {1, (uint32_t)in_stride, (uint32_t)in_stride, 1};
    if (int e = make_tmap_bf16(&tmA, in, 4, dims, str, box, es, row_bytes)) return e;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)N, (uint64_t)wtaps};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)N * Cin * 2};
    uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)p.N, 1};
    if (int e = make_tmap_bf16(&tmB, wpacked, 3, dims, str, box, nullptr, row_bytes)) return e;
  }
  const uint32_t a_slot = (128u * row_bytes + 1023u) & ~1023u;
  const uint32_t b_slot = ((uint32_t)p.N * row_bytes + 1023u) & ~1023u;
  constexpr int STAGES = 4;
  const size_t smem = (size_t)STAGES * (a_slot + b_slot) + 1024;
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(conv_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int tiles = p.tiles_x * p.tiles_y * ((B + p.nb - 1) / p.nb);
  // split-K when the output grid cannot fill the chip but K is long (nn.Linear over a flattened map: 16 tiles,
  // 256-512 K blocks -- 16 CTAs streamed the 134 MB operand at 0.7 TB/s)
  p.ksplit = 1;
  const int nk = ntaps * p.nkc, ctas = tiles * (N / p.N);
  if (out_f32 && !relu_mask && ctas * 4 <= kNumSMs && nk >= 32) {
    int ks = (2 * kNumSMs + ctas - 1) / ctas;
    if (ks > nk / 8) ks = nk / 8;
    const int kper = (nk + ks - 1) / ks;
    p.ksplit = (nk + kper - 1) / kper;
  }
  p.part = nullptr; p.part_stride = 0;
  if (p.ksplit > 1) {
    const int64_t n = (int64_t)B * Ho * Wo * N;
    p.part = scratch_floats((int64_t)p.ksplit * n);
    p.part_stride = n;
    if (!p.part) {
      cudaError_t ce = cudaMemsetAsync(out, 0, (size_t)n * sizeof(float), st);
      if (ce != cudaSuccess) { set_error("tc_conv split-K memset: %s", cudaGetErrorString(ce)); return (int)ce; }
    }
    conv_tc_kernel<STAGES><<<dim3(tiles, N / p.N, p.ksplit), kThreads, smem, st>>>(tmA, tmB, p);
    LIVAE_CUDA_LAUNCH_CHECK();
    if (p.part || bias || act != LIVAE_ACT_NONE) {
      bias_act_rows_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>((float*)out, bias, n, N, act,
                                                                                              p.part, p.ksplit);
      LIVAE_CUDA_LAUNCH_CHECK();
    }
    return 0;
  }
  conv_tc_kernel<STAGES><<<dim3(tiles, N / p.N), kThreads, smem, st>>>(tmA, tmB, p);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace livae

using namespace livae;
using namespace livae::tc;

// Can the tensor-core engine run this convolution (forward AND both gradients)?
extern "C" int livae_tc_conv_supported(const livae_tc_conv_desc* d) {
  if (!d) return 0;
  if (!channels_ok(d->Cin, d->Cout)) return 0;
  if (d->stride != 1 && d->stride != 2) return 0;
  if (d->kh * d->kw > kMaxTaps) return 0;
  int Ho = (d->Hin + 2 * d->pad - d->kh) / d->stride + 1, Wo = (d->Win + 2 * d->pad - d->kw) / d->stride + 1;
  if (Ho <= 0 || Wo <= 0) return 0;
  if (d->stride == 2 && ((d->Hin | d->Win) & 1)) return 0;
  return 1;
}

extern "C" int livae_tc_pack_weights(const float* w, int Cs, int Cb, int kh, int kw, int mode, int Cs_real,
                                     void* out_bf16, livae_stream_t stream) {
  LIVAE_CHECK_ARG(w && out_bf16 && Cs > 0 && Cb > 0 && kh > 0 && kw > 0 && mode >= 0 && mode <= 4 &&
                      Cs_real > 0 && Cs_real <= Cs && (Cs_real == Cs || mode >= 3),
                  "tc_pack_weights: bad args");
  if (int e = require_sm100()) return e;
  int n = Cs * Cb * kh * kw;
  pack_weights_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Cs, Cb, kh, kw, mode == 4 ? 2 : mode,
                                                                          Cs_real, (__nv_bfloat16*)out_bf16);
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_cast(const void* src, int dt_src, void* dst, int dt_dst, int64_t n, livae_stream_t stream) {
  LIVAE_CHECK_ARG(n >= 0, "cast: bad size");
  if (n == 0) return 0;
  LIVAE_CHECK_ARG(src && dst, "cast: null pointer");
  if (int e = require_sm100()) return e;
  int64_t blocks = (n + 1023) / 1024;
  int grid = (int)(blocks < kNumSMs * 16 ? blocks : kNumSMs * 16);
  if (dt_src == LIVAE_F32 && dt_dst == LIVAE_BF16)
    cast_f32_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, n, (__nv_bfloat16*)dst);
  else if (dt_src == LIVAE_BF16 && dt_dst == LIVAE_F32)
    cast_bf16_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, n, (float*)dst);
  else { set_error("cast: unsupported dtype pair %d -> %d", dt_src, dt_dst); return -1; }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// y[B,Ho,Wo,Cout] = act(conv(x[B,Hin,Win,Cin] (bf16), wpacked[tap][Cout][Cin] (bf16)) + bias) [* (mask > 0)]
extern "C" int livae_tc_conv(const livae_tc_conv_desc* d, const void* x, const void* wpacked, const float* bias,
                             void* y, const void* relu_mask, livae_stream_t stream) {
  LIVAE_CHECK_ARG(d, "tc_conv: null descriptor");
  if (d->B == 0) return 0;
  LIVAE_CHECK_ARG(x && wpacked && y, "tc_conv: null pointer");
  LIVAE_CHECK_ARG(channels_ok(d->Cin, d->Cout) && (d->stride == 1 || d->stride == 2) && d->kh * d->kw <= kMaxTaps,
                  "tc_conv: shape not supported by the tensor-core engine");
  LIVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)wpacked | (uintptr_t)y | (uintptr_t)relu_mask) & 15) == 0,
                  "tc_conv: pointers must be 16-byte aligned");
  if (int e = require_sm100()) return e;
  const int Ho = (d->Hin + 2 * d->pad - d->kh) / d->stride + 1, Wo = (d->Win + 2 * d->pad - d->kw) / d->stride + 1;
  LIVAE_CHECK_ARG(Ho > 0 && Wo > 0, "tc_conv: empty output");
  int tdy[kMaxTaps], tdx[kMaxTaps], tw[kMaxTaps];
  for (int ky = 0; ky < d->kh; ++ky)
    for (int kx = 0; kx < d->kw; ++kx) {
      int t = ky * d->kw + kx;
      tdy[t] = ky - d->pad; tdx[t] = kx - d->pad; tw[t] = t;
    }
  return launch_conv_tc(x, d->B, d->Hin, d->Win, d->Cin, wpacked, d->kh * d->kw, d->Cout, Ho, Wo, Ho, Wo, 1, 0, 0,
                        d->stride, d->kh * d->kw, tdy, tdx, tw, y, d->out_f32, bias, d->act, relu_mask,
                        (cudaStream_t)stream);
}

// Data gradient of the convolution described by d (also nn.ConvTranspose2d forward, model.py:90-96):
//   gx[B,Hin,Win,Cin] = act( sum_{ky,kx,co} gy[b,(iy+pad-ky)/s,(ix+pad-kx)/s,co] * w[co,ci,ky,kx] + bias ) [* (mask>0)]
// gy: bf16 [B,Ho,Wo,Cout]; wpacked: mode-2 packing [tap][Cin][Cout].  Stride 2 runs as four
// output-parity phases, each a stride-1 convolution over gy with the taps of matching parity.
extern "C" int livae_tc_conv_dgrad(const livae_tc_conv_desc* d, const void* gy, const void* wpacked,
                                   const float* bias, void* gx, const void* relu_mask, livae_stream_t stream) {
  LIVAE_CHECK_ARG(d, "tc_conv_dgrad: null descriptor");
  if (d->B == 0) return 0;
  LIVAE_CHECK_ARG(gy && wpacked && gx, "tc_conv_dgrad: null pointer");
  LIVAE_CHECK_ARG(channels_ok(d->Cout, d->Cin) && (d->stride == 1 || d->stride == 2) && d->kh * d->kw <= kMaxTaps,
                  "tc_conv_dgrad: shape not supported by the tensor-core engine");
  LIVAE_CHECK_ARG((((uintptr_t)gy | (uintptr_t)wpacked | (uintptr_t)gx | (uintptr_t)relu_mask) & 15) == 0,
                  "tc_conv_dgrad: pointers must be 16-byte aligned");
  if (int e = require_sm100()) return e;
  const int s = d->stride;
  const int Ho = (d->Hin + 2 * d->pad - d->kh) / s + 1, Wo = (d->Win + 2 * d->pad - d->kw) / s + 1;
  LIVAE_CHECK_ARG(Ho > 0 && Wo > 0, "tc_conv_dgrad: empty output");
  LIVAE_CHECK_ARG(s == 1 || ((d->Hin % 2) == 0 && (d->Win % 2) == 0), "tc_conv_dgrad: stride 2 needs even input size");
  for (int py = 0; py < s; ++py)
    for (int px = 0; px < s; ++px) {
      int tdy[kMaxTaps], tdx[kMaxTaps], tw[kMaxTaps], nt = 0;
      for (int ky = 0; ky < d->kh; ++ky) {
        if ((py + d->pad - ky) % s != 0) continue;
        for (int kx = 0; kx < d->kw; ++kx) {
          if ((px + d->pad - kx) % s != 0) continue;
          tdy[nt] = (py + d->pad - ky) / s; tdx[nt] = (px + d->pad - kx) / s; tw[nt] = ky * d->kw + kx;
          ++nt;
        }
      }
      LIVAE_CHECK_ARG(nt > 0, "tc_conv_dgrad: a phase without taps is not supported");
      if (int e = launch_conv_tc(gy, d->B, Ho, Wo, d->Cout, wpacked, d->kh * d->kw, d->Cin, d->Hin / s, d->Win / s,
                                 d->Hin, d->Win, s, py, px, 1, nt, tdy, tdx, tw, gx, d->out_f32, bias, d->act,
                                 relu_mask, (cudaStream_t)stream))
        return e;
    }
  return 0;
}

// tracing hook (tc_common.cuh): device buffer of 4 x 1024 int64, or NULL to switch tracing off
extern "C" void livae_set_probe(void* dev_ptr) { g_probe = (long long*)dev_ptr; }

// Host-side tile geometry of the halo kernel (no device work): for an Hq x Wq output grid and a maximum column shift
// max_sx inside a tap group (kw - 1 for stride 1) -> box width, rows and valid columns per tile, tiles per map
extern "C" int livae_tc_halo_geometry(int Hq, int Wq, int max_sx, int* bw, int* th, int* tw, int* tiles) {
  LIVAE_CHECK_ARG(Hq > 0 && Wq > 0 && max_sx >= 0 && max_sx <= 8 && bw && th && tw && tiles, "tc_halo_geometry: bad args");
  const int b = livae::tc::halo_box_width_public(Hq, Wq, max_sx);
  *bw = b; *th = 128 / b; *tw = b - max_sx;
  *tiles = ((Wq + *tw - 1) / *tw) * ((Hq + *th - 1) / *th);
  return 0;
}

// tuning / test hook: 0 = per-tap boxes only, 1 = halo kernel where eligible
extern "C" void livae_tc_set_halo_mode(int mode) { g_halo_mode = mode ? 1 : 0; g_halo_wide = mode == 2 ? 0 : 1; }

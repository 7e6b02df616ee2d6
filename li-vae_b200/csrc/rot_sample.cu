// a6: fused rotate + bilinear sample == F.affine_grid + F.grid_sample(bilinear, reflection,
// align_corners=False) at reference model.py:254-258, model.py:467-470, train.py:675-677.
//
// Memory-bound.  Algorithmic bytes per [H,W] fp32 image: forward read H*W*4 + write H*W*4;
// backward read gout + read img (+ write gimg).  No [B,H,W,2] grid tensor is ever
// materialised: the source coordinate of every output pixel is recomputed in registers
// from the per-sample (cos, sin).
//
// Forward: one CTA per (sample, channel) image.  The source image is staged ONCE into shared
// memory with coalesced float4 loads (64 KB at 128x128, 3 CTAs resident per SM), the four
// bilinear taps are shared-memory gathers, the output is written with coalesced float4 stores.
//
// Backward: one CTA per sample.  grad_input is a scatter; instead of global atomics the CTA
// accumulates into a shared-memory tile (shared atomics, no HBM traffic) and writes the tile out
// once with float4 stores, so HBM sees exactly one coalesced write of grad_input.  The per-sample
// dL/d(cos), dL/d(sin) are reduced with warp shuffles + one shared exchange.
#include "common.cuh"

namespace livae {

struct SrcCoord {
  float v;     // source coordinate after reflect + clip
  float mult;  // d(v)/d(normalised grid coordinate)
};

// ATen grid_sampler: unnormalise (align_corners=False), reflect about [-0.5, n-0.5], clip to
// [0, n-1] (gradient zero where clipped, clip_coordinates_set_grad uses <= / >=).
// two or more reflections (never for a rotation about the centre, whose coordinates stay within 1.21 n):
// kept out of line so the pixel loops carry no fmodf / division code
__device__ __noinline__ float2 reflect_general(float v, float fn, float mult) {
  const float flips = floorf(v / fn);
  const float extra = fmodf(v, fn);
  if (((int)flips & 1) == 0) return make_float2(extra - 0.5f, mult);
  return make_float2(fn - extra - 0.5f, -mult);
}
__device__ __forceinline__ SrcCoord src_coord(float g, int n) {
  const float fn = (float)n;
  const float u = ((g + 1.f) * fn - 1.f) * 0.5f;
  float mult = fn * 0.5f;
  float v = u + 0.5f;
  mult = v < 0.f ? -mult : mult;
  v = fabsf(v);
  // flips = 0: fmodf(v, fn) = v; flips = 1: fmodf(v, fn) = v - fn, exact (Sterbenz) -- branch-free selects
  const bool flip1 = v >= fn;
  float r = flip1 ? fn - (v - fn) - 0.5f : v - 0.5f;
  mult = flip1 ? -mult : mult;
  if (v >= 2.f * fn) {
    const float2 q = reflect_general(v, fn, flip1 ? -mult : mult);
    r = q.x; mult = q.y;
  }
  if (r <= 0.f) { r = 0.f; mult = 0.f; }
  else if (r >= fn - 1.f) { r = fn - 1.f; mult = 0.f; }
  SrcCoord o; o.v = r; o.mult = mult;
  return o;
}

// Shared-memory tile: (H+1) rows of odd pitch >= W+1 floats, the extra column / row zero-filled.  Source
// coordinates are clipped to [0, n-1], so a +1 tap can only fall outside when its weight is exactly 0 (x0 = W-1,
// fx = 0): with the zero border all four taps are unpredicated loads.  The odd pitch makes bank = (x + y) mod 32,
// so neither near-horizontal nor near-vertical source lines (rotations near 0 / 90 degrees) serialise.
// The kernels are instruction-bound, not gather-bound (measured: the bank layout alone changed nothing), so the
// pixel loop is kept lean: 4 consecutive pixels per thread (one float4 store), no integer divisions, per-row
// terms hoisted, reflection resolved without fmodf for up to one flip.
__device__ __host__ __forceinline__ int tile_pitch(int W) { return (W + 1) | 1; }

__device__ __forceinline__ void stage_image(const float* __restrict__ src, float* __restrict__ tile, int H, int W) {
  const int pitch = tile_pitch(W);
  if ((W & 3) == 0) {
    const int wq = W >> 2;
    int y = threadIdx.x / wq, x4 = threadIdx.x - y * wq;
    const int dy = blockDim.x / wq, dx = blockDim.x - dy * wq;
    // four independent 16-byte loads in flight per thread before the first store (the staging phase is pure latency)
    const int total = H * wq, stride = blockDim.x;
    int i = threadIdx.x;
    for (; i + 3 * stride < total; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(src) + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float* d = tile + y * pitch + 4 * x4;
        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        y += dy; x4 += dx;
        if (x4 >= wq) { x4 -= wq; ++y; }
      }
    }
    for (; i < total; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      float* d = tile + y * pitch + 4 * x4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      y += dy; x4 += dx;
      if (x4 >= wq) { x4 -= wq; ++y; }
    }
  } else {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
      const int y = i / W;
      tile[y * pitch + (i - y * W)] = __ldg(src + i);
    }
  }
  for (int y = threadIdx.x; y < H; y += blockDim.x)
    for (int x = W; x < pitch; ++x) tile[y * pitch + x] = 0.f;
  for (int x = threadIdx.x; x < pitch; x += blockDim.x) tile[H * pitch + x] = 0.f;
}

// (float)k for 0 <= k < 2^23 without the conversion pipe (F2I / I2F / FRND run at a quarter of the FMA rate and
// five of them per pixel made that pipe a co-limiter): k | 0x4B000000 is the float 2^23 + k
__device__ __forceinline__ float small_int_to_float(int k) { return __int_as_float(k | 0x4B000000) - 8388608.f; }

// bilinear taps of one output pixel (ATen grid_sampler_2d, reflection padding, align_corners=False)
struct Taps {
  int o00;                  // tile offset of the (y0, x0) tap; the others are +1, +pitch, +pitch+1
  bool x1ok, y1ok;
  float fx, fy, multx, multy;
};
// gxr = -s*ys, gyr = c*ys: the row terms of the rotated grid coordinate (same products as c*xs - s*ys, s*xs + c*ys)
__device__ __forceinline__ Taps make_taps(float xs, float gxr, float gyr, float c, float s, int H, int W, int pitch) {
  Taps t;
  const SrcCoord cx = src_coord(fmaf(c, xs, gxr), W);
  const SrcCoord cy = src_coord(fmaf(s, xs, gyr), H);
  const int x0 = __float2int_rd(cx.v), y0 = __float2int_rd(cy.v);     // 0 <= v <= n-1 after the clip
  t.fx = cx.v - small_int_to_float(x0); t.fy = cy.v - small_int_to_float(y0);
  t.multx = cx.mult; t.multy = cy.mult;
  t.x1ok = x0 + 1 < W; t.y1ok = y0 + 1 < H;
  t.o00 = y0 * pitch + x0;
  return t;
}
template <bool kPadded>
__device__ __forceinline__ void fetch_taps(const float* __restrict__ tap, const Taps& t, int pitch, float& v00, float& v01,
                                           float& v10, float& v11) {
  v00 = tap[t.o00];
  if (kPadded) {
    v01 = tap[t.o00 + 1]; v10 = tap[t.o00 + pitch]; v11 = tap[t.o00 + pitch + 1];
  } else {
    v01 = t.x1ok ? tap[t.o00 + 1] : 0.f;
    v10 = t.y1ok ? tap[t.o00 + pitch] : 0.f;
    v11 = (t.x1ok && t.y1ok) ? tap[t.o00 + pitch + 1] : 0.f;
  }
}

// Warp block shape: kBW x kBH pixels, lane = 4 consecutive pixels of one row.  16 x 8 rather than 32 x 4: a block that
// touches the image border is never interior (border pixels rotate outside), and with 32-pixel-wide blocks half the
// blocks of a 128-wide image touch the left or right border; 16 x 8 leaves 2 of 8 column blocks and 2 of 16 row blocks.
// A row segment of a warp's store is 64 bytes (two full 32-byte sectors).
static constexpr int kBW = 16, kBH = 128 / kBW, kBLX = kBW / 4;
// ---- interior fast path -------------------------------------------------------------------------------------------
// A rotation about the patch centre maps a warp's block of output pixels onto a rotated rectangle; when its four
// corners land inside [0, n-1] so does every pixel of it (convexity), and then reflect + clip are the identity on
// the coordinate (mult = +n/2) and all four taps are in range.  The per-pixel work shrinks from ~75 to ~40
// instructions: the kernels are issue-bound, not HBM-bound, so this is what moves them towards the copy rate.  The
// arithmetic on the coordinate is the SAME sequence as ATen's ((g+1)*n-1)/2, (v+0.5)-0.5, so results are
// bit-identical to the general path.
__device__ __forceinline__ float unnorm(float g, float fn) {
  const float u = ((g + 1.f) * fn - 1.f) * 0.5f;
  return (u + 0.5f) - 0.5f;               // fabs / reflect of ATen on an in-range coordinate: same two roundings
}
// Warp-cooperative and exact: lanes 0..3 each map one corner of the warp's block through the SAME coordinate
// arithmetic as the pixels (an affine map sends the block onto a parallelogram: corners inside => everything inside),
// one vote decides for the warp.  Must be called by all 32 lanes.  (A test on the distance from the centre is
// cheaper still but can never pass for the two outer 32-pixel column blocks of a 128-wide image -- half the blocks.)
__device__ __forceinline__ bool block_interior(int i0, int j0, float c, float s, int H, int W, float invH, float invW) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  if (lane < 4) {
    const int i = (lane & 2) ? min(i0 + kBH - 1, H - 1) : i0, j = (lane & 1) ? min(j0 + kBW - 1, W - 1) : j0;
    const float ys = (2.f * (float)i + 1.f) * invH - 1.f, xs = (2.f * (float)j + 1.f) * invW - 1.f;
    const float fw = (float)W, fh = (float)H;
    const float ix = unnorm(fmaf(c, xs, -(s * ys)), fw), iy = unnorm(fmaf(s, xs, c * ys), fh);
    ok = ix > 0.01f && ix < fw - 1.01f && iy > 0.01f && iy < fh - 1.01f;      // 1e-2 px guard for rounding
  }
  return __all_sync(0xffffffffu, ok);
}
struct FastTaps { int o00; float fx, fy; };
__device__ __forceinline__ FastTaps make_taps_interior(float xs, float gxr, float gyr, float c, float s, float fh, float fw,
                                                       int pitch) {
  FastTaps t;
  const float ix = unnorm(fmaf(c, xs, gxr), fw), iy = unnorm(fmaf(s, xs, gyr), fh);
  const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);
  t.fx = ix - small_int_to_float(x0); t.fy = iy - small_int_to_float(y0);
  t.o00 = y0 * pitch + x0;
  return t;
}

// Walks the image in kBW x kBH-pixel warp blocks, warps round-robin.
struct BlockWalk {
  int bx, by, bw, bh, nwarps;
  __device__ __forceinline__ BlockWalk(int H, int W) {
    bw = (W + kBW - 1) / kBW; bh = (H + kBH - 1) / kBH; nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    by = warp / bw; bx = warp - by * bw;
  }
  __device__ __forceinline__ bool valid() const { return by < bh; }
  __device__ __forceinline__ void next() {
    bx += nwarps;
    while (bx >= bw) { bx -= bw; ++by; }
  }
};

// 512 threads x 3 CTAs per SM (32 registers, 66 KB of shared memory each): the gathers are dependent shared-memory
// loads, and 8 warps per CTA left each scheduler with two warps in a CTA's compute phase
template <bool kSmem>
__global__ void __launch_bounds__(512, 3) rot_sample_fwd_kernel(
    const float* __restrict__ img, const float* __restrict__ cs, float sgn, int C, int H, int W,
    float* __restrict__ out) {
  extern __shared__ __align__(16) float s_img[];
  const int bc = blockIdx.x;
  const int b = bc / C;
  const float* src = img + (int64_t)bc * H * W;
  float* dst = out + (int64_t)bc * H * W;
  if (kSmem) {
    stage_image(src, s_img, H, W);
    __syncthreads();
  }
  const float* tap = kSmem ? s_img : src;
  const int pitch = kSmem ? tile_pitch(W) : W;
  const float c = cs[2 * b], s = sgn * cs[2 * b + 1];
  const float invW = 1.f / (float)W, invH = 1.f / (float)H;
  const float fw = (float)W, fh = (float)H;
  const int lane = threadIdx.x & 31;
  const int lx = lane % kBLX, ly = lane / kBLX;
  const bool vec_ok = (W & 3) == 0 && (((uintptr_t)dst) & 15) == 0;
  for (BlockWalk w(H, W); w.valid(); w.next()) {
    const int i = w.by * kBH + ly, j0 = w.bx * kBW + lx * 4;
    const bool interior = block_interior(w.by * kBH, w.bx * kBW, c, s, H, W, invH, invW);   // all lanes vote
    if (i >= H || j0 >= W) continue;
    const float ys = (2.f * small_int_to_float(i) + 1.f) * invH - 1.f;
    const float gxr = -(s * ys), gyr = c * ys;
    float o[4];
    if (interior && j0 + 3 < W) {
      const float jf = small_int_to_float(j0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xs = (2.f * (jf + (float)k) + 1.f) * invW - 1.f;
        const FastTaps t = make_taps_interior(xs, gxr, gyr, c, s, fh, fw, pitch);
        const float* q = tap + t.o00;
        const float v00 = q[0], v01 = q[1], v10 = q[pitch], v11 = q[pitch + 1];
        const float ax = 1.f - t.fx, ay = 1.f - t.fy;
        o[k] = v00 * (ax * ay) + v01 * (t.fx * ay) + v10 * (ax * t.fy) + v11 * (t.fx * t.fy);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xs = (2.f * small_int_to_float(j0 + k) + 1.f) * invW - 1.f;
        const Taps t = make_taps(xs, gxr, gyr, c, s, H, W, pitch);
        float v00, v01, v10, v11;
        fetch_taps<kSmem>(tap, t, pitch, v00, v01, v10, v11);
        // same association as ATen: nw*(1-fx)(1-fy) + ne*fx(1-fy) + sw*(1-fx)fy + se*fx*fy
        o[k] = v00 * ((1.f - t.fx) * (1.f - t.fy)) + v01 * (t.fx * (1.f - t.fy)) +
               v10 * ((1.f - t.fx) * t.fy) + v11 * (t.fx * t.fy);
      }
    }
    if (vec_ok) {
      *reinterpret_cast<float4*>(dst + i * W + j0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      for (int k = 0; k < 4 && j0 + k < W; ++k) dst[i * W + j0 + k] = o[k];
    }
  }
}

// Backward, one CTA per sample, ONE shared tile used twice: phase A stages the image and gathers the four taps
// for dL/d(cos), dL/d(sin) (warp shuffles + one shared exchange); phase B (only when grad_input is wanted)
// re-uses the tile as the grad_input accumulator -- shared atomics, no HBM traffic, one coalesced float4 write
// of the tile.  One 66 KB tile instead of image + accumulator side by side keeps 3 CTAs per SM resident.
// kSmem = false (image larger than shared memory): taps from global memory, global atomics on a pre-zeroed gimg.
template <bool kSmem>
__global__ void __launch_bounds__(512, 2) rot_sample_bwd_kernel(
    const float* __restrict__ img, const float* __restrict__ cs, float sgn,
    const float* __restrict__ gout, int C, int H, int W, float* __restrict__ gimg,
    float* __restrict__ gcs) {
  extern __shared__ __align__(16) float s_tile[];
  __shared__ float red[2][32];
  const int b = blockIdx.x;
  const int n = H * W;
  const float c = cs[2 * b], s = sgn * cs[2 * b + 1];
  const float invW = 1.f / (float)W, invH = 1.f / (float)H;
  const float fw = (float)W, fh = (float)H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int lx = lane % kBLX, ly = lane / kBLX;
  const int pitch = kSmem ? tile_pitch(W) : W;
  float acc_c = 0.f, acc_s = 0.f;
  for (int ch = 0; ch < C; ++ch) {
    const float* src = img + ((int64_t)b * C + ch) * n;
    const float* go = gout + ((int64_t)b * C + ch) * n;
    float* gi = gimg ? gimg + ((int64_t)b * C + ch) * n : nullptr;
    const bool vec_ok = (W & 3) == 0 && (((uintptr_t)go) & 15) == 0;
    if (gcs) {
      if (kSmem) {
        __syncthreads();
        stage_image(src, s_tile, H, W);
        __syncthreads();
      }
      const float* tap = kSmem ? s_tile : src;
      for (BlockWalk w(H, W); w.valid(); w.next()) {
        const int i = w.by * kBH + ly, j0 = w.bx * kBW + lx * 4;
        const bool interior = block_interior(w.by * kBH, w.bx * kBW, c, s, H, W, invH, invW);   // all lanes vote
        if (i >= H || j0 >= W) continue;
        const float ys = (2.f * small_int_to_float(i) + 1.f) * invH - 1.f;
        const float gxr = -(s * ys), gyr = c * ys;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec_ok) { const float4 q = __ldg(reinterpret_cast<const float4*>(go + i * W + j0)); g4[0] = q.x; g4[1] = q.y; g4[2] = q.z; g4[3] = q.w; }
        else for (int k = 0; k < 4 && j0 + k < W; ++k) g4[k] = __ldg(go + i * W + j0 + k);
        if (interior && j0 + 3 < W) {
          const float jf = small_int_to_float(j0);
          float bc = 0.f, bs = 0.f;            // multx = W/2, multy = H/2 are applied once per block
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xs = (2.f * (jf + (float)k) + 1.f) * invW - 1.f;
            const FastTaps t = make_taps_interior(xs, gxr, gyr, c, s, fh, fw, pitch);
            const float* q = tap + t.o00;
            const float v00 = q[0], v01 = q[1], v10 = q[pitch], v11 = q[pitch + 1];
            const float g = g4[k];
            const float gix = (-(1.f - t.fy) * v00 + (1.f - t.fy) * v01 - t.fy * v10 + t.fy * v11) * g;
            const float giy = (-(1.f - t.fx) * v00 - t.fx * v01 + (1.f - t.fx) * v10 + t.fx * v11) * g;
            const float ggx = gix * (fw * 0.5f), ggy = giy * (fh * 0.5f);
            bc += ggx * xs + ggy * ys;
            bs += -ggx * ys + ggy * xs;
          }
          acc_c += bc; acc_s += bs;
        } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (j0 + k >= W) break;
          const float xs = (2.f * small_int_to_float(j0 + k) + 1.f) * invW - 1.f;
          const Taps t = make_taps(xs, gxr, gyr, c, s, H, W, pitch);
          float v00, v01, v10, v11;
          fetch_taps<kSmem>(tap, t, pitch, v00, v01, v10, v11);
          const float g = g4[k];
          // d(out)/d(ix), d(out)/d(iy)
          const float gix = (-(1.f - t.fy) * v00 + (1.f - t.fy) * v01 - t.fy * v10 + t.fy * v11) * g;
          const float giy = (-(1.f - t.fx) * v00 - t.fx * v01 + (1.f - t.fx) * v10 + t.fx * v11) * g;
          const float ggx = gix * t.multx, ggy = giy * t.multy;
          // affine_grid backward: base_grid^T @ grad_grid for [[c,-s],[s,c]]
          acc_c += ggx * xs + ggy * ys;
          acc_s += -ggx * ys + ggy * xs;
        }
        }
      }
    }
    if (gi) {
      float* acc = kSmem ? s_tile : gi;
      if (kSmem) {
        __syncthreads();
        for (int k = threadIdx.x; k < (H + 1) * pitch; k += blockDim.x) s_tile[k] = 0.f;
        __syncthreads();
      }
      for (BlockWalk w(H, W); w.valid(); w.next()) {
        const int i = w.by * kBH + ly, j0 = w.bx * kBW + lx * 4;
        const bool interior = block_interior(w.by * kBH, w.bx * kBW, c, s, H, W, invH, invW);   // all lanes vote
        if (i >= H || j0 >= W) continue;
        const float ys = (2.f * small_int_to_float(i) + 1.f) * invH - 1.f;
        const float gxr = -(s * ys), gyr = c * ys;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec_ok) { const float4 q = __ldg(reinterpret_cast<const float4*>(go + i * W + j0)); g4[0] = q.x; g4[1] = q.y; g4[2] = q.z; g4[3] = q.w; }
        else for (int k = 0; k < 4 && j0 + k < W; ++k) g4[k] = __ldg(go + i * W + j0 + k);
        if (interior && j0 + 3 < W) {
          const float jf = small_int_to_float(j0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xs = (2.f * (jf + (float)k) + 1.f) * invW - 1.f;
            const FastTaps t = make_taps_interior(xs, gxr, gyr, c, s, fh, fw, pitch);
            const float g = g4[k];
            float* q = acc + t.o00;
            const float ax = 1.f - t.fx, ay = 1.f - t.fy;
            atomicAdd(q, ax * ay * g);
            atomicAdd(q + 1, t.fx * ay * g);
            atomicAdd(q + pitch, ax * t.fy * g);
            atomicAdd(q + pitch + 1, t.fx * t.fy * g);
          }
        } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (j0 + k >= W) break;
          const float xs = (2.f * small_int_to_float(j0 + k) + 1.f) * invW - 1.f;
          const Taps t = make_taps(xs, gxr, gyr, c, s, H, W, pitch);
          const float g = g4[k];
          atomicAdd(acc + t.o00, (1.f - t.fx) * (1.f - t.fy) * g);
          if (t.x1ok) atomicAdd(acc + t.o00 + 1, t.fx * (1.f - t.fy) * g);
          if (t.y1ok) atomicAdd(acc + t.o00 + pitch, (1.f - t.fx) * t.fy * g);
          if (t.x1ok && t.y1ok) atomicAdd(acc + t.o00 + pitch + 1, t.fx * t.fy * g);
        }
        }
      }
      if (kSmem) {
        __syncthreads();
        if ((W & 3) == 0 && (((uintptr_t)gi) & 15) == 0) {
          const int wq = W >> 2;
          int y = threadIdx.x / wq, x4 = threadIdx.x - y * wq;
          const int dy = blockDim.x / wq, dx = blockDim.x - dy * wq;
          for (int k = threadIdx.x; k < H * wq; k += blockDim.x) {
            const float* sp = s_tile + y * pitch + 4 * x4;
            reinterpret_cast<float4*>(gi)[k] = make_float4(sp[0], sp[1], sp[2], sp[3]);
            y += dy; x4 += dx;
            if (x4 >= wq) { x4 -= wq; ++y; }
          }
        } else {
          for (int k = threadIdx.x; k < n; k += blockDim.x) { const int y = k / W; gi[k] = s_tile[y * pitch + (k - y * W)]; }
        }
      }
    }
  }
  if (!gcs) return;
  acc_c = warp_sum(acc_c);
  acc_s = warp_sum(acc_s);
  if (lane == 0) { red[0][warp] = acc_c; red[1][warp] = acc_s; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < nwarps ? red[0][lane] : 0.f;
    float d = lane < nwarps ? red[1][lane] : 0.f;
    a = warp_sum(a);
    d = warp_sum(d);
    if (lane == 0) { gcs[2 * b] = a; gcs[2 * b + 1] = sgn * d; }
  }
}

static constexpr int kMaxSmemImage = 200 * 1024;  // bytes of one staged image / gradient tile

}  // namespace livae

static size_t tile_bytes(int H, int W) { return (size_t)(H + 1) * livae::tile_pitch(W) * sizeof(float); }

extern "C" int livae_rot_sample_fwd(const float* img, const float* cs, float sgn, int B, int C, int H,
                                    int W, float* out, livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0, "rot_sample_fwd: bad sizes");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && cs && out, "rot_sample_fwd: null pointer");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = tile_bytes(H, W);
  if (bytes <= (size_t)kMaxSmemImage && ((uintptr_t)img & 15) == 0) {
    static OncePerDevice attr_done;
    if (attr_done.first()) {
      cudaFuncSetAttribute(rot_sample_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kMaxSmemImage);
    }
    rot_sample_fwd_kernel<true><<<B * C, 512, bytes, st>>>(img, cs, sgn, C, H, W, out);
  } else {
    rot_sample_fwd_kernel<false><<<B * C, 512, 0, st>>>(img, cs, sgn, C, H, W, out);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int livae_rot_sample_bwd(const float* img, const float* cs, float sgn, const float* gout,
                                    int B, int C, int H, int W, float* gimg, float* gcs,
                                    livae_stream_t stream) {
  using namespace livae;
  LIVAE_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0, "rot_sample_bwd: bad sizes");
  if (B == 0) return 0;
  LIVAE_CHECK_ARG(img && cs && gout, "rot_sample_bwd: null pointer");
  LIVAE_CHECK_ARG(gimg || gcs, "rot_sample_bwd: nothing to compute");
  if (int e = require_sm100()) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = tile_bytes(H, W);
  static OncePerDevice attr_done;
  if (attr_done.first()) {
    cudaFuncSetAttribute(rot_sample_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemImage);
  }
  if (bytes <= (size_t)kMaxSmemImage && ((uintptr_t)img & 15) == 0) {
    rot_sample_bwd_kernel<true><<<B, 512, bytes, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  } else {
    if (gimg) {
      cudaError_t e = cudaMemsetAsync(gimg, 0, (size_t)B * C * H * W * sizeof(float), st);
      if (e != cudaSuccess) { set_error("rot_sample_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    rot_sample_bwd_kernel<false><<<B, 512, 0, st>>>(img, cs, sgn, gout, C, H, W, gimg, gcs);
  }
  LIVAE_CUDA_LAUNCH_CHECK();
  return 0;
}

// Blackwell (sm_100a) primitives used by the tensor-core convolution engine: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and UMMA descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace livae {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(a), "r"(parity)
      : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tcgen05.mma, TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Byte offset of 16-byte chunk `chunk` of row `row` inside a K-major tile whose rows are `rb` = 32 / 64 /
// 128 bytes and hardware-swizzled (CU_TENSOR_MAP_SWIZZLE_32B/64B/128B: address bits [4, 4+n) ^= bits
// [7, 7+n)); the tile base must be 1024-byte aligned.  Threads that build an operand tile themselves
// write through this map so the tile is byte-identical to what TMA would have delivered.
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t chunk, uint32_t rb) {
  const uint32_t off = row * rb + chunk * 16u;
  return off ^ (((off >> 7) & ((rb >> 4) - 1u)) << 4);
}

// One lane of a converged warp.  MMA-issuing warps run their loops as a WHOLE warp and guard only the
// tcgen05.mma / tcgen05.commit with this predicate: in a warp-uniform loop the descriptor arithmetic stays on
// the uniform datapath, and one warp then issues an N <= 64 MMA every 39-48 cycles (the shared-memory operand
// bound); the same loop under `if (lane == 0)` is divergent code and manages one per ~68 (tools/umma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16/fp16 operands, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
      "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets row (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor.
//   bits  0-13 start address >> 4      bits 16-29 leading-dim byte offset >> 4
//   bits 32-45 stride-dim byte offset >> 4   bits 46-47 version (1 on sm_100)
//   bits 61-63 swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
//   bits 49-51 base offset: (start address >> 7) & 7 when the start is not aligned to the swizzle
//   repeat (1024 B for 128B swizzle), e.g. a tile addressed from a shifted row
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16: fp32 accumulator, bf16 A and B.
//   bits 4-5 D format (1 = f32); 7-9 A format (1 = bf16); 10-12 B format; bit 15 A major (1 = MN);
//   bit 16 B major; bits 17-22 N >> 3; bits 24-28 M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same with fp16 A and B (format code 0): 11-bit mantissa for operands of known small range
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

inline uint32_t swizzle_layout_type(int row_bytes) {  // UMMA layout_type for a K-major row pitch
  return row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : row_bytes == 32 ? 6u : 0u;
}

// ---- pipeline tracing ------------------------------------------------------------------------
// livae_set_probe(dev_ptr) hands the kernels a device buffer of int64; CTA 0 of a traced kernel appends
// (role << 60 | slot << 56 | clock64) records at its pipeline points (role-private regions of 1024
// records), which tools/probe.py prints as a timeline.  nullptr (default) = no tracing.
extern long long* g_probe;
__device__ __forceinline__ void probe_rec(long long* buf, int role, int slot, int& n) {
  if (buf && blockIdx.x == 0 && blockIdx.y == 0 && n < 1024)
    buf[role * 1024 + n++] = ((long long)slot << 56) | (clock64() & 0xffffffffffffffll);
}

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// bf16 tensor map: dims/strides innermost first; strides in BYTES for dims 1..rank-1
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, int inner_bytes);

// ---- halo-kernel launchers shared between translation units (conv_tc.cu, wgrad_tc.cu, conv_s2d.cu) ----
// a_s2d: the input operand is read in SPACE-TO-DEPTH form: `in` is a plain NHWC tensor [B, 2*Hin, 2*Win, Cin/4]
//        and a strided tensor map delivers each 2x2 pixel block as 4*C channels (order (ey, ex, c)), one K
//        chunk per pixel row of the block -- no data movement, the re-blocking is done by the TMA strides.
// epi_mode 0: plain NHWC store; 1: N = 4*Co columns are the four pixels of a 2x2 block -> bias + ReLU +
//        max-pool over them, pooled bf16 [B,Hq,Wq,Co] + argmax (pool_idx); 2: N = 4*Ci columns are written
//        back as the four pixels of the block of a plain NHWC tensor [B,2*Hq,2*Wq,Ci] (relu_mask in that layout)
//        3: N = 4*up_cout columns = the four output pixels of the phase-folded up-sampling convolution: bias + ReLU
//        (none on the two outermost rows / columns, finished by upfold.cu), written to [B,2*Hq,2*Wq,up_cout]
struct HaloOpts { int a_s2d; int epi_mode; uint8_t* pool_idx; int up_cout; };
int launch_conv_tc_halo(const void* in, int B, int Hin, int Win, int Cin, const void* wpacked, int wtaps, int N,
                        int Hq, int Wq, int Ho, int Wo, int os, int oy0, int ox0, int in_stride, int ntaps,
                        const int* tdy, const int* tdx, const int* tw_idx, void* out, int out_f32,
                        const float* bias, int act, const void* relu_mask, cudaStream_t st, HaloOpts opts);
int halo_box_width_public(int Hq, int Wq, int max_sx);   // conv_tc.cu: tile box width of the halo kernel (host arithmetic)
// returns 0 = launched, 1 = shape not eligible
int launch_wgrad_halo(const livae_tc_conv_desc* d, const void* x, const void* gy, float* gw_acc, int Ho, int Wo,
                      cudaStream_t st, int x_s2d);   // x_s2d: bit 0 = x, bit 1 = gy read block-wise (2x2 blocks as channels)
void colsum_bf16(const void* g, int64_t R, int C, float* gb, cudaStream_t st);   // gb must be zeroed
void sum_slices(const float* part, int nslices, int64_t n, float* out, cudaStream_t st);

}  // namespace tc
}  // namespace livae

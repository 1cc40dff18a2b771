// Blackwell (sm_100a) primitives used by the tensor-core chain kernels: mbarrier, bulk async copy (TMA engine,
// UBLKCP), tcgen05 TMEM allocation / MMA / commit / load, and the SWIZZLE_128B K-major operand image.
//
// Operand image.  A tf32 operand tile of R rows x 32 columns (one 128-byte swizzle row per matrix row) is stored
// as R/8 groups of 1024 bytes; inside a group row r occupies bytes [128 r, 128 r + 128) and its eight 16-byte
// chunks are XOR-permuted by (r % 8) - the canonical K-major SWIZZLE_128B layout of the UMMA shared-memory
// descriptor (stride-byte-offset 1024).  Weight tiles are written in exactly this image by the packing kernel, so
// they reach shared memory with one linear cp.async.bulk each and need no tensor map.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vdn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, k) in a SWIZZLE_128B K-major tile with 32 fp32 columns
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t k) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 2) ^ r) & 7u) << 4) + ((k & 3u) << 2);
}

// byte offset of element (r, k) in a SWIZZLE_128B K-major tile with 64 16-bit columns
__host__ __device__ __forceinline__ uint32_t sw128_offset_h(uint32_t r, uint32_t k) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 3) ^ r) & 7u) << 4) + ((k & 7u) << 1);
}

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: returns false (and the caller bails out) instead of hanging the GPU if a barrier never flips.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, uint32_t max_spins = 1u << 24) {
#pragma unroll 1
  for (uint32_t i = 0; i < max_spins; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
  }
  return false;
}
// For the epilogue warps of the chain engine: a few tight polls (the barrier usually flips within them), then back off, so
// that warps waiting for an accumulator do not take issue slots from the warps that are still producing it - a spinning
// warp is always "ready" to the scheduler, and twelve of them made the four warps of a ragged skip layer four times slower.
__device__ __forceinline__ bool mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t max_spins = 1u << 22) {
#pragma unroll 1
  for (uint32_t i = 0; i < max_spins; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if (i >= 8) __nanosleep(40);
  }
  return false;
}
// Same, for long waits of single-thread roles (MMA issuer, weight stream): back off between polls so the spinning
// thread does not take issue slots from the warps doing the element-wise work on its scheduler.
__device__ __forceinline__ bool mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns = 64, uint32_t max_spins = 1u << 22) {
#pragma unroll 1
  for (uint32_t i = 0; i < max_spins; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(ns);
  }
  return false;
}

// ---- async proxy ------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// linear global -> shared bulk copy completing on an mbarrier (bytes must be a multiple of 16)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- tensor memory ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t result_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- UMMA descriptors -------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_units = 1) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(lbo_units & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout (CUTLASS: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): rows of the image are K indices, the 128-byte row holds 32
// consecutive M/N indices, rows come in groups of 4 (sbo_bytes apart) and inside a group the four 32-byte blocks
// of row r are XOR-permuted by (r % 4); further 32-index chunks of M/N are lbo_bytes apart.
__host__ __device__ __forceinline__ uint32_t sw128_32b_offset(uint32_t r, uint32_t c) {   // row r (K), column c (<32)
  return (r >> 2) * 512u + (r & 3u) * 128u + ((((c >> 3) ^ r) & 3u) << 5) + ((c & 7u) << 2);
}
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor for kind::tf32 with both operands MN-major (the weight-gradient GEMM)
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// instruction descriptor for kind::f16 with fp16 operands, fp32 accumulate, both operands K-major
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T with fp16 operands: A in tensor memory holds two consecutive K elements per 32-bit
// column (low half = even k), one MMA consumes K = 16 (8 columns)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (lane = row, one 32-bit column per K
// element), which takes its 4 KB per instruction off the shared-memory port
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: lane i of the warp writes 32 consecutive columns of TMEM lane (base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 8-column variants (lane i of the warp <-> TMEM lane base + i, 8 consecutive columns)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// four packed 32-bit words (e.g. 8 fp16 values) per lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 consecutive columns, lane i of the warp reads TMEM lane (base + i) ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace vdn

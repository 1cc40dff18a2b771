// tcgen05 / TMEM GEMMs (tensor-core "tf32" mode of the MLP path), same operand-prologue / epilogue contract as
// the FFMA kernels of gemm_simt.cuh:
//
//   gemm_nt_tc : C[M,N] = epi( pro(A)[M,K] * W[N,K]^T + bias )
//
// One CTA owns a 128-row tile and up to 256 output columns; two CTAs are co-resident per SM so one tile's
// epilogue overlaps the other's MMAs.  Warp roles: warps 0-7 bring the A operand in with cp.async (16 bytes per
// thread and instruction, straight into the SWIZZLE_128B shared-memory image, zero fill outside the operand) and,
// one K block later, apply the fused prologue to their own chunks in place and round them to tf32; thread 0 streams
// the pre-swizzled weight tile of each stage with cp.async.bulk (TMA engine); one thread of warp 8 issues
// tcgen05.mma kind::tf32 into a TMEM accumulator; then the eight warps drain TMEM with tcgen05.ld, transpose
// through shared memory and run the epilogue with coalesced global accesses (its auxiliary operands are pulled
// into L2 a few K blocks ahead).  A 2-stage mbarrier ring (full / empty) connects the roles; every wait is bounded
// and raises a device fault flag instead of hanging.  The prologue / epilogue kind is switched once per tile-row
// group, outside the per-element loops.  History that shaped this: with the kind switch inlined per element the
// kernel was instruction-fetch bound; with register-staged operand loads (address arithmetic, predicates and
// conversion per element, ~300 instructions per thread and K block) it was issue bound in the producers - the
// cp.async form needs ~60.  An output width of 257 (a 256-wide head stacked with a scalar one) is served by the
// producers accumulating the extra column as an fp32 dot product instead of a second N tile.
#pragma once
#include "gemm_simt.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace vdn {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_THREADS = 288;   // warps 0-7: operand producers, then epilogue; warp 8: TMEM allocation + MMA issue

// Round to tf32 (nearest, ties away from zero - what cvt.rna.tf32.f32 does) in two integer instructions instead of
// the three (+ predicate) the conversion compiles to; differs from it only for Inf / NaN inputs, which the MLP
// operands never hold.
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float4 f4_map_sp(float4 a) {
  return make_float4(softplus100_fast(a.x), softplus100_fast(a.y), softplus100_fast(a.z), softplus100_fast(a.w));
}
__device__ __forceinline__ float4 f4_map_sp1(float4 a) {
  return make_float4(softplus100_d1_fast(a.x), softplus100_d1_fast(a.y), softplus100_d1_fast(a.z),
                     softplus100_d1_fast(a.w));
}
__device__ __forceinline__ float4 f4_map_sp2(float4 a) {
  return make_float4(softplus100_d2_fast(a.x), softplus100_d2_fast(a.y), softplus100_d2_fast(a.z),
                     softplus100_d2_fast(a.w));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// Prologue of 4 float4 groups (4 rows, same 4 columns): kind switched once.
__device__ __forceinline__ void tc_prologue4(const Operand& A, const RawLoad (&raw)[4], float4 (&v)[4]) {
  switch (A.kind) {
    case PRO_SOFTPLUS:
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = f4_map_sp(raw[i].a);
      break;
    case PRO_DSIG:
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = f4_scale(f4_mul(f4_map_sp1(raw[i].b), raw[i].a), A.scale);
      break;
    case PRO_DSIGMOID:
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = raw[i].b;
        v[i] = f4_mul(raw[i].a, make_float4(b.x * (1.f - b.x), b.y * (1.f - b.y), b.z * (1.f - b.z), b.w * (1.f - b.w)));
      }
      break;
    case PRO_RELUMASK:
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = raw[i].a, b = raw[i].b;
        v[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
      }
      break;
    default:
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = raw[i].a;
      break;
  }
}

// Element-wise path of the epilogue for 8 rows x 4 columns (ragged column ranges, unaligned or splitting epilogues):
// same access pattern as the vector path (a warp instruction still covers four 128-byte row segments), scalar
// loads / stores.  The kind switch sits outside the loops.
__device__ __forceinline__ void tc_epilogue8_scalar(const Epilogue& e, int mbase, int M, int n, const float4 (&x)[8], int N) {
  const bool rc = e.round_c != 0;
  auto rnd = [rc](float v) { return rc ? to_tf32(v) : v; };
  float bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) bb[j] = e.bias[n + j];
  }
#define VDN_EPI_LOOP(BODY)                                                   \
  _Pragma("unroll") for (int i = 0; i < 8; ++i) {                            \
    const int m = mbase + 4 * i;                                             \
    if (m < M) {                                                             \
      const size_t mm = (size_t)m;                                           \
      const float xv[4] = {x[i].x + bb[0], x[i].y + bb[1], x[i].z + bb[2], x[i].w + bb[3]}; \
      _Pragma("unroll") for (int j = 0; j < 4; ++j) {                        \
        const int nn = n + j;                                                \
        const float v = xv[j];                                               \
        if (nn < N) { BODY }                                                 \
      }                                                                      \
    }                                                                        \
  }
  switch (e.kind) {
    case EPI_STORE: VDN_EPI_LOOP(e.c[mm * e.ldc + e.coff + nn] = v;) break;
    case EPI_RELU: VDN_EPI_LOOP(e.c[mm * e.ldc + e.coff + nn] = rnd(fmaxf(v, 0.0f));) break;
    case EPI_SIGMOID: VDN_EPI_LOOP(e.c[mm * e.ldc + e.coff + nn] = sigmoidf_(v);) break;
    case EPI_SOFTPLUS: VDN_EPI_LOOP(e.c[mm * e.ldc + e.coff + nn] = softplus100(v);) break;
    case EPI_SDF_SKIP:
      VDN_EPI_LOOP(if (e.c) e.c[mm * e.ldc + nn] = v; e.c2[mm * e.ldc2 + nn] = softplus100(v) * e.scale;) break;
    case EPI_SPLIT:
      VDN_EPI_LOOP(if (nn < e.split) { if (e.c2) e.c2[mm * e.ldc2 + nn] = v * e.scale; }
                   else if (e.c) e.c[mm * e.ldc + e.coff + (nn - e.split)] = v;) break;
    case EPI_ADD_SCALED: VDN_EPI_LOOP(e.c[mm * e.ldc + nn] = v + e.scale * e.aux[mm * e.ldaux + e.split + nn];) break;
    case EPI_GRAD_DUAL:
      VDN_EPI_LOOP(const float z = e.aux[mm * e.ldaux + nn]; const float gin = e.aux2[mm * e.ldaux2 + nn] * e.scale2;
                   e.c[mm * e.ldc + nn] = rnd(softplus100_d1(z) * v * e.scale);
                   e.c2[mm * e.ldc2 + nn] = softplus100_d2(z) * gin * v;) break;
    case EPI_BWD_INJECT:
      VDN_EPI_LOOP(const float z = e.aux[mm * e.ldaux + nn]; float r = softplus100_d1(z) * v * e.scale;
                   if (e.aux2) r += e.aux2[mm * e.ldaux2 + nn]; e.c[mm * e.ldc + nn] = rnd(r);) break;
    case EPI_RELU_MASK: VDN_EPI_LOOP(e.c[mm * e.ldc + nn] = rnd(e.aux[mm * e.ldaux + e.split + nn] > 0.0f ? v : 0.0f);) break;
    default: break;
  }
#undef VDN_EPI_LOOP
}

// Pull the lines of the epilogue's auxiliary operands (saved pre-activations, masks, injected cotangents) that this
// thread will read for column chunk `n` into L2 well before the accumulator is complete: their DRAM latency would
// otherwise be exposed once per chunk.
__device__ __forceinline__ void tc_epilogue_prefetch(const Epilogue& e, int mbase, int M, int n, int N) {
  if (n >= N) return;
  const float* p1 = nullptr;
  const float* p2 = nullptr;
  int l1 = 0, l2 = 0;
  switch (e.kind) {
    case EPI_ADD_SCALED: case EPI_RELU_MASK: p1 = e.aux + e.split + n; l1 = e.ldaux; break;
    case EPI_GRAD_DUAL: p1 = e.aux + n; l1 = e.ldaux; if (e.ldaux2) { p2 = e.aux2 + n; l2 = e.ldaux2; } break;
    case EPI_BWD_INJECT: p1 = e.aux + n; l1 = e.ldaux; if (e.aux2) { p2 = e.aux2 + n; l2 = e.ldaux2; } break;
    default: return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = mbase + 4 * i;
    if (m < M) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + (size_t)m * l1));
      if (p2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p2 + (size_t)m * l2));
    }
  }
}

// Vector epilogue for 8 rows x 4 columns (same columns for all rows); kind switched once.
__device__ __forceinline__ void tc_epilogue8(const Epilogue& e, int mbase, int M, int n, float4 (&x)[8], float4 b) {
  // row of group i is mbase + 4 i
#define m_(i) (mbase + 4 * (i))
#define mok_(i) (mbase + 4 * (i) < M)
  auto st4 = [](float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; };
  const bool rc = e.round_c != 0;
  auto st4c = [rc](float* p, float4 v) {   // store to c, optionally rounded to tf32
    if (rc) v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(p) = v;
  };
  auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
  if (e.bias) {   // b = bias[n .. n+3], fetched by the caller ahead of time
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = f4_add(x[i], b);
  }
  switch (e.kind) {
    case EPI_STORE:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) st4(e.c + (size_t)m_(i) * e.ldc + e.coff + n, x[i]);
      break;
    case EPI_RELU:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i))
          st4c(e.c + (size_t)m_(i) * e.ldc + e.coff + n,
               make_float4(fmaxf(x[i].x, 0.f), fmaxf(x[i].y, 0.f), fmaxf(x[i].z, 0.f), fmaxf(x[i].w, 0.f)));
      break;
    case EPI_SOFTPLUS:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) st4(e.c + (size_t)m_(i) * e.ldc + e.coff + n, f4_map_sp(x[i]));
      break;
    case EPI_SIGMOID:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i))
          st4(e.c + (size_t)m_(i) * e.ldc + e.coff + n,
              make_float4(sigmoidf_(x[i].x), sigmoidf_(x[i].y), sigmoidf_(x[i].z), sigmoidf_(x[i].w)));
      break;
    case EPI_SDF_SKIP:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) {
          if (e.c) st4(e.c + (size_t)m_(i) * e.ldc + n, x[i]);
          st4(e.c2 + (size_t)m_(i) * e.ldc2 + n, f4_scale(f4_map_sp(x[i]), e.scale));
        }
      break;
    case EPI_GRAD_DUAL:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) {
          const float4 z = ld4(e.aux + (size_t)m_(i) * e.ldaux + n);
          const float4 g = f4_scale(ld4(e.aux2 + (size_t)m_(i) * e.ldaux2 + n), e.scale2);
          st4c(e.c + (size_t)m_(i) * e.ldc + n, f4_scale(f4_mul(f4_map_sp1(z), x[i]), e.scale));
          st4(e.c2 + (size_t)m_(i) * e.ldc2 + n, f4_mul(f4_mul(f4_map_sp2(z), g), x[i]));
        }
      break;
    case EPI_BWD_INJECT:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) {
          const float4 z = ld4(e.aux + (size_t)m_(i) * e.ldaux + n);
          float4 r = f4_scale(f4_mul(f4_map_sp1(z), x[i]), e.scale);
          if (e.aux2) r = f4_add(r, ld4(e.aux2 + (size_t)m_(i) * e.ldaux2 + n));
          st4c(e.c + (size_t)m_(i) * e.ldc + n, r);
        }
      break;
    case EPI_RELU_MASK:
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (mok_(i)) {
          const float4 h = ld4(e.aux + (size_t)m_(i) * e.ldaux + e.split + n);
          st4c(e.c + (size_t)m_(i) * e.ldc + n, make_float4(h.x > 0.f ? x[i].x : 0.f, h.y > 0.f ? x[i].y : 0.f,
                                                          h.z > 0.f ? x[i].z : 0.f, h.w > 0.f ? x[i].w : 0.f));
        }
      break;
    default: break;
  }
#undef m_
#undef mok_
}

inline bool epilogue_vec_ok(const Epilogue& e) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (e.bias && !al(e.bias)) return false;
  switch (e.kind) {
    case EPI_STORE: case EPI_RELU: case EPI_SOFTPLUS: case EPI_SIGMOID:
      return al(e.c) && !(e.ldc & 3) && !(e.coff & 3);
    case EPI_SDF_SKIP:
      return (!e.c || (al(e.c) && !(e.ldc & 3))) && al(e.c2) && !(e.ldc2 & 3);
    case EPI_GRAD_DUAL:
      return al(e.c) && al(e.c2) && al(e.aux) && al(e.aux2) && !((e.ldc | e.ldc2 | e.ldaux | e.ldaux2) & 3);
    case EPI_BWD_INJECT:
      return al(e.c) && al(e.aux) && (!e.aux2 || al(e.aux2)) && !((e.ldc | e.ldaux | e.ldaux2) & 3);
    case EPI_RELU_MASK:
      return al(e.c) && al(e.aux) && !((e.ldc | e.ldaux | e.split) & 3);
    default: return false;
  }
}

static __global__ void __launch_bounds__(TC_THREADS, 2)
gemm_nt_tc_kernel(int M, int N, int nkb, Operand A, const float* __restrict__ Bimg, int img_rows, int row0,
                  Epilogue E, int vec_ok, const float* __restrict__ wextra, int* __restrict__ fault,
                  long long* __restrict__ dbg) {
  using namespace tc;
  // optional timeline of CTA 0 (debug): dbg[role*64 + event] = clock64()
  const bool rec = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define VDN_TL(role, ev) do { if (rec) dbg[(role) * 64 + (ev)] = clock64(); } while (0)
  if (threadIdx.x == 0) VDN_TL(0, 0);
  if (dbg && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1500) {   // per-CTA start / SM id (debug)
    unsigned long long g; unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    dbg[1024 + 4 * blockIdx.x] = (long long)g; dbg[1024 + 4 * blockIdx.x + 2] = sm;
  }
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int n_base = blockIdx.y * 256;
  const int n_cta = min(256, N - n_base);
  const int n_mma = (n_cta + 15) & ~15;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)n_mma) ncols <<= 1;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_stride = (16384u + (uint32_t)n_mma * 128u + 1023u) & ~1023u;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 8 + 1);     // one arrival per producer warp + the expect_tx arrival of the weight tile
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;
  if (threadIdx.x == 0) VDN_TL(0, 1);

  if (warp < 8) {
    // ---- A producers: warp w owns rows [16w, 16w+16); per instruction the 32 lanes cover 4 rows x 8 chunks of
    // 16 bytes, i.e. four full 128-byte row segments (coalesced global loads, conflict-free swizzled stores).
    // Global loads run two K blocks ahead of their use (double-buffered registers). ----
    const int chunk = lane & 7;
    const int rbase = warp * 16 + (lane >> 3);                    // row of group i: rbase + 4 i
    const uint32_t r7e = (uint32_t)(lane >> 3), r7o = r7e + 4u;   // (row & 7) for even / odd i
    const uint32_t soff_e = (uint32_t)(warp * 2) * 1024u + r7e * 128u + (((uint32_t)chunk ^ r7e) << 4);
    const uint32_t soff_o = (uint32_t)(warp * 2) * 1024u + r7o * 128u + (((uint32_t)chunk ^ r7o) << 4);
    const bool two = A.kind >= PRO_DSIG;
    const float* pa0 = A.p + (size_t)(m0 + rbase) * A.ld + chunk * 4;
    const float* pb0 = two ? A.p2 + (size_t)(m0 + rbase) * A.ld2 + chunk * 4 : nullptr;
    const size_t stepa = (size_t)4 * A.ld, stepb = (size_t)4 * A.ld2;
    const int mlim = M - m0 - rbase;                              // group i is a valid row iff 4 i < mlim
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pf_kb = nkb > 4 ? nkb - 4 : 0;
    // optional output column N (one past the MMA tile, e.g. the 257th output of the stacked sdf|feature and
    // alpha|feature heads): a plain fp32 dot product accumulated by the producers from the operand values they hold,
    // instead of a second N tile that would re-read the whole operand for a single column
    float ext[4] = {0.f, 0.f, 0.f, 0.f};
    // The rows go global -> shared memory with cp.async, 16 bytes per thread and instruction straight into their
    // swizzled positions (zero fill outside the operand): no staging registers, no per-element address arithmetic.
    // The second operand of the two-operand prologues travels through registers (one K block in flight).  One K block
    // later the same thread applies the prologue to its own four chunks in place, rounds them to tf32, clears what
    // lies outside the logical extent, fences the stores for the tensor core and signals the stage.
    float4 b0[4], b1[4];
    auto issue = [&](int kb, float4 (&bq)[4]) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      if (tid == 0) VDN_TL(1, 3 * kb);
      ok = mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      if (tid == 0) VDN_TL(1, 3 * kb + 1);
      const uint32_t base = smem0 + (uint32_t)s * stage_stride;
      if (tid == 0) {                                // the stage is free: stream its weight tile (TMA engine)
        const uint32_t bytes = (uint32_t)n_mma * 128u;
        mbar_arrive_expect_tx(smem_u32(&bar_full[s]), bytes);
        bulk_g2s(base + 16384u, Bimg + ((size_t)kb * img_rows + row0 + n_base) * 32, bytes, smem_u32(&bar_full[s]));
      }
      const bool cok = kb * 32 + chunk * 4 < A.width;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool p = cok && (4 * i < mlim);
        const float* src = p ? pa0 + i * stepa + kb * 32 : A.p;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + ((i & 1) ? soff_o : soff_e) +
                                                                            (uint32_t)(i >> 1) * 1024u),
                     "l"(src), "r"(p ? 16 : 0)
                     : "memory");
        if (two) bq[i] = p ? *reinterpret_cast<const float4*>(pb0 + i * stepb + kb * 32) : zero4;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (kb == pf_kb) {                             // epilogue operands of this thread -> L2, a few K blocks ahead
        const int nch_e = (n_cta + 31) >> 5;
        for (int ch = warp >> 2; ch < nch_e; ch += 2)
          tc_epilogue_prefetch(E, m0 + (warp & 3) * 32 + (lane >> 3), M, n_base + ch * 32 + chunk * 4, N);
      }
    };
    const bool plain = A.kind == PRO_NONE && A.rounded && !wextra;
    auto finish = [&](int kb, const float4 (&bq)[4]) {
      const int s = kb & 1;
      const int col = kb * 32 + chunk * 4;
      const uint32_t base = smem0 + (uint32_t)s * stage_stride;
      if (plain && kb * 32 + 31 < A.kvalid) {        // tf32-representable already, nothing to clear: publish as it landed
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
        if (tid == 0) VDN_TL(1, 3 * kb + 2);
        return;
      }
      RawLoad raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t addr = base + ((i & 1) ? soff_o : soff_e) + (uint32_t)(i >> 1) * 1024u;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(raw[i].a.x), "=f"(raw[i].a.y), "=f"(raw[i].a.z), "=f"(raw[i].a.w)
                     : "r"(addr));
        raw[i].b = bq[i];
      }
      float4 v[4];
      tc_prologue4(A, raw, v);
      if (A.kind == PRO_SOFTPLUS) {                  // the only prologue that does not map 0 to 0: re-clear the padding
        const bool cok = col < A.width;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (!(cok && 4 * i < mlim)) v[i] = zero4;
      }
      if (col + 3 >= A.kvalid) {                     // ragged K edge: zero the columns beyond the logical extent
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (col + 0 >= A.kvalid) v[i].x = 0.f;
          if (col + 1 >= A.kvalid) v[i].y = 0.f;
          if (col + 2 >= A.kvalid) v[i].z = 0.f;
          if (col + 3 >= A.kvalid) v[i].w = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
      if (wextra) {
        const float4 w4 = col < A.width ? __ldg(reinterpret_cast<const float4*>(wextra + col)) : zero4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ext[i] = fmaf(v[i].x, w4.x, fmaf(v[i].y, w4.y, fmaf(v[i].z, w4.z, fmaf(v[i].w, w4.w, ext[i]))));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((i & 1) ? soff_o : soff_e) +
                                                                     (uint32_t)(i >> 1) * 1024u),
                     "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w)
                     : "memory");
      fence_proxy_async();                           // own stores -> visible to the tensor core's (async proxy) reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
      if (tid == 0) VDN_TL(1, 3 * kb + 2);
    };
    for (int kb = 0; kb < nkb && ok; kb += 2) {
      issue(kb, b0);
      if (kb > 0) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        finish(kb - 1, b1);
      }
      if (kb + 1 < nkb && ok) {
        issue(kb + 1, b1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        finish(kb, b0);
      }
    }
    if (ok && nkb > 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (nkb & 1) finish(nkb - 1, b0); else finish(nkb - 1, b1);
    }
    if (wextra) {                                   // reduce over the 8 lanes that share a row, lane 0 of them stores
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float e = ext[i];
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        e += __shfl_xor_sync(0xffffffffu, e, 4);
        if (chunk == 0 && 4 * i < mlim) epi_store(E, m0 + rbase + 4 * i, N, e);
      }
    }
  } else if (tid == 8 * 32) {
    // ---- MMA issuer ------------------------------------------------------------------------------
    const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)n_mma);
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      VDN_TL(2, 2 * kb);
      ok = mbar_wait(smem_u32(&bar_full[s]), ph);
      VDN_TL(2, 2 * kb + 1);
      tc_fence_after();
      const uint32_t a0 = smem0 + (uint32_t)s * stage_stride, b0 = a0 + 16384u;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_tf32(tmem_base, umma_desc_sw128(a0 + ks * 32), umma_desc_sw128(b0 + ks * 32), idesc, (kb | ks) ? 1u : 0u);
      umma_commit(smem_u32(&bar_empty[s]));
    }
    umma_commit(smem_u32(&bar_acc));
  }
  // ---- epilogue: the eight producer warps drain the accumulator ------------------------------------------
  // TMEM hands every lane one row (32 consecutive columns per load).  Each warp transposes its 32x32 block
  // through a private 4 KB staging tile (the pipeline stages are idle by now) so that the epilogue's global
  // loads / stores again cover four full 128-byte row segments per instruction.
  __syncwarp();
  if (warp < 8) {
    const int q = warp & 3, half = warp >> 2;
    const int nch = (n_cta + 31) >> 5;
    const int g = lane & 7;
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(smem_u32(&bar_acc), 0)) {
        __nanosleep(64);
        if (++spins > (1u << 24)) { ok = false; break; }
      }
    }
    tc_fence_after();
    if (threadIdx.x == 0) VDN_TL(0, 2);
    if (ok) {
      const uint32_t stg = smem0 + (uint32_t)warp * 4096u;
      const int mbase = m0 + q * 32 + (lane >> 3);                  // row of group i: mbase + 4 i
      const uint32_t e7e = (uint32_t)(lane >> 3), e7o = e7e + 4u;
      const uint32_t roff_e = e7e * 128u + (((uint32_t)g ^ e7e) << 4);
      const uint32_t roff_o = e7o * 128u + (((uint32_t)g ^ e7o) << 4);
      const uint32_t woff = (uint32_t)lane * 128u;
#pragma unroll 1
      for (int ch = half; ch < nch; ch += 2) {
        {
          const int n = n_base + ch * 32 + g * 4;
          // bias of this thread's columns: issued before the TMEM load so its latency overlaps
          const float4 bias4 = (E.bias && vec_ok && n + 3 < N) ? __ldg(reinterpret_cast<const float4*>(E.bias + n))
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + woff + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4)),
                         "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                         : "memory");
          __syncwarp();
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                         : "r"(stg + ((i & 1) ? roff_o : roff_e) + (uint32_t)(i >> 1) * 1024u));
          __syncwarp();
          if (vec_ok && n + 3 < N) {
            tc_epilogue8(E, mbase, M, n, x, bias4);
          } else if (n < N) {
            tc_epilogue8_scalar(E, mbase, M, n, x, N);
          }
        }
      }
    } else if (fault) {
      *fault = 1;
    }
  }
  if (threadIdx.x == 0) VDN_TL(0, 3);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, ncols);
  if (threadIdx.x == 0) VDN_TL(0, 4);
  if (dbg && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1500) {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    dbg[1024 + 4 * blockIdx.x + 1] = (long long)g;
  }
#undef VDN_TL
}

// Operand of the weight side: plain row-major pointer for the FFMA kernels, swizzled tile image for tcgen05.
struct WeightRef {
  const float* w;    // [rows, ld] row-major, already offset to row0
  int ld;
  const float* img;  // image of the whole matrix: tiles [k-block][img_rows][32] in SWIZZLE_128B order, tf32-rounded
  int img_rows;
  int row0;
};

// W_l as the B operand of a forward-type GEMM, and rows [row0, ...) of W_l^T for the dgrad-type GEMMs.
inline WeightRef wref(const MlpLayout& ly, const float* packed, int l) {
  return WeightRef{packed + ly.off_w[l], ly.in_ld[l], packed + ly.off_iw[l], ly.out_ld[l], 0};
}
inline WeightRef wtref(const MlpLayout& ly, const float* packed, int l, int row0 = 0) {
  return WeightRef{packed + ly.off_wt[l] + (long long)row0 * ly.out_ld[l], ly.out_ld[l], packed + ly.off_iwt[l],
                   ly.in_ld[l], row0};
}

extern int g_mode;           // 0: exact fp32 (FFMA kernels), 1: tf32 tensor cores (tcgen05); set by vdn_set_mode
extern int* g_tc_fault;      // device flag raised by a timed-out barrier wait in a tcgen05 kernel
extern long long* g_tc_dbg;  // optional device buffer (8192 int64) receiving debug time stamps (vdn_debug_timeline)

static inline int launch_gemm_nt_tc(int M, int N, int K, const Operand& A, const WeightRef& B, const Epilogue& E,
                                    cudaStream_t st) {
  const int nkb = (K + TC_BK - 1) / TC_BK;
  const int n_mma_max = ((N < 256 ? N : 256) + 15) & ~15;
  const size_t stage = ((size_t)16384 + (size_t)n_mma_max * 128 + 1023) & ~(size_t)1023;
  const size_t smem = TC_STAGES * stage + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // N = 257 (a 256-wide head stacked with a scalar one): the last column rides along in the producers
  const float* wextra = nullptr;
  int n_main = N;
  if (N == 257 && B.row0 == 0 && B.w && (B.ld & 3) == 0 && ((uintptr_t)B.w & 15) == 0) {
    wextra = B.w + (size_t)256 * B.ld;
    n_main = 256;
  }
  dim3 grid((M + TC_BM - 1) / TC_BM, (n_main + 255) / 256);
  prof_begin(PROF_TC, st, 2.0 * M * N * K, operand_bytes(A, M) + epilogue_bytes(E, M, N));
  // debug timeline: record only the launch selected by VDN_DBG_LAUNCH (index since the buffer was installed)
  long long* dbg = g_tc_dbg;
  if (dbg) {
    static int sel = -2;
    static int count = 0;
    static long long* last = nullptr;
    if (sel == -2) { const char* ev = getenv("VDN_DBG_LAUNCH"); sel = ev ? atoi(ev) : -1; }
    if (last != dbg) { last = dbg; count = 0; }
    if (sel >= 0 && count++ != sel) dbg = nullptr;
  }
  VDN_LAUNCH(gemm_nt_tc_kernel, grid, TC_THREADS, smem, st, M, n_main, nkb, A, B.img, B.img_rows, B.row0, E,
             epilogue_vec_ok(E) ? 1 : 0, wextra, g_tc_fault, dbg);
  prof_end(PROF_TC, st);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

inline int launch_gemm_nt(int M, int N, int K, const Operand& A, const WeightRef& B, const Epilogue& E,
                          cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (g_mode == 1 && B.img && (B.row0 & 15) == 0 && operand_ok(A)) return launch_gemm_nt_tc(M, N, K, A, B, E, st);
  return launch_gemm_nt_simt(M, N, K, A, B.w, B.ld, E, st);
}

}  // namespace vdn

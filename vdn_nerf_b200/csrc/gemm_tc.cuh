// tcgen05 / TMEM GEMMs (tensor-core "tf32" mode of the MLP path), same operand-prologue / epilogue contract as
// the FFMA kernels of gemm_simt.cuh:
//
//   gemm_nt_tc : C[M,N] = epi( pro(A)[M,K] * W[N,K]^T + bias )
//
// One CTA owns a 128-row tile and up to 256 output columns; two CTAs are co-resident per SM so one tile's
// epilogue overlaps the other's MMAs.  Warp roles: warps 0-3 transform the A operand (global -> registers ->
// prologue -> tf32 round -> SWIZZLE_128B shared-memory image, one row per thread), one thread streams the
// pre-swizzled weight tiles with cp.async.bulk (TMA engine), one thread issues tcgen05.mma kind::tf32 into a
// TMEM accumulator, then all eight warps drain TMEM with tcgen05.ld and run the epilogue.  A 2-stage mbarrier
// ring (full / empty) connects them; every wait is bounded and raises a device fault flag instead of hanging.
#pragma once
#include "gemm_simt.cuh"
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace vdn {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 2, TC_THREADS = 256;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Vector (4 consecutive columns) form of epi_store for the elementwise epilogues; scalar fallback otherwise.
__device__ __forceinline__ void epi_store4(const Epilogue& e, bool vec_ok, int m, int n, float4 v, int N) {
  if (!vec_ok || n + 3 >= N) {
    const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) epi_store(e, m, n + j, a[j]);
    return;
  }
  if (e.bias) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  const size_t mm = (size_t)m;
  auto st4 = [](float* p, float4 x) { *reinterpret_cast<float4*>(p) = x; };
  auto ld4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
  switch (e.kind) {
    case EPI_STORE: st4(e.c + mm * e.ldc + e.coff + n, v); break;
    case EPI_RELU:
      st4(e.c + mm * e.ldc + e.coff + n, make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)));
      break;
    case EPI_SOFTPLUS:
      st4(e.c + mm * e.ldc + e.coff + n,
          make_float4(softplus100_fast(v.x), softplus100_fast(v.y), softplus100_fast(v.z), softplus100_fast(v.w)));
      break;
    case EPI_SDF_SKIP:
      if (e.c) st4(e.c + mm * e.ldc + n, v);
      st4(e.c2 + mm * e.ldc2 + n, make_float4(softplus100_fast(v.x) * e.scale, softplus100_fast(v.y) * e.scale,
                                              softplus100_fast(v.z) * e.scale, softplus100_fast(v.w) * e.scale));
      break;
    case EPI_GRAD_DUAL: {
      const float4 z = ld4(e.aux + mm * e.ldaux + n);
      float4 g = ld4(e.aux2 + mm * e.ldaux2 + n);
      g.x *= e.scale2; g.y *= e.scale2; g.z *= e.scale2; g.w *= e.scale2;
      st4(e.c + mm * e.ldc + n, make_float4(softplus100_d1_fast(z.x) * v.x * e.scale, softplus100_d1_fast(z.y) * v.y * e.scale,
                                            softplus100_d1_fast(z.z) * v.z * e.scale, softplus100_d1_fast(z.w) * v.w * e.scale));
      st4(e.c2 + mm * e.ldc2 + n, make_float4(softplus100_d2_fast(z.x) * g.x * v.x, softplus100_d2_fast(z.y) * g.y * v.y,
                                              softplus100_d2_fast(z.z) * g.z * v.z, softplus100_d2_fast(z.w) * g.w * v.w));
    } break;
    case EPI_BWD_INJECT: {
      const float4 z = ld4(e.aux + mm * e.ldaux + n);
      float4 r = make_float4(softplus100_d1_fast(z.x) * v.x * e.scale, softplus100_d1_fast(z.y) * v.y * e.scale,
                             softplus100_d1_fast(z.z) * v.z * e.scale, softplus100_d1_fast(z.w) * v.w * e.scale);
      if (e.aux2) {
        const float4 q = ld4(e.aux2 + mm * e.ldaux2 + n);
        r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
      }
      st4(e.c + mm * e.ldc + n, r);
    } break;
    case EPI_RELU_MASK: {
      const float4 h = ld4(e.aux + mm * e.ldaux + e.split + n);
      st4(e.c + mm * e.ldc + n, make_float4(h.x > 0.f ? v.x : 0.f, h.y > 0.f ? v.y : 0.f, h.z > 0.f ? v.z : 0.f,
                                            h.w > 0.f ? v.w : 0.f));
    } break;
    default: break;
  }
}

inline bool epilogue_vec_ok(const Epilogue& e) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (e.bias && !al(e.bias)) return false;
  switch (e.kind) {
    case EPI_STORE: case EPI_RELU: case EPI_SOFTPLUS:
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
                  Epilogue E, int vec_ok, int* __restrict__ fault, long long* __restrict__ dbg) {
  using namespace tc;
  // optional timeline of CTA 0 (debug): dbg[role*64 + event] = clock64()
  const bool rec = dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define VDN_TL(role, ev) do { if (rec) dbg[(role) * 64 + (ev)] = clock64(); } while (0)
  if (threadIdx.x == 0) VDN_TL(0, 0);
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
  const uint32_t stage_bytes = 16384u + (uint32_t)n_mma * 128u;
  auto sA = [&](int s) { return smem0 + (uint32_t)s * ((stage_bytes + 1023u) & ~1023u); };
  auto sB = [&](int s) { return sA(s) + 16384u; };

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 128 + 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    mbar_fence_init();
  }
  if (warp == 6) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;
  if (threadIdx.x == 0) VDN_TL(0, 1);

  if (warp < 4) {
    // ---- A producers: warp w owns rows [32w, 32w+32); per instruction the 32 lanes cover 4 rows x 8 chunks of
    // 16 bytes, i.e. four full 128-byte row segments (coalesced global loads, conflict-free swizzled stores) ----
    const int chunk = lane & 7;
    int rows[8];
    bool rok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rows[i] = warp * 32 + i * 4 + (lane >> 3);
      rok[i] = (m0 + rows[i]) < M;
    }
    RawLoad raw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) raw[i] = operand_load(A, m0 + rows[i], chunk * 4, rok[i]);
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = operand_finish_fast(A, raw[i], kb * 32 + chunk * 4, rok[i]);
        v[i] = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
      }
      if (kb + 1 < nkb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = operand_load(A, m0 + rows[i], (kb + 1) * 32 + chunk * 4, rok[i]);
      }
      if (tid == 0) VDN_TL(1, 3 * kb);
      ok = mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      if (tid == 0) VDN_TL(1, 3 * kb + 1);
      const uint32_t base = sA(s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r = (uint32_t)rows[i];
        const uint32_t addr = base + (r >> 3) * 1024u + (r & 7u) * 128u + (((uint32_t)chunk ^ (r & 7u)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[i].x), "f"(v[i].y), "f"(v[i].z),
                     "f"(v[i].w)
                     : "memory");
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bar_full[s]));
      if (tid == 0) VDN_TL(1, 3 * kb + 2);
    }
  } else if (tid == 128) {
    // ---- MMA issuer ------------------------------------------------------------------------------
    const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)n_mma);
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      VDN_TL(2, 2 * kb);
      ok = mbar_wait(smem_u32(&bar_full[s]), ph);
      VDN_TL(2, 2 * kb + 1);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_tf32(tmem_base, umma_desc_sw128(sA(s) + ks * 32), umma_desc_sw128(sB(s) + ks * 32), idesc,
                  (kb | ks) ? 1u : 0u);
      umma_commit(smem_u32(&bar_empty[s]));
    }
    umma_commit(smem_u32(&bar_acc));
  } else if (tid == 160) {
    // ---- weight tiles through the TMA engine ------------------------------------------------------
    const uint32_t bytes = (uint32_t)n_mma * 128u;
    for (int kb = 0; kb < nkb && ok; ++kb) {
      const int s = kb & 1, ph = (kb >> 1) & 1;
      ok = mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      VDN_TL(3, kb);
      mbar_arrive_expect_tx(smem_u32(&bar_full[s]), bytes);
      bulk_g2s(sB(s), Bimg + ((size_t)kb * img_rows + row0 + n_base) * 32, bytes, smem_u32(&bar_full[s]));
    }
  }
  // ---- epilogue: all eight warps drain the accumulator ------------------------------------------------
  // TMEM hands every lane one row (32 consecutive columns per load).  Each warp transposes its 32x32 block
  // through a private 4 KB staging tile (the pipeline stages are idle by now) so that the epilogue's global
  // loads / stores again cover four full 128-byte row segments per instruction.
  __syncwarp();
  {
    uint32_t spins = 0;
    while (!mbar_try_wait(smem_u32(&bar_acc), 0)) {
      if (warp != 4) __nanosleep(256);
      if (++spins > (1u << 24)) { ok = false; break; }
    }
  }
  tc_fence_after();
  if (threadIdx.x == 0) VDN_TL(0, 2);
  if (ok) {
    const int q = warp & 3, half = warp >> 2;
    const int nch = (n_cta + 31) >> 5;
    const uint32_t stg = smem0 + (uint32_t)warp * 4096u;
    const int g = lane & 7;
    for (int ch = half; ch < nch; ch += 2) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t addr = stg + (uint32_t)lane * 128u + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                     "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                     : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r = (uint32_t)(i * 4 + (lane >> 3));
        const uint32_t addr = stg + r * 128u + (((uint32_t)g ^ (r & 7u)) << 4);
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
        const int m = m0 + q * 32 + (int)r;
        if (m < M) epi_store4(E, vec_ok != 0, m, n_base + ch * 32 + g * 4, x, N);
      }
      __syncwarp();
    }
  } else if (fault) {
    *fault = 1;
  }
  if (threadIdx.x == 0) VDN_TL(0, 3);
  tc_fence_before();
  __syncthreads();
  if (warp == 6) tmem_dealloc(tmem_base, ncols);
  if (threadIdx.x == 0) VDN_TL(0, 4);
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

extern int g_mode;       // 0: exact fp32 (FFMA kernels), 1: tf32 tensor cores (tcgen05); set by vdn_set_mode
extern int* g_tc_fault;  // device flag raised by a timed-out barrier wait in a tcgen05 kernel
extern long long* g_tc_dbg;  // optional device buffer (256 int64) receiving CTA 0's timeline (vdn_debug_timeline)

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
  dim3 grid((M + TC_BM - 1) / TC_BM, (N + 255) / 256);
  prof_begin(PROF_TC, st, 2.0 * M * N * K);
  VDN_LAUNCH(gemm_nt_tc_kernel, grid, TC_THREADS, smem, st, M, N, nkb, A, B.img, B.img_rows, B.row0, E,
             epilogue_vec_ok(E) ? 1 : 0, g_tc_fault, g_tc_dbg);
  prof_end(PROF_TC, st);
  return (int)cudaGetLastError();
}

inline int launch_gemm_nt(int M, int N, int K, const Operand& A, const WeightRef& B, const Epilogue& E,
                          cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (g_mode == 1 && B.img && (B.row0 & 15) == 0 && operand_ok(A)) return launch_gemm_nt_tc(M, N, K, A, B, E, st);
  return launch_gemm_nt_simt(M, N, K, A, B.w, B.ld, E, st);
}

}  // namespace vdn

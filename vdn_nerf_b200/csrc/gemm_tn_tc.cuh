// tcgen05 weight-gradient GEMM (tensor-core mode):  P[s][n,k] = sum_{m in split s} proA(A)[m,n] * proX(X)[m,k]
//
// The reduction runs over the batch (points), so both operands are MN-major for the MMA: a shared-memory tile is
// [32 points x 32 features] in the SWIZZLE_128B_BASE32B image (rows = points; the only layout tf32 supports for
// MN-major operands), read by the tensor core with K = points.  One CTA owns 128 output rows (n) x up to 256 output columns (k) and a contiguous
// split of the batch; its accumulator stays in TMEM for the whole split and is written once, as a partial slab
// that reduce_partials_kernel sums deterministically (or, in the training path, added straight into dW with fp32
// atomics).  Sixteen producer warps transform both operands (coalesced loads two stages ahead in registers ->
// prologue -> tf32 round -> swizzled store), one thread issues the MMAs, a 4-stage mbarrier ring connects them;
// one CTA per SM, (output tiles x batch splits) sized to one wave over the 148 SMs.
#pragma once
#include "gemm_tc.cuh"

namespace vdn {

constexpr int TN_PWARPS = 16;     // producer (+ epilogue) warps
constexpr int TN_THREADS = TN_PWARPS * 32;         // warp 0 also allocates TMEM and issues the MMAs (one stage behind)
constexpr int TN_P = 32;          // points per pipeline stage
constexpr int TN_STAGES = 4;
constexpr uint32_t TN_TILE = 4096;                 // bytes of one [32 x 32] tile
constexpr uint32_t TN_STAGE = 12 * TN_TILE;        // 4 tiles of A (128 n) + 8 tiles of X (256 k)

__device__ __forceinline__ float4 tn_pro4(int kind, float scale, float4 a, float4 b) {
  switch (kind) {
    case PRO_SOFTPLUS: return f4_map_sp(a);
    case PRO_DSIG: return f4_scale(f4_mul(f4_map_sp1(b), a), scale);
    case PRO_DSIGMOID: return f4_mul(a, make_float4(b.x * (1.f - b.x), b.y * (1.f - b.y), b.z * (1.f - b.z), b.w * (1.f - b.w)));
    case PRO_RELUMASK: return make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
    default: return a;
  }
}

// Raw operand slices of one pipeline stage held by one producer thread (global loads run two stages ahead).
struct TnRaw {
  float4 a[2], b[2], x[4];
};

static __global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_tc_kernel(int M, int N, int K, Operand A, Operand X, float* __restrict__ P, int ldp, int rows_per_split,
                  int atomic_out, float* __restrict__ db, int* __restrict__ fault) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[TN_STAGES], bar_empty[TN_STAGES], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * 128;
  const int split = blockIdx.y;
  const int k0 = blockIdx.z * 256;
  const int kt = min(256, K - k0);
  const int k_mma = (kt + 15) & ~15;
  const int nxt = (kt + 31) >> 5;                       // X tiles per stage actually needed
  const int mbeg = split * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  const int nst = mend > mbeg ? (mend - mbeg + TN_P - 1) / TN_P : 0;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (tid == 0) {
    for (int s = 0; s < TN_STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), TN_PWARPS);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  {
    // ---- producers: warp w owns rows (points) 4(w%8)..+3 of half (w/8) of the stage's tiles (A tiles 2h,2h+1; X tiles
    // 4h..4h+3); a warp instruction covers 4 rows x 128 B.  Global loads run two stages ahead of their use. ----
    const int chunk = lane & 7, hh = warp >> 3;
    const uint32_t r = (uint32_t)((warp & 7) * 4 + (lane >> 3));   // point row inside the stage, 0..31
    // SWIZZLE_128B_BASE32B image: 4-row groups of 512 B, 32-byte blocks of row r permuted by (r % 4)
    const uint32_t soff = (r >> 2) * 512u + (r & 3u) * 128u + (((((uint32_t)chunk >> 1) ^ r) & 3u) << 5) +
                          (((uint32_t)chunk & 1u) << 4);
    const bool twoA = A.kind >= PRO_DSIG;   // the X side only takes single-operand prologues (none / softplus)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ca0 = n0 + (2 * hh) * 32 + chunk * 4;     // first A column of this thread (tile t: + 32 t)
    const int cx0 = k0 + (4 * hh) * 32 + chunk * 4;     // first X column
    // fused bias gradient: column sums of the A side (db[n] += sum_m A[m,n]) ride along for free
    const bool do_bias = db != nullptr && blockIdx.z == 0;
    float4 bsum[2] = {zero4, zero4};
    // column groups of this thread that touch the ragged edge of an operand (loop invariant; the common case is none)
    bool a_edge[2], x_edge[4];
#pragma unroll
    for (int t = 0; t < 2; ++t) a_edge[t] = ca0 + t * 32 + 3 >= (A.kvalid < A.width ? A.kvalid : A.width);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      x_edge[t] = (4 * hh + t >= nxt) || cx0 + t * 32 + 3 >= (X.kvalid < X.width ? X.kvalid : X.width);
    auto gload = [&](int st, TnRaw& R) {
      const int m = mbeg + st * TN_P + (int)r;
      const bool rok = m < mend;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = ca0 + t * 32;
        const bool p = rok && c < A.width;
        R.a[t] = p ? *reinterpret_cast<const float4*>(A.p + (size_t)m * A.ld + c) : zero4;
        R.b[t] = (p && twoA) ? *reinterpret_cast<const float4*>(A.p2 + (size_t)m * A.ld2 + c) : zero4;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = cx0 + t * 32;
        const bool p = (4 * hh + t < nxt) && rok && c < X.width;
        R.x[t] = p ? *reinterpret_cast<const float4*>(X.p + (size_t)m * X.ld + c) : zero4;
      }
    };
    // MMA issue (warp 0 only, one stage behind its own production): 4 MMAs (8 points each) per stage, both operands
    // MN-major
    const uint32_t idesc = umma_idesc_tf32_mn(128, (uint32_t)k_mma);
    auto issue = [&](int st) {
      const int s = st % TN_STAGES, ph = (st / TN_STAGES) & 1;
      ok = mbar_wait(smem_u32(&bar_full[s]), ph) && ok;
      tc_fence_after();
      if (lane == 0 && ok) {
        const uint32_t a0 = smem0 + (uint32_t)s * TN_STAGE, x0 = a0 + 4 * TN_TILE;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          umma_tf32(tmem_base, umma_desc_sw128_mn(a0 + g * 1024, TN_TILE, 512), umma_desc_sw128_mn(x0 + g * 1024, TN_TILE, 512),
                    idesc, (st | g) ? 1u : 0u);
        umma_commit(smem_u32(&bar_empty[s]));
      }
      __syncwarp();
    };
    auto produce = [&](int st, TnRaw& R) {
      const int s = st % TN_STAGES, ph = (st / TN_STAGES) & 1;
      const bool rok = mbeg + st * TN_P + (int)r < mend;
      const bool tail = mbeg + (st + 1) * TN_P > mend;     // uniform: only the split's last stage can hold invalid rows
      float4 va[2], vx[4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float4 v = tn_pro4(A.kind, A.scale, R.a[t], R.b[t]);
        if (a_edge[t] | tail) {                        // ragged column edge of the operand or the split's last rows
          const int c = ca0 + t * 32;
          if (!rok || c >= A.width) v = zero4;
          if (c + 0 >= A.kvalid) v.x = 0.f;
          if (c + 1 >= A.kvalid) v.y = 0.f;
          if (c + 2 >= A.kvalid) v.z = 0.f;
          if (c + 3 >= A.kvalid) v.w = 0.f;
        }
        bsum[t] = f4_add(bsum[t], v);
        va[t] = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float4 v = (X.kind == PRO_SOFTPLUS) ? f4_map_sp(R.x[t]) : R.x[t];
        if (x_edge[t] | tail) {
          const int c = cx0 + t * 32;
          if (!((4 * hh + t < nxt) && rok && c < X.width)) v = zero4;
          if (c + 0 >= X.kvalid) v.x = 0.f;
          if (c + 1 >= X.kvalid) v.y = 0.f;
          if (c + 2 >= X.kvalid) v.z = 0.f;
          if (c + 3 >= X.kvalid) v.w = 0.f;
        }
        vx[t] = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
      if (st + 2 < nst) gload(st + 2, R);              // refill the register buffer just consumed
      ok = mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
      const uint32_t base = smem0 + (uint32_t)s * TN_STAGE + soff;
#pragma unroll
      for (int t = 0; t < 2; ++t)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(2 * hh + t) * TN_TILE), "f"(va[t].x),
                     "f"(va[t].y), "f"(va[t].z), "f"(va[t].w)
                     : "memory");
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * hh + t < nxt)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(4 + 4 * hh + t) * TN_TILE),
                       "f"(vx[t].x), "f"(vx[t].y), "f"(vx[t].z), "f"(vx[t].w)
                       : "memory");
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
      if (warp == 0 && st > 0 && ok) issue(st - 1);
    };
    TnRaw R0, R1;
    if (nst > 0) gload(0, R0);
    if (nst > 1) gload(1, R1);
    for (int st = 0; st < nst && ok; st += 2) {
      produce(st, R0);
      if (st + 1 < nst && ok) produce(st + 1, R1);
    }
    if (warp == 0) {
      if (nst > 0 && ok) issue(nst - 1);
      if (lane == 0) umma_commit(smem_u32(&bar_acc));
      __syncwarp();
    }
    if (do_bias) {   // reduce over the 4 row lanes of the warp, then one atomic per column and warp
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float4 v = bsum[t];
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
          v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
          v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
          v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
          v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
        }
        const int n = ca0 + t * 32;
        if ((lane >> 3) == 0) {
          if (n + 0 < N) atomicAdd(db + n + 0, v.x);
          if (n + 1 < N) atomicAdd(db + n + 1, v.y);
          if (n + 2 < N) atomicAdd(db + n + 2, v.z);
          if (n + 3 < N) atomicAdd(db + n + 3, v.w);
        }
      }
    }
  }
  // ---- epilogue: the producer warps write the partial tile ----------------------------------------------
  __syncwarp();
  {
    uint32_t spins = 0;
    while (!mbar_try_wait(smem_u32(&bar_acc), 0)) {
      __nanosleep(100);
      if (++spins > (1u << 24)) { ok = false; break; }
    }
    tc_fence_after();
    if (ok) {
      const int q = warp & 3, cset = warp >> 2;
      const int nch = (kt + 31) >> 5;
      const uint32_t stg = smem0 + (uint32_t)warp * 4096u;
      const int g = lane & 7;
      const int nbase = n0 + q * 32 + (lane >> 3);
      // atomic_out: every split adds its tile straight into the gradient matrix (red.global.add, no partial slabs)
      float* Ps = atomic_out ? P : P + (size_t)split * N * ldp;
      const bool skip_all = atomic_out && nst == 0;   // an empty split has nothing to add
      for (int ch = cset; ch < nch && !skip_all; ch += TN_PWARPS / 4) {
        float v[32];
        if (nst > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 128u +
                                                                       (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4)),
                       "f"(v[4 * c]), "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t rr = (uint32_t)(i * 4 + (lane >> 3));
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"(stg + rr * 128u + (((uint32_t)g ^ (rr & 7u)) << 4)));
          const int n = nbase + 4 * i;
          const int k = k0 + ch * 32 + g * 4;
          if (n < N && atomic_out) {
            float* dst = Ps + (size_t)n * ldp + k;
            if (k + 3 < K && (ldp & 3) == 0) {
              atomicAdd(reinterpret_cast<float4*>(dst), x);
            } else {
              if (k + 0 < K) atomicAdd(dst + 0, x.x);
              if (k + 1 < K) atomicAdd(dst + 1, x.y);
              if (k + 2 < K) atomicAdd(dst + 2, x.z);
              if (k + 3 < K) atomicAdd(dst + 3, x.w);
            }
          } else if (n < N) {
            float* dst = Ps + (size_t)n * ldp + k;
            if (k + 3 < K && (ldp & 3) == 0) {
              *reinterpret_cast<float4*>(dst) = x;
            } else {
              if (k + 0 < K) dst[0] = x.x;
              if (k + 1 < K) dst[1] = x.y;
              if (k + 2 < K) dst[2] = x.z;
              if (k + 3 < K) dst[3] = x.w;
            }
          }
        }
        __syncwarp();
      }
    } else if (fault) {
      *fault = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// Row splits of the batch for the tensor-core weight gradient: one wave of (output tiles x splits) CTAs over the
// 148 SMs, at least 128 points per split.  tiles = 1 gives the upper bound used to size the partial slabs.
inline int wgrad_splits_tc(int M, int tiles = 1) {
  int s = 148 / (tiles < 1 ? 1 : tiles);
  const int cap = (M + 127) / 128;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return s;
}

// In the tensor-core mode the splits accumulate into dW with fp32 atomics (red.global.add.v4.f32): no partial slabs,
// no second kernel; the summation order over splits is then not deterministic (differences at the 1e-7 level).
// `accumulate` must be 1 (the packed gradient buffer is zero-initialised by the caller).
static inline int launch_wgrad_tc(int M, int N, int K, const Operand& A0, const Operand& X0, float* partials, float* dW,
                                  int ldd, int accumulate, float* db, cudaStream_t st) {
  const int S = wgrad_splits_tc(M, ((N + 127) / 128) * ((K + 255) / 256));
  int rows = (M + S - 1) / S;
  rows = (rows + TN_P - 1) / TN_P * TN_P;
  static bool attr_set = false;
  const size_t smem = TN_STAGES * TN_STAGE + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((N + 127) / 128, S, (K + 255) / 256);
  prof_begin(PROF_WGRAD, st, 2.0 * M * N * K, operand_bytes(A0, M) + operand_bytes(X0, M));
  if (accumulate) {
    VDN_LAUNCH(gemm_tn_tc_kernel, grid, TN_THREADS, smem, st, M, N, K, A0, X0, dW, ldd, rows, 1, db, g_tc_fault);
  } else {
    const int ldp = (K + 3) & ~3;
    VDN_LAUNCH(gemm_tn_tc_kernel, grid, TN_THREADS, smem, st, M, N, K, A0, X0, partials, ldp, rows, 0, db, g_tc_fault);
    int e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
    const int total = N * K;
    VDN_LAUNCH(reduce_partials_kernel, (total + 255) / 256, 256, 0, st, partials, S, N, K, ldp, dW, ldd, 0);
  }
  prof_end(PROF_WGRAD, st);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// Mode dispatch for the weight gradient (single operand pair).  db (nullable): bias gradient, db[n] += sum_m A0[m,n];
// the tensor-core kernel fuses it, the fp32 path runs the column-sum kernels.
inline int launch_wgrad_any(int M, int N, int K, const Operand& A0, const Operand& X0, float* partials, float* dW, int ldd,
                            int accumulate, float* db, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (g_mode == 1 && operand_ok(A0) && operand_ok(X0) && X0.kind <= PRO_SOFTPLUS)
    return launch_wgrad_tc(M, N, K, A0, X0, partials, dW, ldd, accumulate, db, st);
  int e = launch_wgrad(M, N, K, A0, X0, nullptr, nullptr, partials, dW, ldd, accumulate, st);
  if (e || !db) return e;
  return launch_colsum(M, N, A0, partials, db, 1, st);
}

// Number of partial slabs either path may write (for workspace sizing).
inline long long wgrad_max_splits(long long M) {
  int a = wgrad_splits((int)M), b = wgrad_splits_tc((int)M);
  return a > b ? a : b;
}

}  // namespace vdn

// Exact-fp32 GEMM building blocks (CUDA-core FFMA) with fused operand prologues and output epilogues.
//
// These kernels are the "fp32 mode" of the MLP path (parity <= 1e-5 relative against the reference's
// fp32 PyTorch path) and the fallback for shapes the tcgen05 chain kernels do not cover.  Two forms:
//
//   gemm_nt : C[M,N] = epi( pro(A)[M,K] * B[N,K]^T + bias )      forward / dgrad (B = W or W^T, packed)
//   gemm_tn : P[s][N,K] = sum_{m in split s} proA(A)[m,n] * proX(X)[m,k]   wgrad partials (deterministic)
//
// Operand prologues recompute activations from stored pre-activations (softplus, softplus', ...), so the
// layer-wise path stores one tensor per layer.  All operand buffers have leading dimensions that are
// multiples of 4 floats and are 16-byte aligned (checked by the launchers).
#pragma once
#include "common.cuh"

namespace vdn {

enum ProKind : int {
  PRO_NONE = 0,      // p[m,k]
  PRO_SOFTPLUS = 1,  // softplus100(p[m,k])
  PRO_DSIG = 2,      // softplus100'(p2[m,k]) * p[m,k] * scale      (p may be a broadcast row: ld == 0)
  PRO_DSIGMOID = 3,  // p[m,k] * p2[m,k] * (1 - p2[m,k])            (sigmoid backward from its output)
  PRO_RELUMASK = 4,  // p2[m,k] > 0 ? p[m,k] : 0
};

struct Operand {
  const float* p;
  const float* p2;
  int ld, ld2;
  int width;   // readable columns (multiple of 4); columns >= width read as 0 without touching memory
  int kvalid;  // logical columns; columns >= kvalid read as 0
  int kind;
  float scale;
  int rounded;  // values are already tf32-representable (written by an epilogue with round_c): the tensor-core
                // kernels may feed them to the MMA without their own rounding pass
};

__host__ __device__ inline Operand make_operand(const float* p, int ld, int width, int kvalid, int kind = PRO_NONE,
                                                const float* p2 = nullptr, int ld2 = 0, float scale = 1.0f) {
  Operand o;
  o.p = p; o.p2 = p2; o.ld = ld; o.ld2 = ld2; o.width = width; o.kvalid = kvalid; o.kind = kind; o.scale = scale; o.rounded = 0;
  return o;
}

struct RawLoad {
  float4 a, b;
};

__device__ __forceinline__ RawLoad operand_load(const Operand& o, int m, int c, bool row_ok) {
  RawLoad r;
  r.a = make_float4(0.f, 0.f, 0.f, 0.f);
  r.b = r.a;
  if (row_ok && c < o.width) {
    r.a = *reinterpret_cast<const float4*>(o.p + (size_t)m * o.ld + c);
    if (o.kind >= PRO_DSIG) r.b = *reinterpret_cast<const float4*>(o.p2 + (size_t)m * o.ld2 + c);
  }
  return r;
}

__device__ __forceinline__ float pro_apply(int kind, float a, float b, float scale) {
  switch (kind) {
    case PRO_SOFTPLUS: return softplus100(a);
    case PRO_DSIG: return softplus100_d1(b) * a * scale;
    case PRO_DSIGMOID: return a * b * (1.0f - b);
    case PRO_RELUMASK: return b > 0.0f ? a : 0.0f;
    default: return a;
  }
}

__device__ __forceinline__ float pro_apply_fast(int kind, float a, float b, float scale) {
  switch (kind) {
    case PRO_SOFTPLUS: return softplus100_fast(a);
    case PRO_DSIG: return softplus100_d1_fast(b) * a * scale;
    case PRO_DSIGMOID: return a * b * (1.0f - b);
    case PRO_RELUMASK: return b > 0.0f ? a : 0.0f;
    default: return a;
  }
}

// tensor-core mode: same as operand_finish with the MUFU-based activations
__device__ __forceinline__ float4 operand_finish_fast(const Operand& o, const RawLoad& r, int c, bool row_ok) {
  float4 v;
  if (!row_ok || c >= o.width) return make_float4(0.f, 0.f, 0.f, 0.f);
  v.x = (c + 0 < o.kvalid) ? pro_apply_fast(o.kind, r.a.x, r.b.x, o.scale) : 0.f;
  v.y = (c + 1 < o.kvalid) ? pro_apply_fast(o.kind, r.a.y, r.b.y, o.scale) : 0.f;
  v.z = (c + 2 < o.kvalid) ? pro_apply_fast(o.kind, r.a.z, r.b.z, o.scale) : 0.f;
  v.w = (c + 3 < o.kvalid) ? pro_apply_fast(o.kind, r.a.w, r.b.w, o.scale) : 0.f;
  return v;
}

__device__ __forceinline__ float4 operand_finish(const Operand& o, const RawLoad& r, int c, bool row_ok) {
  float4 v;
  if (!row_ok || c >= o.width) return make_float4(0.f, 0.f, 0.f, 0.f);
  v.x = (c + 0 < o.kvalid) ? pro_apply(o.kind, r.a.x, r.b.x, o.scale) : 0.f;
  v.y = (c + 1 < o.kvalid) ? pro_apply(o.kind, r.a.y, r.b.y, o.scale) : 0.f;
  v.z = (c + 2 < o.kvalid) ? pro_apply(o.kind, r.a.z, r.b.z, o.scale) : 0.f;
  v.w = (c + 3 < o.kvalid) ? pro_apply(o.kind, r.a.w, r.b.w, o.scale) : 0.f;
  return v;
}

enum EpiKind : int {
  EPI_STORE = 0,      // c[m, coff+n] = v
  EPI_RELU = 1,       // c[m, coff+n] = max(v, 0)
  EPI_SIGMOID = 2,    // c[m, coff+n] = sigmoid(v)
  EPI_SDF_SKIP = 3,   // c[m,n] = v ; c2[m,n] = softplus100(v) * scale
  EPI_SPLIT = 4,      // n < split: c2[m*ldc2+n] = v*scale ; else c[m*ldc + coff + n-split] = v   (null ptr: skip)
  EPI_ADD_SCALED = 5, // c[m,n] = v + scale * aux[m*ldaux + split + n]
  EPI_GRAD_DUAL = 6,  // z=aux[m,n], gin=aux2[m*ldaux2+n]*scale2 : c[m,n] = sp'(z)*v*scale ; c2[m,n] = sp''(z)*gin*v
  EPI_BWD_INJECT = 7, // c[m,n] = sp'(aux[m,n]) * v * scale + (aux2 ? aux2[m,n] : 0)
  EPI_RELU_MASK = 8,  // c[m,n] = aux[m*ldaux + split + n] > 0 ? v : 0
  EPI_SOFTPLUS = 9,   // c[m, coff+n] = softplus100(v)
};

struct Epilogue {
  int kind;
  const float* bias;
  float* c;
  float* c2;
  const float* aux;
  const float* aux2;
  int ldc, ldc2, ldaux, ldaux2;
  int split, coff;
  float scale, scale2;
  int round_c;  // tensor-core mode: store c rounded to tf32 (intermediates that only ever feed GEMM operands; the
                // consumer would round them anyway, so the numerics are unchanged and its rounding pass is saved)
};

__host__ __device__ inline Epilogue make_epilogue(int kind, const float* bias, float* c, int ldc) {
  Epilogue e;
  e.kind = kind; e.bias = bias; e.c = c; e.c2 = nullptr; e.aux = nullptr; e.aux2 = nullptr;
  e.ldc = ldc; e.ldc2 = 0; e.ldaux = 0; e.ldaux2 = 0; e.split = 0; e.coff = 0; e.scale = 1.0f; e.scale2 = 1.0f; e.round_c = 0;
  return e;
}

__device__ __forceinline__ void epi_store(const Epilogue& e, int m, int n, float v) {
  if (e.bias) v += e.bias[n];
  const size_t mm = (size_t)m;
  switch (e.kind) {
    case EPI_STORE: e.c[mm * e.ldc + e.coff + n] = v; break;
    case EPI_RELU: e.c[mm * e.ldc + e.coff + n] = fmaxf(v, 0.0f); break;
    case EPI_SIGMOID: e.c[mm * e.ldc + e.coff + n] = sigmoidf_(v); break;
    case EPI_SOFTPLUS: e.c[mm * e.ldc + e.coff + n] = softplus100(v); break;
    case EPI_SDF_SKIP:
      if (e.c) e.c[mm * e.ldc + n] = v;
      e.c2[mm * e.ldc2 + n] = softplus100(v) * e.scale;
      break;
    case EPI_SPLIT:
      if (n < e.split) {
        if (e.c2) e.c2[mm * e.ldc2 + n] = v * e.scale;
      } else {
        if (e.c) e.c[mm * e.ldc + e.coff + (n - e.split)] = v;
      }
      break;
    case EPI_ADD_SCALED: e.c[mm * e.ldc + n] = v + e.scale * e.aux[mm * e.ldaux + e.split + n]; break;
    case EPI_GRAD_DUAL: {
      float z = e.aux[mm * e.ldaux + n];
      float gin = e.aux2[mm * e.ldaux2 + n] * e.scale2;
      e.c[mm * e.ldc + n] = softplus100_d1(z) * v * e.scale;
      e.c2[mm * e.ldc2 + n] = softplus100_d2(z) * gin * v;
    } break;
    case EPI_BWD_INJECT: {
      float z = e.aux[mm * e.ldaux + n];
      float r = softplus100_d1(z) * v * e.scale;
      if (e.aux2) r += e.aux2[mm * e.ldaux2 + n];
      e.c[mm * e.ldc + n] = r;
    } break;
    case EPI_RELU_MASK: e.c[mm * e.ldc + n] = e.aux[mm * e.ldaux + e.split + n] > 0.0f ? v : 0.0f; break;
    default: break;
  }
}

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16, GEMM_THREADS = 256, GEMM_PAD = 4;

// 8x8 micro-tile FMA on one BK slab held in shared memory (sa: [BK][BM+PAD], sb: [BK][BN+PAD]).
__device__ __forceinline__ void tile_fma(const float (*sa)[GEMM_BM + GEMM_PAD], const float (*sb)[GEMM_BN + GEMM_PAD],
                                         int ty, int tx, float (&acc)[8][8]) {
#pragma unroll
  for (int k = 0; k < GEMM_BK; ++k) {
    float4 a0 = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
    float4 a1 = *reinterpret_cast<const float4*>(&sa[k][64 + ty * 4]);
    float4 b0 = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
    float4 b1 = *reinterpret_cast<const float4*>(&sb[k][64 + tx * 4]);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// C = epi(pro(A) * B^T).  A: [M, K] through an Operand; B: [>=N rows, ldb] row-major, zero padded in K.
// K must be a multiple of GEMM_BK; A.width / kvalid mask the logical extent.
static __global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_nt_kernel(int M, int N, int K, Operand A, const float* __restrict__ B, int ldb, Epilogue E) {
  __shared__ __align__(16) float sa[2][GEMM_BK][GEMM_BM + GEMM_PAD];
  __shared__ __align__(16) float sb[2][GEMM_BK][GEMM_BN + GEMM_PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;
  const int tx = tid & 15, ty = tid >> 4;

  // loader mapping: 2 float4 per thread per operand; row = idx/4, k4 = (idx%4)*4
  int lrow[2], lk[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * GEMM_THREADS;
    lrow[i] = idx >> 2;
    lk[i] = (idx & 3) * 4;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  RawLoad ra[2];
  float4 rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = m0 + lrow[i];
      ra[i] = operand_load(A, m, k0 + lk[i], m < M);
      int n = n0 + lrow[i];
      rb[i] = (n < N) ? *reinterpret_cast<const float4*>(B + (size_t)n * ldb + k0 + lk[i])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf, int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = m0 + lrow[i];
      float4 v = operand_finish(A, ra[i], k0 + lk[i], m < M);
      sa[buf][lk[i] + 0][lrow[i]] = v.x;
      sa[buf][lk[i] + 1][lrow[i]] = v.y;
      sa[buf][lk[i] + 2][lrow[i]] = v.z;
      sa[buf][lk[i] + 3][lrow[i]] = v.w;
      sb[buf][lk[i] + 0][lrow[i]] = rb[i].x;
      sb[buf][lk[i] + 1][lrow[i]] = rb[i].y;
      sb[buf][lk[i] + 2][lrow[i]] = rb[i].z;
      sb[buf][lk[i] + 3][lrow[i]] = rb[i].w;
    }
  };

  const int nk = K / GEMM_BK;
  gload(0);
  sstore(0, 0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GEMM_BK);
    tile_fma(sa[cur], sb[cur], ty, tx, acc);
    if (kt + 1 < nk) sstore(cur ^ 1, (kt + 1) * GEMM_BK);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) epi_store(E, m, n, acc[i][j]);
    }
  }
}

// Split-M weight-gradient partials: P[s][n][k] = sum over the rows of split s and over up to two operand
// pairs of proA(A)[m,n] * proX(X)[m,k].  Grid: (ceil(K/BN), ceil(N/BM), S).  P is [S][N][ldp].
static __global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tn_kernel(int M, int N, int K, Operand A0, Operand X0, Operand A1, Operand X1, int npairs,
               float* __restrict__ P, int ldp, int rows_per_split) {
  __shared__ __align__(16) float sa[2][GEMM_BK][GEMM_BM + GEMM_PAD];
  __shared__ __align__(16) float sb[2][GEMM_BK][GEMM_BN + GEMM_PAD];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * GEMM_BN, n0 = blockIdx.y * GEMM_BM;
  const int split = blockIdx.z;
  const int mbeg = split * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  const int tx = tid & 15, ty = tid >> 4;
  // loader mapping: rows r = idx/32 (0..15), col4 = (idx%32)*4
  int lr[2], lc[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * GEMM_THREADS;
    lr[i] = idx >> 5;
    lc[i] = (idx & 31) * 4;
  }
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  const int steps_per_pair = (mend > mbeg) ? (mend - mbeg + GEMM_BK - 1) / GEMM_BK : 0;
  const int nsteps = steps_per_pair * npairs;
  RawLoad ra[2], rx[2];
  auto gload = [&](int step) {
    const int pair = step / steps_per_pair;
    const int mm = mbeg + (step - pair * steps_per_pair) * GEMM_BK;
    const Operand& A = pair ? A1 : A0;
    const Operand& X = pair ? X1 : X0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = mm + lr[i];
      ra[i] = operand_load(A, m, n0 + lc[i], m < mend);
      rx[i] = operand_load(X, m, k0 + lc[i], m < mend);
    }
  };
  auto sstore = [&](int buf, int step) {
    const int pair = step / steps_per_pair;
    const int mm = mbeg + (step - pair * steps_per_pair) * GEMM_BK;
    const Operand& A = pair ? A1 : A0;
    const Operand& X = pair ? X1 : X0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = mm + lr[i];
      float4 va = operand_finish(A, ra[i], n0 + lc[i], m < mend);
      float4 vx = operand_finish(X, rx[i], k0 + lc[i], m < mend);
      *reinterpret_cast<float4*>(&sa[buf][lr[i]][lc[i]]) = va;
      *reinterpret_cast<float4*>(&sb[buf][lr[i]][lc[i]]) = vx;
    }
  };
  if (nsteps > 0) {
    gload(0);
    sstore(0, 0);
  }
  __syncthreads();
  for (int st = 0; st < nsteps; ++st) {
    const int cur = st & 1;
    if (st + 1 < nsteps) gload(st + 1);
    tile_fma(sa[cur], sb[cur], ty, tx, acc);
    if (st + 1 < nsteps) sstore(cur ^ 1, st + 1);
    __syncthreads();
  }
  float* Ps = P + (size_t)split * N * ldp;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (k < K) Ps[(size_t)n * ldp + k] = acc[i][j];
    }
  }
}

// dst[i] (+)= sum_s P[s][i]  for i in [0, rows*ldp) restricted to k < K; dst is [rows, ldd].
static __global__ void reduce_partials_kernel(const float* __restrict__ P, int S, int rows, int K, int ldp,
                                       float* __restrict__ dst, int ldd, int accumulate) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = rows * K;
  if (idx >= total) return;
  int r = idx / K, k = idx - r * K;
  float s = 0.0f;
  for (int i = 0; i < S; ++i) s += P[((size_t)i * rows + r) * ldp + k];
  float* d = dst + (size_t)r * ldd + k;
  *d = accumulate ? (*d + s) : s;
}

// Column sums of an operand: P[s][n] = sum over rows of split s of pro(A)[m,n].  Grid (ceil(N/128), S), 256 threads:
// 32 lanes x float4 cover 128 columns, 8 row lanes stride the split with four independent loads in flight.
static __global__ void colsum_partial_kernel(int M, int N, Operand A, float* __restrict__ P, int rows_per_split) {
  __shared__ float red[8][128];
  const int n4 = (threadIdx.x & 31) * 4;
  const int rl = threadIdx.x >> 5;  // 0..7
  const int n0 = blockIdx.x * 128;
  const int mbeg = blockIdx.y * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int m = mbeg + rl; m < mend; m += 32) {
    RawLoad r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = operand_load(A, m + 8 * u, n0 + n4, m + 8 * u < mend);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 v = operand_finish(A, r[u], n0 + n4, m + 8 * u < mend);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[rl][n4 + 0] = s.x; red[rl][n4 + 1] = s.y; red[rl][n4 + 2] = s.z; red[rl][n4 + 3] = s.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    int n = n0 + threadIdx.x;
    if (n < N) P[(size_t)blockIdx.y * N + n] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------------
inline bool operand_ok(const Operand& o) {
  if (!o.p) return false;
  if (((uintptr_t)o.p & 15) || (o.ld & 3) || (o.width & 3)) return false;
  if (o.kind >= PRO_DSIG && (!o.p2 || ((uintptr_t)o.p2 & 15) || (o.ld2 & 3))) return false;
  return true;
}

// Optional per-family timing (vdn_prof_enable): CUDA events around the GEMM launches on the launching stream.
void prof_begin(int family, cudaStream_t st, double flops, double bytes = 0.0);
void prof_end(int family, cudaStream_t st);
enum { PROF_GEMM_NT = 0, PROF_WGRAD = 1, PROF_TC = 2, PROF_CHAIN = 3, PROF_CHAIN_TRAIN = 4, PROF_WGRAD16 = 5, PROF_FAMILIES = 6 };

// Algorithmic HBM bytes of one GEMM launch: every operand / output element moved once (weights are L2 resident).
inline double operand_bytes(const Operand& A, double M) {
  return 4.0 * M * (A.kvalid < A.width ? A.kvalid : A.width) * (A.kind >= PRO_DSIG ? 2.0 : 1.0);
}
inline double epilogue_bytes(const Epilogue& E, double M, double N) {
  double per = 1.0;
  switch (E.kind) {
    case EPI_SDF_SKIP: per = E.c ? 2.0 : 1.0; break;
    case EPI_ADD_SCALED: case EPI_RELU_MASK: per = 2.0; break;
    case EPI_GRAD_DUAL: per = (E.ldaux2 ? 4.0 : 3.0); break;
    case EPI_BWD_INJECT: per = E.aux2 ? 3.0 : 2.0; break;
    default: break;
  }
  return 4.0 * M * N * per;
}

inline int launch_gemm_nt_simt(int M, int N, int K, const Operand& A, const float* B, int ldb, const Epilogue& E,
                          cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (K % GEMM_BK != 0 || !operand_ok(A) || ((uintptr_t)B & 15) || (ldb & 3)) return (int)cudaErrorInvalidValue;
  dim3 grid((M + GEMM_BM - 1) / GEMM_BM, (N + GEMM_BN - 1) / GEMM_BN);
  prof_begin(PROF_GEMM_NT, st, 2.0 * M * N * K, operand_bytes(A, M) + epilogue_bytes(E, M, N));
  VDN_LAUNCH(gemm_nt_kernel, grid, GEMM_THREADS, 0, st, M, N, K, A, B, ldb, E);
  prof_end(PROF_GEMM_NT, st);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// Number of row splits used for a wgrad over M rows (also the number of partial slabs needed).
inline int wgrad_splits(int M) {
  int s = (M + 1023) / 1024;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}

// dW[N, K] (+)= sum_m A[m,n] X[m,k] (+ second pair).  `partials` must hold wgrad_splits(M)*N*ldp floats.
inline int launch_wgrad(int M, int N, int K, const Operand& A0, const Operand& X0, const Operand* A1,
                        const Operand* X1, float* partials, float* dW, int ldd, int accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (!operand_ok(A0) || !operand_ok(X0)) return (int)cudaErrorInvalidValue;
  if (A1 && (!operand_ok(*A1) || !operand_ok(*X1))) return (int)cudaErrorInvalidValue;
  const int S = wgrad_splits(M);
  const int rows = (M + S - 1) / S;
  const int rows_per_split = (rows + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
  const int ldp = K;
  dim3 grid((K + GEMM_BN - 1) / GEMM_BN, (N + GEMM_BM - 1) / GEMM_BM, S);
  prof_begin(PROF_WGRAD, st, 2.0 * M * N * K * (A1 ? 2 : 1), operand_bytes(A0, M) + operand_bytes(X0, M));
  VDN_LAUNCH(gemm_tn_kernel, grid, GEMM_THREADS, 0, st, M, N, K, A0, X0, A1 ? *A1 : A0, X1 ? *X1 : X0, A1 ? 2 : 1,
                                                partials, ldp, rows_per_split);
  int e = (int)(cudaError_t)::vdn::take_launch_error();
  if (e) return e;
  int total = N * K;
  VDN_LAUNCH(reduce_partials_kernel, (total + 255) / 256, 256, 0, st, partials, S, N, K, ldp, dW, ldd, accumulate);
  prof_end(PROF_WGRAD, st);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// out[n] (+)= sum_m pro(A)[m,n];  `partials` must hold wgrad_splits(M)*N floats.
inline int colsum_splits(int M) {
  int s = (M + 255) / 256;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}

// out[n] (+)= sum_m pro(A)[m,n];  `partials` must hold colsum_splits(M)*N floats.
inline int launch_colsum(int M, int N, const Operand& A, float* partials, float* out, int accumulate,
                         cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (!operand_ok(A)) return (int)cudaErrorInvalidValue;
  const int S = colsum_splits(M);
  const int rows_per_split = (M + S - 1) / S;
  dim3 grid((N + 127) / 128, S);
  VDN_LAUNCH(colsum_partial_kernel, grid, 256, 0, st, M, N, A, partials, rows_per_split);
  int e = (int)(cudaError_t)::vdn::take_launch_error();
  if (e) return e;
  VDN_LAUNCH(reduce_partials_kernel, (N + 255) / 256, 256, 0, st, partials, S, 1, N, N, out, N, accumulate);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

}  // namespace vdn

// Stand-alone probe for the tcgen05 building blocks in csrc/tc_common.cuh (build: tools/build_probe.sh; run on
// the B200 box).  Validates, against an exact integer-valued reference GEMM:
//   - the SWIZZLE_128B K-major operand image + shared-memory descriptor + tf32 instruction descriptor,
//   - K-stepping inside and across 32-column K blocks, N = 256 / 224 / 16,
//   - cp.async.bulk of a pre-swizzled weight tile completing on an mbarrier,
//   - tcgen05.commit -> mbarrier and tcgen05.ld 32x32b accumulator read-back,
// and measures the issue rate of back-to-back tf32 MMAs.  Every wait is bounded; nothing can hang.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../vdn_nerf_b200/csrc/tc_common.cuh"

using namespace vdn::tc;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

// D[128, N] = A[128, K] * B[N, K]^T with K a multiple of 32.  mode bit0: B tiles arrive by cp.async.bulk from a
// pre-swizzled global image (Bimg) instead of being written by the threads.
__global__ void __launch_bounds__(128) probe_gemm(const float* __restrict__ A, const float* __restrict__ B,
                                                 const float* __restrict__ Bimg, int N, int K, int mode,
                                                 int lbo_units, float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 128 x 32 fp32 = 16 KB
  uint8_t* sB = smem + 16384;         // 256 x 32 fp32 = 32 KB
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_init(smem_u32(&bar_tma), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(128, N);
  const int nkb = K / 32;
  bool ok = true;
  for (int kb = 0; kb < nkb && ok; ++kb) {
    // A tile by the threads (the epilogue's job in the real kernel)
    for (int i = tid; i < 128 * 32; i += 128) {
      int r = i >> 5, k = i & 31;
      *reinterpret_cast<float*>(sA + sw128_offset(r, k)) = A[r * K + kb * 32 + k];
    }
    if (mode & 1) {
      if (tid == 0) {
        mbar_arrive_expect_tx(smem_u32(&bar_tma), N * 128);
        bulk_g2s(smem_u32(sB), Bimg + (size_t)kb * N * 32, N * 128, smem_u32(&bar_tma));
      }
      ok = mbar_wait(smem_u32(&bar_tma), kb & 1);
    } else {
      for (int i = tid; i < N * 32; i += 128) {
        int r = i >> 5, k = i & 31;
        *reinterpret_cast<float*>(sB + sw128_offset(r, k)) = B[r * K + kb * 32 + k];
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t ad = umma_desc_sw128(smem_u32(sA) + ks * 32, lbo_units);
        uint64_t bd = umma_desc_sw128(smem_u32(sB) + ks * 32, lbo_units);
        umma_tf32(tmem_base, ad, bd, idesc, (kb | ks) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar_mma));
    }
    ok = ok && mbar_wait(smem_u32(&bar_mma), kb & 1);   // smem tiles are reused next iteration
    tc_fence_after();
    __syncthreads();
  }
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      int r = warp * 32 + (tid & 31);
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[r * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// D[128, N] = sum_p A[p, 0:128] * X[p, 0:N] over P points (a multiple of 32): both operands MN-major, the layout of
// the weight-gradient GEMM.  Tiles are [32 points x 32 features] SWIZZLE_128B images, feature chunks 4 KB apart.
__global__ void __launch_bounds__(128) probe_gemm_mn(const float* __restrict__ A, const float* __restrict__ X, int N,
                                                    int P, int sbo, float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;            // 4 chunks x 4 KB
  uint8_t* sX = smem + 16384;    // 8 chunks x 4 KB
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(&bar_mma), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32_mn(128, N);
  bool ok = true;
  for (int st = 0; st < P / 32 && ok; ++st) {
    for (int i = tid; i < 32 * 128; i += 128) {
      int p = i / 128, n = i % 128;
      *reinterpret_cast<float*>(sA + (n >> 5) * 4096 + sw128_32b_offset(p, n & 31)) = A[(st * 32 + p) * 128 + n];
    }
    for (int i = tid; i < 32 * N; i += 128) {
      int p = i / N, n = i % N;
      *reinterpret_cast<float*>(sX + (n >> 5) * 4096 + sw128_32b_offset(p, n & 31)) = X[(st * 32 + p) * N + n];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int g = 0; g < 4; ++g)
        umma_tf32(tmem_base, umma_desc_sw128_mn(smem_u32(sA) + g * 1024, 4096, sbo),
                  umma_desc_sw128_mn(smem_u32(sX) + g * 1024, 4096, sbo), idesc, (st | g) ? 1u : 0u);
      umma_commit(smem_u32(&bar_mma));
    }
    ok = mbar_wait(smem_u32(&bar_mma), st & 1);
    tc_fence_after();
    __syncthreads();
  }
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      int r = warp * 32 + (tid & 31);
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[r * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static int run_gemm_mn(int N, int P, int sbo = 512) {
  std::vector<float> A((size_t)P * 128), X((size_t)P * N), D(128 * N, -1.f), R(128 * N, 0.f);
  srand(99 + N + P);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : X) v = (float)(rand() % 7 - 3);
  for (int p = 0; p < P; ++p)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) R[m * N + n] += A[p * 128 + m] * X[p * N + n];
  float *dA, *dX, *dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4)); CK(cudaMemset(dS, 0, 4));
  const int smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe_gemm_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm_mn<<<1, 128, smem>>>(dA, dX, N, P, sbo, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_gemm_mn N=%d P=%d: LAUNCH FAILED %s\n", N, P, cudaGetErrorString(e)); return 2; }
  int st = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0, first = -1;
  double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    double er = fabs((double)D[i] - R[i]);
    if (!(er <= 1e-3)) { if (first < 0) first = (int)i; ++bad; }
    if (er > maxerr) maxerr = er;
  }
  printf("probe_gemm_mn sbo=%d N=%d P=%d: %s (timeout=%d, mismatches=%d/%zu, maxerr=%g", sbo, N, P, (bad == 0 && st == 0) ? "PASS" : "FAIL",
         st, bad, D.size(), maxerr);
  if (first >= 0) printf(", first bad (m=%d,n=%d) got %g want %g", first / N, first % N, D[first], R[first]);
  printf(")\n");
  if (bad) {
    for (int m = 0; m < 4; ++m) {
      printf("  row %d got:", m);
      for (int n = 0; n < 8; ++n) printf(" %6.0f", D[m * N + n]);
      printf("   want:");
      for (int n = 0; n < 8; ++n) printf(" %6.0f", R[m * N + n]);
      printf("\n");
    }
    for (int m : {32, 64}) {
      printf("  row %d got:", m);
      for (int n : {0, 1, 32, 33, 64, 128}) if (n < N) printf(" %6.0f", D[m * N + n]);
      printf("   want:");
      for (int n : {0, 1, 32, 33, 64, 128}) if (n < N) printf(" %6.0f", R[m * N + n]);
      printf("\n");
    }
  }
  cudaFree(dA); cudaFree(dX); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0 ? 0 : 1;
}

// D[128, N] = A[128, K] * B[N, K]^T with the A operand in TENSOR MEMORY (written with tcgen05.st), B in shared memory.
__global__ void __launch_bounds__(128) probe_gemm_ts(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                                                    float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;   // 256 x 32 fp32 = 32 KB per K block
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(smem_u32(&bar_mma), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tA = tmem_base + 256;   // A lives in columns [256, 256 + K)
  // every thread owns one row of A: write it to TMEM, 32 columns at a time
  for (int c0 = 0; c0 < K; c0 += 32) {
    float v[32];
    for (int j = 0; j < 32; ++j) v[j] = A[(warp * 32 + lane) * K + c0 + j];
    tmem_st32(tA + ((uint32_t)(warp * 32) << 16) + c0, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  const uint32_t idesc = umma_idesc_tf32(128, N);
  bool ok = true;
  for (int kb = 0; kb < K / 32 && ok; ++kb) {
    for (int i = tid; i < N * 32; i += 128) {
      int r = i >> 5, k = i & 31;
      *reinterpret_cast<float*>(sB + sw128_offset(r, k)) = B[r * K + kb * 32 + k];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int ks = 0; ks < 4; ++ks)
        umma_tf32_ts(tmem_base, tA + kb * 32 + ks * 8, umma_desc_sw128(smem_u32(sB) + ks * 32), idesc, (kb | ks) ? 1u : 0u);
      umma_commit(smem_u32(&bar_mma));
    }
    ok = mbar_wait(smem_u32(&bar_mma), kb & 1);
    tc_fence_after();
    __syncthreads();
  }
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      int r = warp * 32 + lane;
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[r * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static int run_gemm_ts(int N, int K) {
  std::vector<float> A(128 * K), B(256 * K), D(128 * N, -1.f), R(128 * N);
  srand(4321 + N + K);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : B) v = (float)(rand() % 7 - 3);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4)); CK(cudaMemset(dS, 0, 4));
  const int smem = 32768 + 1024;
  probe_gemm_ts<<<1, 128, smem>>>(dA, dB, N, K, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_gemm_ts N=%d K=%d: LAUNCH FAILED %s\n", N, K, cudaGetErrorString(e)); return 2; }
  int st = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0, first = -1;
  double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    double er = fabs((double)D[i] - R[i]);
    if (!(er <= 1e-3)) { if (first < 0) first = (int)i; ++bad; }
    if (er > maxerr) maxerr = er;
  }
  printf("probe_gemm_ts (A in TMEM) N=%d K=%d: %s (timeout=%d, mismatches=%d/%zu, maxerr=%g", N, K,
         (bad == 0 && st == 0) ? "PASS" : "FAIL", st, bad, D.size(), maxerr);
  if (first >= 0) printf(", first bad (m=%d,n=%d) got %g want %g", first / N, first % N, D[first], R[first]);
  printf(")\n");
  if (bad)
    for (int m : {0, 1, 32, 33}) {
      printf("  row %d got:", m);
      for (int n = 0; n < 6; ++n) printf(" %6.0f", D[m * N + n]);
      printf("   want:");
      for (int n = 0; n < 6; ++n) printf(" %6.0f", R[m * N + n]);
      printf("\n");
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0 ? 0 : 1;
}

// Issue `iters` x 4 back-to-back MMAs (N=256, K=8 each) on fixed operands and report cycles per MMA.
__global__ void __launch_bounds__(128) probe_rate(int iters, long long* cycles, int* status, int mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (tid == 0) {
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = mn ? umma_idesc_tf32_mn(128, 256) : umma_idesc_tf32(128, 256);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int ks = 0; ks < 4; ++ks) {
        if (mn)   // both operands MN-major (SWIZZLE_128B_BASE32B tiles of 32 points x 32 features): the wgrad layout
          umma_tf32(tmem_base + (it & 1) * 256, umma_desc_sw128_mn(smem_u32(smem) + ks * 1024, 4096, 512),
                    umma_desc_sw128_mn(smem_u32(smem + 16384) + ks * 1024, 4096, 512), idesc, 1u);
        else
          umma_tf32(tmem_base + (it & 1) * 256, umma_desc_sw128(smem_u32(smem) + ks * 32),
                    umma_desc_sw128(smem_u32(smem + 16384) + ks * 32), idesc, 1u);
      }
    umma_commit(smem_u32(&bar_mma));
    bool ok = mbar_wait(smem_u32(&bar_mma), 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    if (!ok) status[0] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static int run_gemm(int N, int K, int mode, int lbo) {
  std::vector<float> A(128 * K), B(256 * K), Bimg((size_t)(K / 32) * N * 32), D(128 * N, -1.f), R(128 * N);
  srand(1234 + N + K);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : B) v = (float)(rand() % 7 - 3);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  for (int kb = 0; kb < K / 32; ++kb)
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 32; ++k)
        Bimg[(size_t)kb * N * 32 + sw128_offset(n, k) / 4] = B[n * K + kb * 32 + k];
  float *dA, *dB, *dBi, *dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dBi, Bimg.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBi, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, D.size() * 4)); CK(cudaMemset(dS, 0, 4));
  const int smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm<<<1, 128, smem>>>(dA, dB, dBi, N, K, mode, lbo, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("probe_gemm N=%d K=%d mode=%d lbo=%d: LAUNCH FAILED %s\n", N, K, mode, lbo, cudaGetErrorString(e));
    return 2;
  }
  int st = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0, first = -1;
  for (size_t i = 0; i < D.size(); ++i) {
    double er = fabs((double)D[i] - R[i]);
    if (!(er <= 1e-3)) { if (first < 0) first = (int)i; ++bad; }
    if (er > maxerr) maxerr = er;
  }
  printf("probe_gemm N=%d K=%d mode=%d lbo=%d: %s (timeout=%d, mismatches=%d/%zu, maxerr=%g", N, K, mode, lbo,
         (bad == 0 && st == 0) ? "PASS" : "FAIL", st, bad, D.size(), maxerr);
  if (first >= 0) printf(", first bad (m=%d,n=%d) got %g want %g", first / N, first % N, D[first], R[first]);
  printf(")\n");
  if (bad) {  // dump a corner to help infer a layout error
    for (int m = 0; m < 4; ++m) {
      printf("  row %d got:", m);
      for (int n = 0; n < 8; ++n) printf(" %6.0f", D[m * N + n]);
      printf("   want:");
      for (int n = 0; n < 8; ++n) printf(" %6.0f", R[m * N + n]);
      printf("\n");
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dBi); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0 ? 0 : 1;
}

int main() {
  int fails = 0;
  fails += run_gemm(256, 32, 0, 1) != 0;
  fails += run_gemm(256, 64, 0, 1) != 0;
  fails += run_gemm(256, 256, 0, 1) != 0;
  fails += run_gemm(256, 64, 0, 0) != 0;
  fails += run_gemm(224, 64, 0, 1) != 0;
  fails += run_gemm(16, 64, 0, 1) != 0;
  fails += run_gemm(256, 64, 1, 1) != 0;
  fails += run_gemm(224, 256, 1, 1) != 0;
  fails += run_gemm_ts(256, 32) != 0;
  fails += run_gemm_ts(256, 256) != 0;
  fails += run_gemm_ts(16, 64) != 0;
  fails += run_gemm_mn(256, 32) != 0;
  fails += run_gemm_mn(256, 128) != 0;
  fails += run_gemm_mn(64, 64) != 0;
  fails += run_gemm_mn(16, 64) != 0;
  {
    long long* dC;
    int* dS;
    CK(cudaMalloc(&dC, 148 * 8)); CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4));
    const int smem = 16384 + 32768 + 1024;
    CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mn = 0; mn < 2; ++mn)
    for (int grid : {1, 148}) {
      const int iters = 2048;
      probe_rate<<<grid, 128, smem>>>(iters, dC, dS, mn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("probe_rate: LAUNCH FAILED %s\n", cudaGetErrorString(e)); ++fails; break; }
      long long c[148];
      CK(cudaMemcpy(c, dC, grid * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < grid; ++i) if (c[i] > mx) mx = c[i];
      printf("probe_rate grid=%d %s: %.1f cycles per 128x256x8 tf32 MMA (max over CTAs)\n", grid,
             mn ? "MN-major operands" : "K-major operands", (double)mx / (iters * 4));
    }
  }
  printf("probe_tc: %s (%d failing)\n", fails ? "FAIL" : "ALL PASS", fails);
  return fails ? 1 : 0;
}

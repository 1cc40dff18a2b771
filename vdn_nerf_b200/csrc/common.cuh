// Shared device helpers for the vdn_nerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>

#define VDN_CHECK(expr)                                   \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return (int)_e;                \
  } while (0)

#define VDN_LAUNCH_CHECK()                                \
  do {                                                    \
    const int _e = ::vdn::take_launch_error();            \
    if (_e) return _e;                                    \
  } while (0)

namespace vdn {

extern std::atomic<long long> g_launches;  // defined in api.cu; read through vdn_launch_count()

// Launch errors are collected per thread by VDN_LAUNCH itself and fetched with take_launch_error(): the thread's CUDA
// "last error" may hold a stale, harmless error left by another library of the process, which must not be reported
// as a failure of this one.
inline thread_local int g_launch_error = 0;
inline int take_launch_error() {
  const int e = g_launch_error;
  g_launch_error = 0;
  return e;
}

// Every kernel launch of the library goes through this macro so the launch count is exact.
#define VDN_LAUNCH(kernel, grid, block, smem, stream, ...)                                     \
  do {                                                                                         \
    ++::vdn::g_launches;                                                                       \
    (void)cudaGetLastError();                                                                  \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                \
    const cudaError_t _le = cudaGetLastError();                                                \
    if (_le != cudaSuccess && !::vdn::g_launch_error) ::vdn::g_launch_error = (int)_le;        \
  } while (0)

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kSoftplusBeta = 100.0f;
constexpr float kSoftplusThreshold = 20.0f;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

// softplus(beta=100, threshold=20) as nn.Softplus computes it (reference fields.py:70).
__device__ __forceinline__ float softplus100(float z) {
  float t = z * kSoftplusBeta;
  return t > kSoftplusThreshold ? z : log1pf(expf(t)) * (1.0f / kSoftplusBeta);
}
// d softplus / dz  (= sigmoid(100 z) below the threshold, 1 above it)
__device__ __forceinline__ float softplus100_d1(float z) {
  float t = z * kSoftplusBeta;
  return t > kSoftplusThreshold ? 1.0f : 1.0f / (1.0f + expf(-t));
}
// d^2 softplus / dz^2 (= 100 s (1-s) below the threshold, 0 above it)
__device__ __forceinline__ float softplus100_d2(float z) {
  float t = z * kSoftplusBeta;
  if (t > kSoftplusThreshold) return 0.0f;
  float s = 1.0f / (1.0f + expf(-t));
  return kSoftplusBeta * s * (1.0f - s);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// MUFU-based, branch-free variants for the tensor-core mode, whose operands are rounded to tf32 (2^-11) anyway:
// absolute error of softplus ~1e-8, relative error of its derivatives ~1e-6.  They are written with raw
// ex2/lg2/rcp.approx.ftz and folded constants (8 / 4 / 7 instructions): per 128x32 operand block the tensor pipe
// needs 512 cycles, i.e. 2048 issue slots of the SM - at the ~20 instructions per element of __expf/__logf the
// element-wise work, not the MMA, bounds the kernel.
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float kBetaLog2e = 144.26950408889634f;          // 100 * log2(e)
constexpr float kThrLog2e = 28.853900817779268f;           // 20 * log2(e)
constexpr float kLn2OverBeta = 0.0069314718055994531f;     // ln(2) / 100
__device__ __forceinline__ float softplus100_fast(float z) {
  // softplus(z) >= z everywhere and the clamped branch never exceeds z above the clamp, so max() selects exactly
  // like the reference's threshold (7 instructions)
  const float s = lg2_approx(1.0f + ex2_approx(fminf(z * kBetaLog2e, kThrLog2e))) * kLn2OverBeta;
  return fmaxf(z, s);
}
// sigmoid(100 z); above the threshold 1/(1+e^-20) already rounds to 1.0f, matching the reference's branch
__device__ __forceinline__ float softplus100_d1_fast(float z) { return rcp_approx(1.0f + ex2_approx(-z * kBetaLog2e)); }
__device__ __forceinline__ float softplus100_d2_fast(float z) {
  const float s = rcp_approx(1.0f + ex2_approx(-z * kBetaLog2e));
  return kSoftplusBeta * s * (1.0f - s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vdn

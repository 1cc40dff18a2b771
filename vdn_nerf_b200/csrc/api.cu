// Library-level entry points and the stand-alone embedder kernels' launchers.
#include "pointwise.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {
std::atomic<long long> g_launches{0};
}
using namespace vdn;

extern "C" int vdn_abi_version(void) { return 1; }
extern "C" long long vdn_launch_count(void) { return g_launches.load(); }
extern "C" const char* vdn_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

extern "C" int vdn_embed_fwd(const float* x, long long N, int d, int multires, float* out, void* stream) {
  if (N <= 0) return 0;
  if (d < 1 || d > 8 || multires < 0 || multires > 16) return (int)cudaErrorInvalidValue;
  const int d_e = d * (1 + 2 * multires);
  VDN_LAUNCH(embed_rows_kernel, (unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream, x, d, N, d, multires, 1.0f, out, d_e,
                                                                                 nullptr, 0, 0, 1.0f, 0);
  return (int)cudaGetLastError();
}

extern "C" int vdn_embed_bwd(const float* x, long long N, int d, int multires, const float* d_out, float* d_x,
                             void* stream) {
  if (N <= 0) return 0;
  if (d < 1 || d > 8 || multires < 0 || multires > 16) return (int)cudaErrorInvalidValue;
  const int d_e = d * (1 + 2 * multires);
  long long tot = N * d;
  VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream, x, d, N, d, multires, 1.0f, d_out,
                                                                                  d_e, nullptr, 0, 0.0f, 1.0f, d_x, d,
                                                                                  0);
  return (int)cudaGetLastError();
}

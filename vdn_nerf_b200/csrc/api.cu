// Library-level entry points and the stand-alone embedder kernels' launchers.
#include "pointwise.cuh"
#include "../../include/vdn_b200.h"

#include <vector>
#include "gemm_tc.cuh"

namespace vdn {
std::atomic<long long> g_launches{0};
int g_mode = 0;
int g_chain = 1;
int* g_tc_fault = nullptr;
long long* g_tc_dbg = nullptr;

// ---- optional profiling: CUDA events around kernel families, summed on read ----------------------------
static bool g_prof_on = false;
struct ProfSpan { cudaEvent_t a, b; };
static std::vector<ProfSpan> g_spans[PROF_FAMILIES];
static double g_flops[PROF_FAMILIES];
static double g_bytes[PROF_FAMILIES];
static cudaEvent_t g_open[PROF_FAMILIES];

void prof_begin(int family, cudaStream_t st, double flops, double bytes) {
  if (!g_prof_on) return;
  cudaEvent_t a;
  cudaEventCreate(&a);
  cudaEventRecord(a, st);
  g_open[family] = a;
  g_flops[family] += flops;
  g_bytes[family] += bytes;
}
void prof_end(int family, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, st);
  g_spans[family].push_back({g_open[family], b});
}
}  // namespace vdn
using namespace vdn;

extern "C" int vdn_abi_version(void) { return 5; }
extern "C" long long vdn_launch_count(void) { return g_launches.load(); }
extern "C" const char* vdn_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

extern "C" int vdn_set_mode(int mode) {
  if (mode != 0 && mode != 1) return (int)cudaErrorInvalidValue;
  g_mode = mode;
  return 0;
}
// The 4-byte device flag a tcgen05 kernel raises when one of its bounded barrier waits times out.  The caller owns the
// memory (a torch tensor on the device the kernels run on); null switches the reporting off.
extern "C" int vdn_set_fault_flag(int* device_flag) {
  g_tc_fault = device_flag;
  return 0;
}
// Enqueue a device -> host copy of the flag on `stream` (host_dst should be pinned); no synchronisation.
extern "C" int vdn_tc_fault_async(int* host_dst, void* stream) {
  if (!g_tc_fault || !host_dst) return 0;
  return (int)cudaMemcpyAsync(host_dst, g_tc_fault, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
extern "C" int vdn_get_mode(void) { return g_mode; }
extern "C" int vdn_set_chain(int on) { g_chain = on ? 1 : 0; return 0; }
extern "C" int vdn_get_chain(void) { return g_chain; }
extern "C" int vdn_tc_fault(void) {
  if (!g_tc_fault) return 0;
  int h = 0;
  if (cudaMemcpy(&h, g_tc_fault, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return h;
}

extern "C" int vdn_debug_timeline(long long* device_buf) {
  g_tc_dbg = device_buf;
  return 0;
}

extern "C" int vdn_prof_enable(int on) {
  g_prof_on = on != 0;
  for (int f = 0; f < PROF_FAMILIES; ++f) {
    for (auto& s : g_spans[f]) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_spans[f].clear();
    g_flops[f] = 0.0;
    g_bytes[f] = 0.0;
  }
  return 0;
}

extern "C" int vdn_prof_read(int family, double* ms, long long* spans, double* flops) {
  if (family < 0 || family >= PROF_FAMILIES) return (int)cudaErrorInvalidValue;
  double total = 0.0;
  for (auto& s : g_spans[family]) {
    cudaError_t e = cudaEventSynchronize(s.b);
    if (e != cudaSuccess) return (int)e;
    float t = 0.f;
    cudaEventElapsedTime(&t, s.a, s.b);
    total += t;
  }
  *ms = total;
  *spans = (long long)g_spans[family].size();
  *flops = g_flops[family];
  return 0;
}

extern "C" int vdn_prof_read_bytes(int family, double* bytes) {
  if (family < 0 || family >= PROF_FAMILIES) return (int)cudaErrorInvalidValue;
  *bytes = g_bytes[family];
  return 0;
}

extern "C" int vdn_embed_fwd(const float* x, long long N, int d, int multires, float* out, void* stream) {
  if (N <= 0) return 0;
  if (d < 1 || d > 8 || multires < 0 || multires > 16) return (int)cudaErrorInvalidValue;
  const int d_e = d * (1 + 2 * multires);
  return launch_embed_rows(x, d, N, d, multires, 1.0f, out, d_e, nullptr, 0, 0, 1.0f, 0, (cudaStream_t)stream);
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
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// Weight materialisation: one launch per network and optimiser step turns the raw parameters
// (weight_v / weight_g / bias of old-style nn.utils.weight_norm, reference fields.py:65-66, 141-142; plain
// weight / bias for the NeRF background field) into the packed layout of mlp_layout.cuh - fp32 W and W^T for the
// FFMA kernels plus SWIZZLE_128B tile images of both (tf32-rounded for the layer-wise tcgen05 kernels, fp16 for the
// chain engine) - and one launch turns
// a packed gradient back into per-parameter gradients (weight-norm backward).
#include "mlp_layout.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include "../../include/vdn_b200.h"

namespace vdn {

struct PackLayer {
  const float* v[2];
  const float* g[2];
  const float* b[2];
  int rows[2];
  int in_dim, in_ld, out_dim, out_ld, rot, orot;
  long long off_w, off_wt, off_b, off_iw, off_iwt, off_ih, off_iht;
};
struct PackArgs {
  int L;
  int row_start[VDN_MAX_LAYERS + 1];
  PackLayer layer[VDN_MAX_LAYERS];
};

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// One warp per (padded) output row.  The tile-image regions are zero-filled by the caller beforehand.
__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a, float* __restrict__ packed) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.row_start[a.L]) return;
  int l = 0;
  while (warp >= a.row_start[l + 1]) ++l;
  const int r = warp - a.row_start[l];
  const PackLayer& P = a.layer[l];
  float* W = packed + P.off_w + (long long)r * P.in_ld;
  float* WT = packed + P.off_wt;
  float* B = packed + P.off_b;
  int src = -1, rr = r;
  if (r < P.rows[0]) src = 0;
  else if (r < P.rows[0] + P.rows[1]) { src = 1; rr = r - P.rows[0]; }
  if (src < 0) {
    for (int k = lane; k < P.in_ld; k += 32) { W[k] = 0.0f; WT[(long long)k * P.out_ld + r] = 0.0f; }
    if (lane == 0) B[r] = 0.0f;
    return;
  }
  const float* v = P.v[src] + (long long)rr * P.in_dim;
  float sc = 1.0f;
  if (P.g[src]) {
    float ss = 0.0f;
    for (int k = lane; k < P.in_dim; k += 32) ss += v[k] * v[k];
    ss = warp_sum(ss);
    sc = P.g[src][rr] / sqrtf(ss);
  }
  float* IW = packed + P.off_iw;
  float* IWT = packed + P.off_iwt;
  __half* IH = reinterpret_cast<__half*>(packed + P.off_ih);
  __half* IHT = reinterpret_cast<__half*>(packed + P.off_iht);
  int rj = r - P.orot;                      // position of output r in the (rotated) fp16 images
  if (rj < 0) rj += P.out_dim;
  for (int k = lane; k < P.in_ld; k += 32) {
    float w = 0.0f;
    if (k < P.in_dim) {
      int c = k + P.rot;
      if (c >= P.in_dim) c -= P.in_dim;
      w = v[c] * sc;
    }
    W[k] = w;
    WT[(long long)k * P.out_ld + r] = w;
    if (k < P.in_dim) {
      const float t = round_tf32(w);
      IW[(long long)(k >> 5) * P.out_ld * 32 + (tc::sw128_offset((uint32_t)r, (uint32_t)(k & 31)) >> 2)] = t;
      IWT[(long long)(r >> 5) * P.in_ld * 32 + (tc::sw128_offset((uint32_t)k, (uint32_t)(r & 31)) >> 2)] = t;
      const __half hw = __float2half_rn(w);
      const long long i_w = (long long)(k >> 6) * P.out_ld * 64 + (tc::sw128_offset_h((uint32_t)rj, (uint32_t)(k & 63)) >> 1);
      const long long i_t = (long long)(rj >> 6) * P.in_ld * 64 + (tc::sw128_offset_h((uint32_t)k, (uint32_t)(rj & 63)) >> 1);
      IH[i_w] = hw;
      IHT[i_t] = hw;
    }
  }
  if (lane == 0) B[r] = P.b[src] ? P.b[src][rr] : 0.0f;
}

struct UnpackLayer {
  const float* v[2];
  const float* g[2];
  float* dv[2];
  float* dg[2];
  float* db[2];
  int rows[2];
  int in_dim, in_ld, rot;
  long long off_w, off_b;
};
struct UnpackArgs {
  int L;
  int row_start[VDN_MAX_LAYERS + 1];
  UnpackLayer layer[VDN_MAX_LAYERS];
};

__global__ void unpack_grads_kernel(const __grid_constant__ UnpackArgs a, const float* __restrict__ dpacked) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.row_start[a.L]) return;
  int l = 0;
  while (warp >= a.row_start[l + 1]) ++l;
  const int r = warp - a.row_start[l];
  const UnpackLayer& P = a.layer[l];
  int src = 0, rr = r;
  if (r >= P.rows[0]) { src = 1; rr = r - P.rows[0]; }
  const float* dWrow = dpacked + P.off_w + (long long)r * P.in_ld;
  auto dW = [&](int c) {  // gradient w.r.t. source column c (undo the column rotation)
    int k = c - P.rot;
    if (k < 0) k += P.in_dim;
    return dWrow[k];
  };
  const float* v = P.v[src] + (long long)rr * P.in_dim;
  float* dv = P.dv[src] ? P.dv[src] + (long long)rr * P.in_dim : nullptr;
  if (P.g[src]) {
    float dot = 0.0f, ss = 0.0f;
    for (int k = lane; k < P.in_dim; k += 32) { float vv = v[k]; dot += dW(k) * vv; ss += vv * vv; }
    dot = warp_sum(dot);
    ss = warp_sum(ss);
    const float norm = sqrtf(ss);
    const float gn = P.g[src][rr] / norm;
    if (dv) for (int k = lane; k < P.in_dim; k += 32) dv[k] = gn * (dW(k) - v[k] * (dot / ss));
    if (lane == 0 && P.dg[src]) P.dg[src][rr] = dot / norm;
  } else if (dv) {
    for (int k = lane; k < P.in_dim; k += 32) dv[k] = dW(k);
  }
  if (lane == 0 && P.db[src]) P.db[src][rr] = dpacked[P.off_b + r];
}

}  // namespace vdn
using namespace vdn;

extern "C" long long vdn_mlp_layout(int L, const int* in_dims, const int* out_dims, long long* off_w,
                                    long long* off_wt, long long* off_b) {
  MlpLayout ly;
  if (make_layout(L, in_dims, out_dims, &ly)) return -1;
  for (int l = 0; l < L; ++l) {
    if (off_w) off_w[l] = ly.off_w[l];
    if (off_wt) off_wt[l] = ly.off_wt[l];
    if (off_b) off_b[l] = ly.off_b[l];
  }
  return ly.total;
}

extern "C" int vdn_mlp_pack(int L, const int* in_dims, const int* out_dims, const float* const* v,
                            const float* const* g, const float* const* b, const int* rows, const int* rot,
                            const int* orot, float* packed, void* stream) {
  MlpLayout ly;
  if (make_layout(L, in_dims, out_dims, &ly)) return (int)cudaErrorInvalidValue;
  PackArgs a;
  a.L = L;
  int start = 0;
  for (int l = 0; l < L; ++l) {
    a.row_start[l] = start;
    start += ly.out_ld[l];
    PackLayer& P = a.layer[l];
    for (int s = 0; s < 2; ++s) {
      P.v[s] = v[2 * l + s]; P.g[s] = g[2 * l + s]; P.b[s] = b[2 * l + s]; P.rows[s] = rows[2 * l + s];
    }
    if (P.rows[0] + P.rows[1] != out_dims[l] || !P.v[0] || (P.rows[1] > 0 && !P.v[1]))
      return (int)cudaErrorInvalidValue;
    P.in_dim = ly.in_dim[l]; P.in_ld = ly.in_ld[l]; P.out_dim = ly.out_dim[l]; P.out_ld = ly.out_ld[l];
    P.rot = rot ? rot[l] : 0;
    if (P.rot < 0 || P.rot >= P.in_dim) return (int)cudaErrorInvalidValue;
    P.orot = orot ? orot[l] : 0;
    if (P.orot < 0 || P.orot >= P.out_dim) return (int)cudaErrorInvalidValue;
    P.off_w = ly.off_w[l]; P.off_wt = ly.off_wt[l]; P.off_b = ly.off_b[l];
    P.off_iw = ly.off_iw[l]; P.off_iwt = ly.off_iwt[l]; P.off_ih = ly.off_ih[l]; P.off_iht = ly.off_iht[l];
  }
  a.row_start[L] = start;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(packed + ly.off_iw[0], 0, (size_t)(ly.total - ly.off_iw[0]) * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const int threads = 256;
  const int blocks = (start * 32 + threads - 1) / threads;
  VDN_LAUNCH(pack_weights_kernel, blocks, threads, 0, st, a, packed);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_mlp_unpack_grads(int L, const int* in_dims, const int* out_dims, const float* const* v,
                                    const float* const* g, const int* rows, const int* rot, const float* dpacked,
                                    float* const* dv, float* const* dg, float* const* db, void* stream) {
  MlpLayout ly;
  if (make_layout(L, in_dims, out_dims, &ly)) return (int)cudaErrorInvalidValue;
  UnpackArgs a;
  a.L = L;
  int start = 0;
  for (int l = 0; l < L; ++l) {
    a.row_start[l] = start;
    start += ly.out_dim[l];
    UnpackLayer& P = a.layer[l];
    for (int s = 0; s < 2; ++s) {
      P.v[s] = v[2 * l + s]; P.g[s] = g[2 * l + s]; P.rows[s] = rows[2 * l + s];
      P.dv[s] = dv[2 * l + s]; P.dg[s] = dg[2 * l + s]; P.db[s] = db[2 * l + s];
    }
    if (P.rows[0] + P.rows[1] != out_dims[l]) return (int)cudaErrorInvalidValue;
    P.in_dim = ly.in_dim[l]; P.in_ld = ly.in_ld[l];
    P.rot = rot ? rot[l] : 0;
    P.off_w = ly.off_w[l]; P.off_b = ly.off_b[l];
  }
  a.row_start[L] = start;
  const int threads = 256;
  const int blocks = (start * 32 + threads - 1) / threads;
  VDN_LAUNCH(unpack_grads_kernel, blocks, threads, 0, (cudaStream_t)stream, a, dpacked);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

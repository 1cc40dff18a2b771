// Marching cubes on the device-resident field of extract_fields: the reference hands the 512^3
// field to the third-party host library `mcubes` (dpt_models/renderer.py:33-41) after a 537 MB device -> host copy.
// Two kernels around an exclusive scan (torch.cumsum on the caller's side): count the triangles of every cell, then emit
// them as (grid-edge key, interpolated position) triples; the caller welds vertices by key (torch.unique).  The case table
// is generated on the host (vdn_nerf_b200/mcubes_table.py) and passed in as device arrays.
#include "common.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {

__device__ __forceinline__ int mc_case(const float* __restrict__ u, int ny, int nz, int i, int j, int k, float thr) {
  int m = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float v = u[((size_t)(i + (c & 1)) * ny + (j + ((c >> 1) & 1))) * nz + (k + ((c >> 2) & 1))];
    m |= (v > thr ? 1 : 0) << c;
  }
  return m;
}

__global__ void mc_count_kernel(const float* __restrict__ u, int nx, int ny, int nz, float thr,
                                const int* __restrict__ tri_count, int* __restrict__ counts) {
  const long long cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  const int k = (int)(idx % (nz - 1));
  const long long t = idx / (nz - 1);
  const int j = (int)(t % (ny - 1)), i = (int)(t / (ny - 1));
  counts[idx] = tri_count[mc_case(u, ny, nz, i, j, k, thr)];
}

// offsets = exclusive prefix sum of counts (triangles before this cell).  keys[3 t + v] = 3 * (linear index of the lower
// end point of the cut grid edge) + axis; pos[3 t + v] = interpolated vertex in grid-index coordinates.
__global__ void mc_emit_kernel(const float* __restrict__ u, int nx, int ny, int nz, float thr, const int* __restrict__ tri_count,
                               const int* __restrict__ tri_table, const int* __restrict__ edge_corner,
                               const int* __restrict__ edge_axis, const long long* __restrict__ offsets,
                               long long* __restrict__ keys, float* __restrict__ pos) {
  const long long cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells) return;
  const int k = (int)(idx % (nz - 1));
  const long long t = idx / (nz - 1);
  const int j = (int)(t % (ny - 1)), i = (int)(t / (ny - 1));
  const int m = mc_case(u, ny, nz, i, j, k, thr);
  const int n = tri_count[m];
  if (n == 0) return;
  long long o = offsets[idx] * 3;
  for (int v = 0; v < 3 * n; ++v, ++o) {
    const int e = tri_table[m * 16 + v];
    const int a = edge_corner[e], axis = edge_axis[e];
    const int x0 = i + (a & 1), y0 = j + ((a >> 1) & 1), z0 = k + ((a >> 2) & 1);
    const int x1 = x0 + (axis == 0), y1 = y0 + (axis == 1), z1 = z0 + (axis == 2);
    const float u0 = u[((size_t)x0 * ny + y0) * nz + z0], u1 = u[((size_t)x1 * ny + y1) * nz + z1];
    const float tt = (thr - u0) / (u1 - u0);
    keys[o] = (((long long)x0 * ny + y0) * nz + z0) * 3 + axis;
    pos[o * 3 + 0] = (float)x0 + (axis == 0 ? tt : 0.0f);
    pos[o * 3 + 1] = (float)y0 + (axis == 1 ? tt : 0.0f);
    pos[o * 3 + 2] = (float)z0 + (axis == 2 ? tt : 0.0f);
  }
}

}  // namespace vdn
using namespace vdn;

extern "C" int vdn_mc_count(const float* u, int nx, int ny, int nz, float threshold, const int* tri_count, int* counts,
                            void* stream) {
  if (nx < 2 || ny < 2 || nz < 2) return (int)cudaErrorInvalidValue;
  const long long cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  VDN_LAUNCH(mc_count_kernel, (unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream, u, nx, ny, nz, threshold, tri_count,
             counts);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_mc_emit(const float* u, int nx, int ny, int nz, float threshold, const int* tri_count, const int* tri_table,
                           const int* edge_corner, const int* edge_axis, const long long* offsets, long long* keys, float* pos,
                           void* stream) {
  if (nx < 2 || ny < 2 || nz < 2) return (int)cudaErrorInvalidValue;
  const long long cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  VDN_LAUNCH(mc_emit_kernel, (unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream, u, nx, ny, nz, threshold, tri_count,
             tri_table, edge_corner, edge_axis, offsets, keys, pos);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

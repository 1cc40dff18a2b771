// Callers on either side of the rendering path ("next" rows N1 and N2), as kernels:
//
//   vdn_adam_step        torch.optim.Adam's update (dpt_runner.py:88, 251-253) for ALL parameter tensors of the path in
//                        ONE launch (the reference's step is ~12 small ATen kernels per tensor group)
//   vdn_color_loss       the driver's masked L1 colour loss + PSNR sums and its gradient (dpt_runner.py:228-232)
//   vdn_raygen_fwd/bwd   pixel -> world ray for one camera (dpt_models/poses.py:189-212: K^-1 p, normalise, rotate by
//                        the camera-to-world pose) and the cotangent of that pose, so that rays never leave the device
//                        (the reference assembles them on the CPU and copies them back every step)
#include "common.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {

constexpr int ADAM_MAX_TENSORS = 96;
struct AdamArgs {
  int n_tensors;
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  int n[ADAM_MAX_TENSORS];
};

// hyper (device): {lr, beta1, beta2, eps, 1 - beta1, 1 - beta2 (rounded from double like torch's scalars),
// (1 - beta1^t_i, 1 - beta2^t_i) for every tensor i} - torch keeps one step
// counter per parameter (a parameter without gradient skips the step); device resident so that a captured CUDA graph
// follows the learning-rate schedule.
__global__ void adam_step_kernel(const __grid_constant__ AdamArgs a, const float* __restrict__ hyper) {
  const int t = blockIdx.y;
  const int n = a.n[t];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], omb1 = hyper[4], omb2 = hyper[5];
  const float bc1 = hyper[6 + 2 * t], bc2 = hyper[7 + 2 * t];
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= step_size * (mi / denom);
  }
}

// sums[0] += sum |c - t| * mask, sums[1] += sum ((c - t) * mask)^2, sums[2] += sum mask;  d_color = sign(c - t) * mask
// (the caller divides by mask_sum).  One thread per ray.
__global__ void color_loss_kernel(const float* __restrict__ color, const float* __restrict__ rgb,
                                  const float* __restrict__ mask, long long B, float* __restrict__ sums,
                                  float* __restrict__ d_color) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float l1 = 0.f, l2 = 0.f, ms = 0.f;
  if (b < B) {
    const float mk = mask ? mask[b] : 1.0f;
    ms = mk;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float e = (color[b * 3 + j] - rgb[b * 3 + j]) * mk;
      l1 += fabsf(e);
      l2 += e * e;
      d_color[b * 3 + j] = (e > 0.f ? 1.0f : (e < 0.f ? -1.0f : 0.0f)) * mk;
    }
  }
  l1 = warp_sum(l1); l2 = warp_sum(l2); ms = warp_sum(ms);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums + 0, l1);
    atomicAdd(sums + 1, l2);
    atomicAdd(sums + 2, ms);
  }
}

// rays_d[b] = R * normalize(Kinv * [px, py, 1]),  rays_o[b] = t,  with pose = [R | t] (3 x 4 row-major, device)
__global__ void raygen_fwd_kernel(const float* __restrict__ px, const float* __restrict__ py, long long B,
                                  const float* __restrict__ kinv, const float* __restrict__ pose, float* __restrict__ rays_o,
                                  float* __restrict__ rays_d) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float x = px[b], y = py[b];
  float p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = kinv[i * 3 + 0] * x + kinv[i * 3 + 1] * y + kinv[i * 3 + 2];
  const float inv = 1.0f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] *= inv;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rays_d[b * 3 + i] = pose[i * 4 + 0] * p[0] + pose[i * 4 + 1] * p[1] + pose[i * 4 + 2] * p[2];
    rays_o[b * 3 + i] = pose[i * 4 + 3];
  }
}
// d_pose[i][j] += sum_b d_rays_d[b][i] * v_cam[b][j] (j < 3),  d_pose[i][3] += sum_b d_rays_o[b][i]
__global__ void raygen_bwd_kernel(const float* __restrict__ px, const float* __restrict__ py, long long B,
                                  const float* __restrict__ kinv, const float* __restrict__ d_rays_o,
                                  const float* __restrict__ d_rays_d, float* __restrict__ d_pose) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  if (b < B) {
    const float x = px[b], y = py[b];
    float p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = kinv[i * 3 + 0] * x + kinv[i * 3 + 1] * y + kinv[i * 3 + 2];
    const float inv = 1.0f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float gd = d_rays_d ? d_rays_d[b * 3 + i] : 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[i * 4 + j] = gd * p[j] * inv;
      acc[i * 4 + 3] = d_rays_o ? d_rays_o[b * 3 + i] : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float s = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(d_pose + i, s);
  }
}

}  // namespace vdn
using namespace vdn;

extern "C" int vdn_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int* numel, const float* hyper_dev, void* stream) {
  if (n_tensors <= 0) return 0;
  if (n_tensors > ADAM_MAX_TENSORS || !hyper_dev) return (int)cudaErrorInvalidValue;
  AdamArgs a;
  a.n_tensors = n_tensors;
  int maxn = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0 || (numel[i] > 0 && (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i])))
      return (int)cudaErrorInvalidValue;      // numel 0: tensor skipped this step (no gradient)
    a.p[i] = params[i]; a.g[i] = grads[i]; a.m[i] = exp_avg[i]; a.v[i] = exp_avg_sq[i]; a.n[i] = numel[i];
    if (numel[i] > maxn) maxn = numel[i];
  }
  int bx = (maxn + 1023) / 1024;          // 256 threads x 4 elements per pass
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  VDN_LAUNCH(adam_step_kernel, dim3(bx, n_tensors), 256, 0, (cudaStream_t)stream, a, hyper_dev);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_color_loss(const float* color, const float* true_rgb, const float* mask, long long B, float* sums,
                              float* d_color, void* stream) {
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, 3 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  VDN_LAUNCH(color_loss_kernel, (unsigned)((B + 255) / 256), 256, 0, st, color, true_rgb, mask, B, sums, d_color);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_raygen_fwd(const float* px, const float* py, long long B, const float* kinv, const float* pose,
                              float* rays_o, float* rays_d, void* stream) {
  if (B <= 0) return 0;
  VDN_LAUNCH(raygen_fwd_kernel, (unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream, px, py, B, kinv, pose, rays_o, rays_d);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_raygen_bwd(const float* px, const float* py, long long B, const float* kinv, const float* d_rays_o,
                              const float* d_rays_d, float* d_pose, void* stream) {
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_pose, 0, 12 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  VDN_LAUNCH(raygen_bwd_kernel, (unsigned)((B + 255) / 256), 256, 0, st, px, py, B, kinv, d_rays_o, d_rays_d, d_pose);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

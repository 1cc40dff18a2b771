// HBM read throughput of a [M x 256] fp32 matrix under the access orders the GEMM operand producers can use.
//   mode 0: per CTA a 128-row tile, walked K block by K block (128 B per row and step)   - the layer-wise GEMM today
//   mode 1: same tile, walked row by row (1 KB contiguous per row)
//   mode 2: same tile, 4 K blocks (512 B per row) per step
//   mode 3: flat grid-stride copy-like read (reference for the achievable peak)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_hbm tools/probe_hbm.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) read_tiles(const float4* __restrict__ a, long long M, int mode, int inflight,
                                                   float* __restrict__ sink) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc = 0.f;
  for (long long tile = blockIdx.x; tile < M / 128; tile += gridDim.x) {
    const float4* base = a + tile * 128 * 64;   // 64 float4 per row
    if (mode == 0) {
      // K block kb: thread covers rows warp*16 + (lane>>3) + 4i, chunk lane&7
      for (int kb = 0; kb < 8; kb += inflight) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < inflight) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j * 4 + i] = base[(size_t)(warp * 16 + (lane >> 3) + 4 * i) * 64 + (kb + j) * 8 + (lane & 7)];
          }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < inflight) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc += v[j * 4 + i].x + v[j * 4 + i].w;
          }
      }
    } else if (mode == 1) {
      // row by row: warp w rows w*16..+15, a warp instruction reads 512 B contiguous
      for (int r = 0; r < 16; r += 4) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[2 * i] = base[(size_t)(warp * 16 + r + i) * 64 + lane];
          v[2 * i + 1] = base[(size_t)(warp * 16 + r + i) * 64 + 32 + lane];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].w;
      }
    } else if (mode == 2) {
      for (int kq = 0; kq < 2; ++kq) {
        for (int r = 0; r < 16; r += 8) {
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = base[(size_t)(warp * 16 + r + i) * 64 + kq * 32 + lane];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].w;
        }
      }
    }
  }
  if (mode == 3) {
    const long long n4 = M * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n4; i += (long long)gridDim.x * blockDim.x * 4) {
      float4 v0 = a[i], v1 = i + (long long)gridDim.x * blockDim.x < n4 ? a[i + (long long)gridDim.x * blockDim.x] : v0;
      float4 v2 = i + 2LL * gridDim.x * blockDim.x < n4 ? a[i + 2LL * gridDim.x * blockDim.x] : v0;
      float4 v3 = i + 3LL * gridDim.x * blockDim.x < n4 ? a[i + 3LL * gridDim.x * blockDim.x] : v0;
      acc += v0.x + v1.x + v2.x + v3.x;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const long long M = 1 << 20;   // 1 GiB matrix (>> L2)
  float4* a; float* sink;
  cudaMalloc(&a, M * 1024); cudaMalloc(&sink, 4);
  cudaMemset(a, 0, M * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { int mode, inflight, ctas_per_sm; const char* name; } cfg[] = {
    {0, 1, 2, "K-block walk, 1 K block in flight per thread, 2 CTA/SM"},
    {0, 2, 2, "K-block walk, 2 K blocks at once, 2 CTA/SM"},
    {0, 4, 2, "K-block walk, 4 K blocks at once, 2 CTA/SM"},
    {0, 1, 4, "K-block walk, 1 K block, 4 CTA/SM"},
    {0, 1, 8, "K-block walk, 1 K block, 8 CTA/SM"},
    {2, 0, 2, "512 B per row per step, 2 CTA/SM"},
    {1, 0, 2, "row by row (1 KB contiguous), 2 CTA/SM"},
    {1, 0, 4, "row by row, 4 CTA/SM"},
    {3, 0, 8, "flat grid-stride read, 8 CTA/SM"},
  };
  cudaFuncSetAttribute(read_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int pass = 0; pass < 2; ++pass)
  for (auto& c : cfg) {
    const int grid = 148 * c.ctas_per_sm;
    // pass 1: the same with the GEMM's shared-memory footprint (198 KB per SM), which shrinks L1 to ~30 KB
    const size_t smem = pass ? (size_t)(198 * 1024 / c.ctas_per_sm) : 0;
    if (pass && c.ctas_per_sm > 2) continue;
    if (pass) printf("[%zu KB dynamic smem per CTA] ", smem / 1024);
    read_tiles<<<grid, 256, smem>>>(a, M, c.mode, c.inflight, sink);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) read_tiles<<<grid, 256, smem>>>(a, M, c.mode, c.inflight, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-60s %8.1f GB/s\n", c.name, 3.0 * M * 1024 / (ms * 1e-3) / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

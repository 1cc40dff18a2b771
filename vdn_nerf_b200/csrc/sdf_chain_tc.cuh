// Fused SDF value chain on tcgen05 (tensor-core mode): positional encoding + all softplus(beta=100) layers of the
// SDF network for a tile of 128 points in ONE kernel - activations never leave the SM.
//
//   reference: SDFNetwork.sdf (dpt_models/fields.py:72-92) as used by the hierarchical sampler
//   (renderer.py:369-370, 201) and by extract_fields (renderer.py:10-30, 446).
//
// Per CTA (persistent, one per SM, tiles of 128 points):
//   * the activation tile A [128 x 256] lives in TENSOR MEMORY (columns 256..511; lane = point, one column per
//     feature) and is the A operand of the MMAs - with tf32 a 128x256x8 MMA would otherwise read 4 KB of A plus 8 KB
//     of B from shared memory every 128 cycles, i.e. 75 % of the shared-memory port, before any staging traffic;
//   * weight tiles ([n x 32] pre-swizzled tf32 images, mlp_layout.cuh) stream through a 6-stage ring (192 KB),
//     fetched by one thread with cp.async.bulk (TMA engine) running ahead across layers and tiles;
//   * one thread issues tcgen05.mma kind::tf32 into the 256-column accumulator D (TMEM columns 0..255);
//   * sixteen warps run the epilogue of layer l: drain D into registers (tcgen05.ld) and release it, then - all warps
//     on the same 32-column chunk, chunk after chunk - +bias -> softplus -> tf32 -> tcgen05.st into A, signalling each
//     finished K-block on its own mbarrier, so the MMAs of layer l+1 trail the epilogue of layer l by one chunk;
//   * the skip connection cat[h, e]/sqrt2 (fields.py:82-83) is formed in the epilogue of the preceding layer, the
//     embedding e is recomputed from the point (no extra buffer);
//   * in grid mode the lattice point is generated from its index (no point tensor in HBM).
// Supported shape: d_in = 3, d_hidden = 256, d_e <= 64, one skip layer; anything else takes the layer-wise path.
#pragma once
#include "gemm_tc.cuh"

namespace vdn {

constexpr int CH_EPI_WARPS = 16;    // four warps per scheduler: the epilogue is latency bound with fewer
constexpr int CH_THREADS = (CH_EPI_WARPS + 2) * 32;   // + warp 16: TMEM alloc + MMA issue; warp 17: weight stream
constexpr int CH_WSTAGES = 6;
constexpr uint32_t CH_W_STAGE = 32768;
constexpr size_t CH_SMEM = CH_WSTAGES * CH_W_STAGE + 1024;

struct ChainLayer {
  long long img_off, bias_off;   // float offsets into the packed buffer
  int out_ld, out_dim, n_mma, nkb;
};
struct ChainArgs {
  int L, skip, d_e, multires;
  float scale, out_mul;
  const float* packed;
  const float* x;                // [N,3] points, or null in grid mode
  const float *xs, *ys, *zs;     // grid mode: coordinate vectors
  int ny, nz, i0;
  long long N;
  float* out;
  int lds;
  ChainLayer layer[VDN_MAX_LAYERS];
};

// embedding column c (< d_e) of point y: [y | sin(2^k y) | cos(2^k y)]_k, d = 3 (embedder.py:15-36)
static __device__ __noinline__ float chain_embed_col(float y0, float y1, float y2, int c) {
  const float y[3] = {y0, y1, y2};
  if (c < 3) return y[c];
  const int k = (c - 3) / 6, rem = (c - 3) - 6 * k;
  const float f = (float)(1 << k);
  return rem < 3 ? sinf(y[rem] * f) : cosf(y[rem - 3] * f);
}

// One output element of a chunk that is not entirely real outputs: the tail of the layer before the skip connection
// carries the embedding (fields.py:82-83), anything beyond is zero padding.  Cold path.
static __device__ __noinline__ float chain_ragged_elem(float acc, int nn, int out_dim, const float* bias, float osc,
                                                       int d_e_tail, float y0, float y1, float y2) {
  if (nn < out_dim) return softplus100_fast(acc + bias[nn]) * osc;
  if (nn - out_dim < d_e_tail) return chain_embed_col(y0, y1, y2, nn - out_dim) * osc;
  return 0.0f;
}

static __global__ void __launch_bounds__(CH_THREADS, 1)
sdf_chain_tc_kernel(const __grid_constant__ ChainArgs a, int* __restrict__ fault) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[CH_WSTAGES], w_empty[CH_WSTAGES], a_ready[8], d_full, d_drained;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sW = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const long long ntiles = (a.N + 127) / 128;

  if (tid == 0) {
    for (int s = 0; s < CH_WSTAGES; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int j = 0; j < 8; ++j) mbar_init(smem_u32(&a_ready[j]), CH_EPI_WARPS * 32);
    mbar_init(smem_u32(&d_full), 1);
    mbar_init(smem_u32(&d_drained), CH_EPI_WARPS * 32);
    mbar_fence_init();
  }
  if (warp == CH_EPI_WARPS) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;   // accumulator D: columns [0,256); activation tile A: columns [256,512)
  bool ok = true;

  if (warp < CH_EPI_WARPS) {
    // ================= embedding + epilogue warps =================
    // warp (q, h): TMEM lane quarter q (rows 32q..32q+31), columns [32 ch + 8 h, +8) of every 32-column chunk ch.
    // All sixteen warps work on the same chunk, chunk after chunk, so K-block ch of the next layer's operand is
    // complete after 1/8 of the epilogue and the tensor pipe trails the epilogue by one chunk.
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t tD = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 8);
    const uint32_t tA = tD + 256u;
    uint32_t dcnt = 0;                                // completions of d_full consumed
    for (long long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
      const long long m = tile * 128 + row;
      const bool valid = m < a.N;
      float y[3] = {0.f, 0.f, 0.f};
      if (valid) {
        if (a.x) {
          y[0] = a.x[m * 3] * a.scale; y[1] = a.x[m * 3 + 1] * a.scale; y[2] = a.x[m * 3 + 2] * a.scale;
        } else {
          const int k = (int)(m % a.nz);
          const long long t = m / a.nz;
          const int j = (int)(t % a.ny);
          const int i = a.i0 + (int)(t / a.ny);
          y[0] = a.xs[i] * a.scale; y[1] = a.ys[j] * a.scale; y[2] = a.zs[k] * a.scale;
        }
      }
      // ---- positional encoding -> K-blocks 0 and 1 of A (activation tile in TMEM) ----
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = ch * 32 + h * 8 + j;
          v[j] = c < a.d_e ? to_tf32(chain_embed_col(y[0], y[1], y[2], c)) : 0.0f;
        }
        tmem_st8(tA + (uint32_t)(ch * 32), v);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&a_ready[ch]));
      }
      // ---- layer epilogues ----
      for (int l = 0; l < a.L && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        ok = mbar_wait(smem_u32(&d_full), dcnt & 1);
        ++dcnt;
        tc_fence_after();
        const float* bias = a.packed + Ly.bias_off;
        if (l == a.L - 1) {
          if (h == 0) {
            float v[8];
            tmem_ld8(tD, v);
            tmem_ld_wait();
            if (valid) a.out[m * a.lds] = (v[0] + bias[0]) * (a.out_mul / a.scale);
          }
          tc_fence_before();
          mbar_arrive(smem_u32(&d_drained));
          continue;
        }
        // drain this thread's 8 x 8 accumulator columns into registers, then release D so the MMAs of the next layer
        // may overwrite it while the activations are still being computed
        float v[8][8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) tmem_ld8(tD + (uint32_t)(ch * 32), v[ch]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&d_drained));
        const bool skip_next = (l + 1 == a.skip);
        const float osc = skip_next ? kInvSqrt2 : 1.0f;
        const int d_e_tail = skip_next ? a.d_e : 0;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int n0 = ch * 32 + h * 8;
          if (ch * 32 + 32 <= Ly.out_dim) {      // chunk of real outputs (uniform over the CTA)
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0) + 1);
            v[ch][0] = to_tf32(softplus100_fast(v[ch][0] + b0.x) * osc);
            v[ch][1] = to_tf32(softplus100_fast(v[ch][1] + b0.y) * osc);
            v[ch][2] = to_tf32(softplus100_fast(v[ch][2] + b0.z) * osc);
            v[ch][3] = to_tf32(softplus100_fast(v[ch][3] + b0.w) * osc);
            v[ch][4] = to_tf32(softplus100_fast(v[ch][4] + b1.x) * osc);
            v[ch][5] = to_tf32(softplus100_fast(v[ch][5] + b1.y) * osc);
            v[ch][6] = to_tf32(softplus100_fast(v[ch][6] + b1.z) * osc);
            v[ch][7] = to_tf32(softplus100_fast(v[ch][7] + b1.w) * osc);
          } else {                                // tail of the layer before the skip connection / zero padding
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[ch][j] = to_tf32(chain_ragged_elem(n0 + j < Ly.n_mma ? v[ch][j] : 0.0f, n0 + j, Ly.out_dim, bias, osc,
                                                   d_e_tail, y[0], y[1], y[2]));
          }
          tmem_st8(tA + (uint32_t)(ch * 32), v[ch]);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(&a_ready[ch]));
        }
      }
    }
  } else if (tid == CH_EPI_WARPS * 32) {
    // ================= MMA issuer: A from tensor memory, weights from shared memory =================
    uint32_t acnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t wt = 0, drained = 0;
    const uint32_t tAcol = tmem_base + 256u;
    for (long long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
      for (int l = 0; l < a.L && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)Ly.n_mma);
        // the accumulator of the previous layer must have been drained (first layer ever: passes immediately)
        ok = mbar_wait(smem_u32(&d_drained), (drained & 1) ^ 1);
        ++drained;
        for (int kb = 0; kb < Ly.nkb && ok; ++kb, ++wt) {
          const uint32_t ws = wt % CH_WSTAGES, wph = (wt / CH_WSTAGES) & 1;
          ok = mbar_wait(smem_u32(&w_full[ws]), wph);
          ok = ok && mbar_wait(smem_u32(&a_ready[kb]), acnt[kb] & 1);
          ++acnt[kb];
          tc_fence_after();
          const uint32_t b0 = sW + ws * CH_W_STAGE;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_ts(tmem_base, tAcol + (uint32_t)(kb * 32 + ks * 8), umma_desc_sw128(b0 + ks * 32), idesc,
                         (kb | ks) ? 1u : 0u);
          umma_commit(smem_u32(&w_empty[ws]));
        }
        umma_commit(smem_u32(&d_full));
      }
    }
  } else if (tid == (CH_EPI_WARPS + 1) * 32) {
    // ================= weight stream (TMA engine) =================
    uint32_t wt = 0;
    for (long long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
      for (int l = 0; l < a.L && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const uint32_t bytes = (uint32_t)Ly.n_mma * 128u;
        for (int kb = 0; kb < Ly.nkb && ok; ++kb, ++wt) {
          const uint32_t ws = wt % CH_WSTAGES, wph = (wt / CH_WSTAGES) & 1;
          ok = mbar_wait(smem_u32(&w_empty[ws]), wph ^ 1);
          mbar_arrive_expect_tx(smem_u32(&w_full[ws]), bytes);
          bulk_g2s(sW + ws * CH_W_STAGE, a.packed + Ly.img_off + (size_t)kb * Ly.out_ld * 32, bytes, smem_u32(&w_full[ws]));
        }
      }
    }
  }
  if (!ok && fault) *fault = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == CH_EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

// Returns -1 when the configuration is not supported by the fused chain (caller falls back to the layer-wise path).
static inline int launch_sdf_chain(const MlpLayout& ly, int d_in, int multires, int d_hidden, int skip, float scale,
                                   const float* packed, const float* x, const float* xs, const float* ys,
                                   const float* zs, int ny, int nz, int i0, long long N, float* out, int lds, float out_mul,
                                   cudaStream_t st) {
  const int d_e = d_in * (1 + 2 * multires);
  if (d_in != 3 || d_hidden != 256 || d_e > 64 || ly.L < 2 || ly.L > VDN_MAX_LAYERS) return -1;
  for (int l = 1; l < ly.L; ++l)
    if (ly.in_dim[l] != 256) return -1;
  ChainArgs a;
  a.L = ly.L; a.skip = skip; a.d_e = d_e; a.multires = multires; a.scale = scale; a.out_mul = out_mul;
  a.packed = packed; a.x = x; a.xs = xs; a.ys = ys; a.zs = zs; a.ny = ny; a.nz = nz; a.i0 = i0; a.N = N; a.out = out; a.lds = lds;
  for (int l = 0; l < ly.L; ++l) {
    ChainLayer& c = a.layer[l];
    c.img_off = ly.off_iw[l]; c.bias_off = ly.off_b[l]; c.out_ld = ly.out_ld[l]; c.out_dim = ly.out_dim[l];
    c.n_mma = (l == ly.L - 1) ? 16 : ((ly.out_dim[l] + 15) & ~15);
    c.nkb = (ly.in_dim[l] + 31) / 32;
    if (c.n_mma > 256 || c.nkb > 8) return -1;
  }
  static int num_sms = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sdf_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long ntiles = (N + 127) / 128;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  double flops = 0.0;
  for (int l = 0; l < ly.L; ++l) flops += 2.0 * (double)N * a.layer[l].n_mma * a.layer[l].nkb * 32;
  prof_begin(PROF_TC, st, flops);
  VDN_LAUNCH(sdf_chain_tc_kernel, grid, CH_THREADS, CH_SMEM, st, a, g_tc_fault);
  prof_end(PROF_TC, st);
  return (int)cudaGetLastError();
}

}  // namespace vdn

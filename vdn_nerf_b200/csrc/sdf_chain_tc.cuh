// Fused SDF value chain on tcgen05 (tensor-core mode): positional encoding + all softplus(beta=100) layers of the
// SDF network for a tile of 128 points in ONE kernel - activations never leave the SM.
//
//   reference: SDFNetwork.sdf (dpt_models/fields.py:72-92) as used by the hierarchical sampler
//   (renderer.py:369-370, 201) and by extract_fields (renderer.py:10-30, 446).
//
// Per CTA (persistent, one per SM, tiles of 128 points):
//   * the activation tile A [128 x 256] lives in shared memory as eight SWIZZLE_128B K-blocks (128 KB);
//   * weight tiles ([n x 32] pre-swizzled tf32 images, mlp_layout.cuh) stream through a 3-stage ring (96 KB),
//     fetched by one thread with cp.async.bulk (TMA engine) running ahead across layers and tiles;
//   * one thread issues tcgen05.mma kind::tf32; the two 256-column TMEM accumulators alternate between layers;
//   * eight warps run the epilogue of layer l (TMEM -> +bias -> softplus -> tf32 -> swizzled store into A) chunk by
//     chunk (32 columns) and signal each finished K-block on its own mbarrier, so the MMAs of layer l+1 start
//     while the epilogue of layer l is still running (the tensor pipe trails the epilogue by one chunk);
//   * the skip connection cat[h, e]/sqrt2 (fields.py:82-83) is formed in the epilogue of the preceding layer, the
//     embedding e is recomputed from the point (no extra buffer);
//   * in grid mode the lattice point is generated from its index (no point tensor in HBM).
// Supported shape: d_in = 3, d_hidden = 256, d_e <= 64, one skip layer; anything else takes the layer-wise path.
#pragma once
#include "gemm_tc.cuh"

namespace vdn {

constexpr int CH_THREADS = 320;     // warps 0-7: embedding + epilogue; warp 8: TMEM alloc + MMA issue; warp 9: weight stream
constexpr int CH_WSTAGES = 3;
constexpr uint32_t CH_A_BYTES = 8 * 16384;
constexpr uint32_t CH_W_STAGE = 32768;
constexpr size_t CH_SMEM = CH_A_BYTES + CH_WSTAGES * CH_W_STAGE + 1024;

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

// Chunk of 32 output columns that is not entirely real outputs: the tail of the layer before the skip connection
// (remaining columns carry the embedding, fields.py:82-83) or zero padding.  Cold path, kept out of line.
static __device__ __noinline__ void chain_ragged_chunk(uint32_t tacc, int n0, int n_mma, int out_dim, const float* bias,
                                                       float osc, int d_e_tail, float y0, float y1, float y2,
                                                       uint32_t abase, uint32_t r7) {
  using namespace tc;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.0f;
  if (n0 < n_mma) {
    tmem_ld32(tacc + (uint32_t)n0, v);
    tmem_ld_wait();
  }
  for (int c4 = 0; c4 < 8; ++c4) {
    float o[4];
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + c4 * 4 + j;
      float r = 0.0f;
      if (nn < out_dim) {
        float acc = 0.0f;
#pragma unroll
        for (int t = 0; t < 32; ++t) acc = (t == c4 * 4 + j) ? v[t] : acc;
        r = softplus100_fast(acc + bias[nn]) * osc;
      } else if (nn - out_dim < d_e_tail) {
        r = chain_embed_col(y0, y1, y2, nn - out_dim) * osc;
      }
      o[j] = to_tf32(r);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(abase + (((uint32_t)c4 ^ r7) << 4)), "f"(o[0]), "f"(o[1]),
                 "f"(o[2]), "f"(o[3])
                 : "memory");
  }
}

static __global__ void __launch_bounds__(CH_THREADS, 1)
sdf_chain_tc_kernel(const __grid_constant__ ChainArgs a, int* __restrict__ fault) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[CH_WSTAGES], w_empty[CH_WSTAGES], a_ready[8], d_full[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem0, sW = smem0 + CH_A_BYTES;
  const long long ntiles = (a.N + 127) / 128;

  if (tid == 0) {
    for (int s = 0; s < CH_WSTAGES; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int j = 0; j < 8; ++j) mbar_init(smem_u32(&a_ready[j]), 128);
    mbar_init(smem_u32(&d_full[0]), 1);
    mbar_init(smem_u32(&d_full[1]), 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  if (warp < 8) {
    // ================= embedding + epilogue warps =================
    const int q = warp & 3, h = warp >> 2;           // TMEM lane quarter, column half
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t r7 = (uint32_t)(row & 7);
    uint32_t dcnt0 = 0, dcnt1 = 0;                    // completions consumed of d_full[0], d_full[1]
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
      // ---- positional encoding -> K-block h of A (columns 32h .. 32h+31) ----
      {
        const uint32_t base = sA + (uint32_t)h * 16384u + rowoff;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = h * 32 + c4 * 4 + j;
            v[j] = c < a.d_e ? to_tf32(chain_embed_col(y[0], y[1], y[2], c)) : 0.0f;
          }
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((uint32_t)c4 ^ r7) << 4)), "f"(v[0]),
                       "f"(v[1]), "f"(v[2]), "f"(v[3])
                       : "memory");
        }
        fence_proxy_async();
        mbar_arrive(smem_u32(&a_ready[h]));
      }
      // ---- layer epilogues ----
      for (int l = 0; l < a.L && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const int acc = l & 1;
        uint32_t& dc = acc ? dcnt1 : dcnt0;
        ok = mbar_wait(smem_u32(&d_full[acc]), dc & 1);
        ++dc;
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16);
        const float* bias = a.packed + Ly.bias_off;
        if (l == a.L - 1) {
          if (h == 0) {
            float v[32];
            tmem_ld32(tacc, v);
            tmem_ld_wait();
            if (valid) a.out[m * a.lds] = (v[0] + bias[0]) * (a.out_mul / a.scale);
          }
          tc_fence_before();
          continue;
        }
        const bool skip_next = (l + 1 == a.skip);
        const float osc = skip_next ? kInvSqrt2 : 1.0f;
        for (int ch = h; ch < 8; ch += 2) {
          const int n0 = ch * 32;
          const uint32_t abase = sA + (uint32_t)ch * 16384u + rowoff;
          if (n0 + 32 <= Ly.out_dim) {
            // hot path: a full chunk of real outputs; bias loads are issued before waiting on the TMEM load
            float v[32];
            tmem_ld32(tacc + (uint32_t)n0, v);
            float4 b[8];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) b[c4] = __ldg(reinterpret_cast<const float4*>(bias + n0) + c4);
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float o0 = softplus100_fast(v[c4 * 4 + 0] + b[c4].x) * osc;
              const float o1 = softplus100_fast(v[c4 * 4 + 1] + b[c4].y) * osc;
              const float o2 = softplus100_fast(v[c4 * 4 + 2] + b[c4].z) * osc;
              const float o3 = softplus100_fast(v[c4 * 4 + 3] + b[c4].w) * osc;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(abase + (((uint32_t)c4 ^ r7) << 4)),
                           "f"(to_tf32(o0)), "f"(to_tf32(o1)), "f"(to_tf32(o2)), "f"(to_tf32(o3))
                           : "memory");
            }
          } else {
            chain_ragged_chunk(tacc, n0, Ly.n_mma, Ly.out_dim, bias, osc, skip_next ? a.d_e : 0, y[0], y[1], y[2], abase, r7);
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(smem_u32(&a_ready[ch]));
        }
      }
    }
  } else if (tid == 8 * 32) {
    // ================= MMA issuer =================
    uint32_t acnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t wt = 0;   // weight tiles consumed
    for (long long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
      for (int l = 0; l < a.L && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)Ly.n_mma);
        const uint32_t dacc = tmem_base + (uint32_t)(l & 1) * 256u;
        for (int kb = 0; kb < Ly.nkb && ok; ++kb, ++wt) {
          const uint32_t ws = wt % CH_WSTAGES, wph = (wt / CH_WSTAGES) & 1;
          ok = mbar_wait(smem_u32(&w_full[ws]), wph);
          ok = ok && mbar_wait(smem_u32(&a_ready[kb]), acnt[kb] & 1);
          ++acnt[kb];
          tc_fence_after();
          const uint32_t a0 = sA + (uint32_t)kb * 16384u, b0 = sW + ws * CH_W_STAGE;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(dacc, umma_desc_sw128(a0 + ks * 32), umma_desc_sw128(b0 + ks * 32), idesc, (kb | ks) ? 1u : 0u);
          umma_commit(smem_u32(&w_empty[ws]));
        }
        umma_commit(smem_u32(&d_full[l & 1]));
      }
    }
  } else if (tid == 9 * 32) {
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
  if (warp == 8) tmem_dealloc(tmem_base, 512);
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

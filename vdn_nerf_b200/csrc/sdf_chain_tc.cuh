// Fused SDF value chain on tcgen05 (tensor-core mode): positional encoding + all softplus(beta=100) layers of the
// SDF network for a tile of 128 points in ONE kernel - activations never leave the SM.
//
//   reference: SDFNetwork.sdf (dpt_models/fields.py:72-92) as used by the hierarchical sampler
//   (renderer.py:369-370, 201) and by extract_fields (renderer.py:10-30, 446).
//
// Operand precision: fp16 activations and weights (10-bit mantissa, the same as tf32; every magnitude on this path
// sits well inside the fp16 range), exact products, fp32 accumulation: kind::f16 runs at twice the kind::tf32 rate,
// halves the weight bytes per MMA and packs the activation tile into half the tensor-memory columns.
//
// The chain runs in "base-2 softplus units": with t = z * beta/ln2 the activation is a' = log2(1 + 2^t) =
// max(t,0) + log2(1 + 2^-|t|) and the next layer's t' = W a' + (beta/ln2) b - the beta scaling cancels, so an element
// costs one FFMA (bias), one MUFU.EX2, a degree-4 polynomial for log2(1+w)/w on [0,1] (|err| < 3.3e-5 in a', i.e.
// 2e-7 in softplus units, an order of magnitude below the fp16 rounding of a'), FMNMX, FFMA and half a pack
// instruction.  The embedding enters scaled by beta/ln2, the last layer scales back.
//
// Per CTA (persistent, one per SM), TWO tiles of 128 points (slots X and Y) are in flight and alternate phases:
// while the tensor pipe runs layer l of one slot, the epilogue warps turn the other slot's accumulator into its
// next operand, so neither the MMAs nor the element-wise work wait on each other's latencies.
//   * the activation tiles A_X, A_Y [128 x 256] fp16 live in TENSOR MEMORY (columns 256..383 and 384..511; lane =
//     point, two features per column) and are the A operands of the MMAs (no shared-memory traffic for A);
//   * one 256-column fp32 accumulator D (TMEM columns 0..255) is shared by both slots: the epilogue warps drain it
//     into registers (tcgen05.ld) as soon as a phase completes and release it for the other slot's MMAs;
//   * weight tiles ([n x 64] pre-swizzled fp16 images, mlp_layout.cuh) stream through a 6-stage ring (192 KB),
//     fetched by one thread with cp.async.bulk (TMA engine); the first two tiles of a layer serve both slots, the
//     others are fetched once per slot so the ring can run ahead into the next layer;
//   * one thread issues tcgen05.mma kind::f16; sixteen warps run the epilogue (bias from shared memory, activation,
//     fp16x2 pack, tcgen05.st into A_slot);
//   * the skip connection cat[h, e]/sqrt2 (fields.py:82-83): the layer before it writes [a' | (beta/ln2) e] and the
//     1/sqrt2 is applied to the skip layer's accumulator (same FFMA as the bias), the embedding e is recomputed from
//     the point (no extra buffer);
//   * in grid mode the lattice point is generated from its index (no point tensor in HBM).
// Supported shape: d_in = 3, d_hidden = 256, d_e <= 64, one skip layer; anything else takes the layer-wise path.
#pragma once
#include "gemm_tc.cuh"
#include <cuda_fp16.h>

namespace vdn {

constexpr int CH_EPI_WARPS = 16;    // four warps per scheduler: the epilogue is latency bound with fewer
constexpr int CH_THREADS = (CH_EPI_WARPS + 2) * 32;   // + warp 16: TMEM alloc + MMA issue; warp 17: weight stream
constexpr int CH_WSTAGES = 6;
constexpr int CH_KEEP = 4;         // weight tiles of a layer kept in the ring for the second slot
constexpr uint32_t CH_W_STAGE = 32768;
constexpr size_t CH_SMEM = CH_WSTAGES * CH_W_STAGE + 1024 + VDN_MAX_LAYERS * 256 * sizeof(float) + 64 * sizeof(float4);

struct ChainLayer {
  long long img_off, bias_off;   // float offsets into the packed buffer
  int out_ld, out_dim, n_mma, nkb;
  int row0;                      // first image row of the B tile (the sdf row of a rotated last layer)
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
  // training forward: Lrun < L layers are run (the multi-head last layer is left to the layer-wise kernel) and the
  // pre-activations z_l (fields.py:79-87, in softplus units) are stored for the analytic gradient passes
  int Lrun;
  float* save_z[VDN_MAX_LAYERS];   // [N, ldz] per layer or null
  float* save_u;                   // head of the skip layer's input, softplus(z_{skip-1}) / sqrt2, or null
  int ldz;
  ChainLayer layer[VDN_MAX_LAYERS];
};

// embedding column c (< d_e) of point y: [y | sin(2^k y) | cos(2^k y)]_k, d = 3 (embedder.py:15-36), through a
// per-CTA table {frequency, phase, coordinate, identity flag} (cos x = sin(x + pi/2)).  MUFU sin: |2^k y| stays below
// ~100 rad on this path, where sin.approx is good to ~1e-5 absolute - far below the fp16 rounding of the operand.
__device__ __forceinline__ float chain_embed_col(const float4* lut, float y0, float y1, float y2, int c) {
  const float4 e = lut[c];
  const float y = e.z == 0.0f ? y0 : (e.z == 1.0f ? y1 : y2);
  return e.w != 0.0f ? y : __sinf(fmaf(y, e.x, e.y));
}

constexpr float kB2 = 144.26950408889634f;     // beta / ln 2 for beta = 100 (fields.py:50 Softplus(beta=100))
constexpr float kInvB2 = 1.0f / 144.26950408889634f;

// a' = log2(1 + 2^t) = max(t, 0) + w q(w), w = 2^-|t|; q = degree-4 fit of log2(1+w)/w on [0,1], |w q - log2(1+w)| < 3.3e-5
__device__ __forceinline__ float softplus_base2(float t) {
  float w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(-fabsf(t)));
  float q = fmaf(0.04008112847805023f, w, -0.1803952157497406f);
  q = fmaf(q, w, 0.4036492109298706f);
  q = fmaf(q, w, -0.7047332525253296f);
  q = fmaf(q, w, 1.4414016008377075f);
  return fmaf(w, q, fmaxf(t, 0.0f));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One output element of a chunk that is not entirely real outputs: the tail of the layer before the skip connection
// carries the embedding (fields.py:82-83), anything beyond is zero padding.  t = pre-activation in base-2 units (bias
// included).
__device__ __forceinline__ float chain_ragged_elem(const float4* lut, float t, int nn, int out_dim, int d_e_tail, float y0,
                                                   float y1, float y2) {
  if (nn < out_dim) return softplus_base2(t);
  if (nn - out_dim < d_e_tail) return chain_embed_col(lut, y0, y1, y2, nn - out_dim) * kB2;
  return 0.0f;
}

// Hot path: 8 consecutive real outputs of one row -> 4 packed fp16x2 words.  sb = bias * beta/ln2 in shared memory,
// dsc = scale of the accumulator (1/sqrt2 on the skip layer, else 1).  SAVE: also store the pre-activations (zrow) and,
// on the layer before the skip connection, the scaled activations (urow), both in softplus units.
template <bool SAVE>
__device__ __forceinline__ void chain_epi8(const float (&v)[8], const float* sb, float dsc, uint32_t (&p)[4], float* zrow,
                                           float* urow) {
  const float4 b0 = *reinterpret_cast<const float4*>(sb);
  const float4 b1 = *reinterpret_cast<const float4*>(sb + 4);
  float t[8], r[8];
  t[0] = fmaf(v[0], dsc, b0.x); t[1] = fmaf(v[1], dsc, b0.y); t[2] = fmaf(v[2], dsc, b0.z); t[3] = fmaf(v[3], dsc, b0.w);
  t[4] = fmaf(v[4], dsc, b1.x); t[5] = fmaf(v[5], dsc, b1.y); t[6] = fmaf(v[6], dsc, b1.z); t[7] = fmaf(v[7], dsc, b1.w);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = softplus_base2(t[j]);
  if (SAVE) {
    if (zrow) {
      reinterpret_cast<float4*>(zrow)[0] = make_float4(t[0] * kInvB2, t[1] * kInvB2, t[2] * kInvB2, t[3] * kInvB2);
      reinterpret_cast<float4*>(zrow)[1] = make_float4(t[4] * kInvB2, t[5] * kInvB2, t[6] * kInvB2, t[7] * kInvB2);
    }
    if (urow) {
      const float us = kInvB2 * kInvSqrt2;
      reinterpret_cast<float4*>(urow)[0] = make_float4(r[0] * us, r[1] * us, r[2] * us, r[3] * us);
      reinterpret_cast<float4*>(urow)[1] = make_float4(r[4] * us, r[5] * us, r[6] * us, r[7] * us);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = pack_half2(r[2 * i], r[2 * i + 1]);
}

template <bool SAVE>
static __global__ void __launch_bounds__(CH_THREADS, 1)
sdf_chain_tc_kernel(const __grid_constant__ ChainArgs a, int* __restrict__ fault, long long* __restrict__ dbg) {
  using namespace tc;
  // optional timeline of CTA 0 (debug, vdn_debug_timeline): dbg[role * 256 + 4 * phase + ev] = clock64()
  const bool rec = dbg != nullptr && blockIdx.x == 0;
#define CH_TL(role, ph, ev) do { if (rec && (ph) < 64) dbg[(role) * 256 + 4 * (ph) + (ev)] = clock64(); } while (0)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[CH_WSTAGES], w_empty[CH_WSTAGES], a_ready[2], d_full, d_drained;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sW = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* sB = reinterpret_cast<float*>(smem_raw + (sW - smem_u32(smem_raw)) + CH_WSTAGES * CH_W_STAGE);
  const float4* sE = reinterpret_cast<const float4*>(sB + VDN_MAX_LAYERS * 256);
  const long long ntiles = (a.N + 127) / 128;
  const long long G = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < CH_WSTAGES; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int j = 0; j < 2; ++j) mbar_init(smem_u32(&a_ready[j]), CH_EPI_WARPS * 32);
    mbar_init(smem_u32(&d_full), 1);
    mbar_init(smem_u32(&d_drained), CH_EPI_WARPS * 32);
    mbar_fence_init();
  }
  for (int i = tid; i < a.L * 256; i += CH_THREADS) {   // biases in base-2 units
    const int l = i >> 8, n = i & 255;
    sB[i] = n < a.layer[l].out_dim ? a.packed[a.layer[l].bias_off + n] * kB2 : 0.0f;
  }
  if (tid < 64) {                                        // embedding table (zero rows beyond d_e)
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = tid;
    if (c < 3) {
      e = make_float4(1.f, 0.f, (float)c, 1.f);
    } else if (c < a.d_e) {
      const int k = (c - 3) / 6, rem = (c - 3) - 6 * k;
      e = make_float4((float)(1 << k), rem < 3 ? 0.f : 1.5707963267948966f, (float)(rem < 3 ? rem : rem - 3), 0.f);
    }
    const_cast<float4*>(sE)[c] = e;
  }
  if (warp == CH_EPI_WARPS) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // accumulator D: columns [0,256); fp16 activation tiles: slot X columns [256,384), slot Y columns [384,512)
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  if (warp < CH_EPI_WARPS) {
    // ================= embedding + epilogue warps =================
    // warp (q, h): TMEM lane quarter q (rows 32q..32q+31), columns [32 ch + 8 h, +8) of every 32-column chunk ch of D,
    // i.e. packed columns [16 ch + 4 h, +4) of the slot's A tile.
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t tD = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 8);
    const uint32_t tA0 = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(h * 4);
    uint32_t dcnt = 0;                                // completions of d_full consumed
    float yX0 = 0.f, yX1 = 0.f, yX2 = 0.f, yY0 = 0.f, yY1 = 0.f, yY2 = 0.f;
    auto load_point = [&](long long tile, float& y0, float& y1, float& y2) {
      const long long m = tile * 128 + row;
      y0 = y1 = y2 = 0.f;
      if (m < a.N) {
        if (a.x) {
          y0 = a.x[m * 3] * a.scale; y1 = a.x[m * 3 + 1] * a.scale; y2 = a.x[m * 3 + 2] * a.scale;
        } else {
          const unsigned mu = (unsigned)m, t = mu / (unsigned)a.nz;   // N < 2^31 (checked by the launcher)
          const int k = (int)(mu - t * (unsigned)a.nz);
          const unsigned ti = t / (unsigned)a.ny;
          const int j = (int)(t - ti * (unsigned)a.ny);
          const int i = a.i0 + (int)ti;
          y0 = a.xs[i] * a.scale; y1 = a.ys[j] * a.scale; y2 = a.zs[k] * a.scale;
        }
      }
    };
    // positional encoding (times beta/ln2) -> chunks 0 and 1 of the slot's A tile, then signal the slot
    auto write_pe = [&](int s, float y0, float y1, float y2) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = ch * 32 + h * 8 + j;
          v[j] = c < a.d_e ? chain_embed_col(sE, y0, y1, y2, c) * kB2 : 0.0f;
        }
        uint32_t p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = pack_half2(v[2 * i], v[2 * i + 1]);
        tmem_st4(tA0 + (uint32_t)(s * 128 + ch * 16), p);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&a_ready[s]));
    };
    if ((long long)blockIdx.x < ntiles) { load_point(blockIdx.x, yX0, yX1, yX2); write_pe(0, yX0, yX1, yX2); }
    if (blockIdx.x + G < ntiles) { load_point(blockIdx.x + G, yY0, yY1, yY2); write_pe(1, yY0, yY1, yY2); }
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int l = 0; l < a.Lrun && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const float* sb = sB + l * 256 + h * 8;
        const float dsc = (l == a.skip) ? kInvSqrt2 : 1.0f;
        for (int s = 0; s < 2 && ok; ++s) {
          if (s && !hasY) break;
          float y0 = s ? yY0 : yX0, y1 = s ? yY1 : yX1, y2 = s ? yY2 : yX2;
          if (tid == 0) CH_TL(0, dcnt, 0);
          ok = mbar_wait(smem_u32(&d_full), dcnt & 1);
          if (tid == 0) CH_TL(0, dcnt, 1);
          ++dcnt;
          tc_fence_after();
          const bool last_run = (l == a.Lrun - 1);
          const long long m = (tX + s * G) * 128 + row;
          if (l == a.L - 1) {                          // sdf output layer (value-only mode)
            if (h == 0) {
              float v[8];
              tmem_ld8(tD, v);
              tmem_ld_wait();
              if (m < a.N) a.out[m * a.lds] = fmaf(v[0] * dsc, kInvB2, a.packed[Ly.bias_off]) * (a.out_mul / a.scale);
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&d_drained));
          } else {
            // drain this thread's 8 x 8 accumulator columns into registers, then release D for the other slot's MMAs
            float v[8][8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) tmem_ld8(tD + (uint32_t)(ch * 32), v[ch]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(smem_u32(&d_drained));
            if (tid == 0) CH_TL(0, dcnt - 1, 2);
            if (l == a.Lrun - 2 && a.x) {              // the slot's next point: pull its line towards L1 a phase early
              const long long mn = (tX + (2 + s) * G) * 128 + row;
              if (mn < a.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.x + mn * 3));
            }
            const int d_e_tail = (l + 1 == a.skip) ? a.d_e : 0;
            const uint32_t tA = tA0 + (uint32_t)(s * 128);
            float* zrow = nullptr;
            float* urow = nullptr;
            if (SAVE && m < a.N) {
              if (a.save_z[l]) zrow = a.save_z[l] + m * a.ldz + h * 8;
              if (a.save_u && l + 1 == a.skip) urow = a.save_u + m * a.ldz + h * 8;
            }
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              uint32_t p[4];
              if (ch * 32 + 32 <= Ly.out_dim) {      // chunk of real outputs (uniform over the CTA)
                chain_epi8<SAVE>(v[ch], sb + ch * 32, dsc, p, zrow ? zrow + ch * 32 : nullptr, urow ? urow + ch * 32 : nullptr);
              } else {                                // tail of the layer before the skip connection / zero padding
                const int n0 = ch * 32 + h * 8;
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float t = n0 + j < Ly.n_mma ? fmaf(v[ch][j], dsc, sb[ch * 32 + j]) : 0.0f;
                  r[j] = chain_ragged_elem(sE, t, n0 + j, Ly.out_dim, d_e_tail, y0, y1, y2);
                  if (SAVE && n0 + j < Ly.out_dim) {
                    if (zrow) zrow[ch * 32 + j] = t * kInvB2;
                    if (urow) urow[ch * 32 + j] = r[j] * (kInvB2 * kInvSqrt2);
                  }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = pack_half2(r[2 * i], r[2 * i + 1]);
              }
              if (!last_run) tmem_st4(tA + (uint32_t)(ch * 16), p);
            }
            if (!last_run) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(smem_u32(&a_ready[s]));
            }
            if (tid == 0) CH_TL(0, dcnt - 1, 3);
          }
          if (last_run) {                              // the slot's next tile
            const long long tn = tX + (2 + s) * G;
            if (tn < ntiles) {
              load_point(tn, y0, y1, y2);
              if (s) { yY0 = y0; yY1 = y1; yY2 = y2; } else { yX0 = y0; yX1 = y1; yX2 = y2; }
              write_pe(s, y0, y1, y2);
            }
          }
        }
      }
    }
  } else if (tid == CH_EPI_WARPS * 32) {
    // ================= MMA issuer: A from tensor memory, weights from shared memory =================
    uint32_t acnt[2] = {0, 0};
    uint32_t wt = 0, drained = 0;
    const uint32_t tAcol = tmem_base + 256u;
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int l = 0; l < a.Lrun && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const uint32_t idesc = umma_idesc_f16(128, (uint32_t)Ly.n_mma);
        for (int s = 0; s < 2 && ok; ++s) {
          if (s && !hasY) break;
          // the accumulator of the previous phase must have been drained (first phase ever: passes immediately)
          CH_TL(1, drained, 0);
          ok = mbar_wait(smem_u32(&d_drained), (drained & 1) ^ 1);
          CH_TL(1, drained, 1);
          ++drained;
          ok = ok && mbar_wait(smem_u32(&a_ready[s]), acnt[s] & 1);
          ++acnt[s];
          tc_fence_after();
          CH_TL(1, drained - 1, 2);
          // K blocks of 64.  Slot X streams all of the layer's weight tiles; the first `keep` stay in the ring for
          // slot Y, the others are released at once and fetched again for Y, so that the ring always has room to run
          // ahead into the next layer.
          const int keep = Ly.nkb < CH_KEEP ? Ly.nkb : CH_KEEP;
          for (int kb = 0; kb < Ly.nkb && ok; ++kb) {
            const bool held = (s == 1 && kb < keep);       // tile slot X has already seen arrive
            const uint32_t w = wt + (uint32_t)(s == 0 || held ? kb : Ly.nkb + kb - keep);
            const uint32_t ws = w % CH_WSTAGES, wph = (w / CH_WSTAGES) & 1;
            if (!held) {
              ok = mbar_wait(smem_u32(&w_full[ws]), wph);
              tc_fence_after();
            }
            const uint32_t b0 = sW + ws * CH_W_STAGE;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16_ts(tmem_base, tAcol + (uint32_t)(s * 128 + kb * 32 + ks * 8), umma_desc_sw128(b0 + ks * 32), idesc,
                          (kb | ks) ? 1u : 0u);
            if (s == 1 || !hasY || kb >= keep) umma_commit(smem_u32(&w_empty[ws]));
          }
          umma_commit(smem_u32(&d_full));
          CH_TL(1, drained - 1, 3);
        }
        const int keep = Ly.nkb < CH_KEEP ? Ly.nkb : CH_KEEP;
        wt += (uint32_t)(hasY ? 2 * Ly.nkb - keep : Ly.nkb);
      }
    }
  } else if (tid == (CH_EPI_WARPS + 1) * 32) {
    // ================= weight stream (TMA engine), in the order the MMA issuer consumes the tiles =================
    uint32_t wt = 0;
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int l = 0; l < a.Lrun && ok; ++l) {
        const ChainLayer& Ly = a.layer[l];
        const uint32_t bytes = (uint32_t)Ly.n_mma * 128u;
        const int keep = Ly.nkb < CH_KEEP ? Ly.nkb : CH_KEEP;
        const int nfetch = hasY ? 2 * Ly.nkb - keep : Ly.nkb;
        for (int f = 0; f < nfetch && ok; ++f, ++wt) {
          const int kb = f < Ly.nkb ? f : f - Ly.nkb + keep;
          const uint32_t ws = wt % CH_WSTAGES, wph = (wt / CH_WSTAGES) & 1;
          ok = mbar_wait_backoff(smem_u32(&w_empty[ws]), wph ^ 1, 256);
          mbar_arrive_expect_tx(smem_u32(&w_full[ws]), bytes);
          bulk_g2s(sW + ws * CH_W_STAGE, a.packed + Ly.img_off + ((size_t)kb * Ly.out_ld + Ly.row0) * 32, bytes, smem_u32(&w_full[ws]));
        }
      }
    }
  }
  if (!ok && fault) *fault = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == CH_EPI_WARPS) tmem_dealloc(tmem_base, 512);
#undef CH_TL
}

// Returns -1 when the configuration is not supported by the fused chain (caller falls back to the layer-wise path).
static inline int launch_sdf_chain(const MlpLayout& ly, int d_in, int multires, int d_hidden, int skip, float scale,
                                   const float* packed, const float* x, const float* xs, const float* ys,
                                   const float* zs, int ny, int nz, int i0, long long N, float* out, int lds, float out_mul,
                                   cudaStream_t st, int orot_last = 0, float* const* save_z = nullptr, float* save_u = nullptr,
                                   int ldz = 0) {
  const int d_e = d_in * (1 + 2 * multires);
  if (d_in != 3 || d_hidden != 256 || d_e > 64 || ly.L < 2 || ly.L > VDN_MAX_LAYERS) return -1;
  for (int l = 1; l < ly.L; ++l)
    if (ly.in_dim[l] != 256) return -1;
  if (N > 0x7fffffffLL) return (int)cudaErrorInvalidValue;   // 32-bit point indices inside the kernel
  ChainArgs a;
  a.L = ly.L; a.skip = skip; a.d_e = d_e; a.multires = multires; a.scale = scale; a.out_mul = out_mul;
  a.packed = packed; a.x = x; a.xs = xs; a.ys = ys; a.zs = zs; a.ny = ny; a.nz = nz; a.i0 = i0; a.N = N; a.out = out; a.lds = lds;
  for (int l = 0; l < ly.L; ++l) {
    ChainLayer& c = a.layer[l];
    c.img_off = ly.off_ih[l]; c.bias_off = ly.off_b[l]; c.out_ld = ly.out_ld[l]; c.out_dim = ly.out_dim[l];
    c.n_mma = (l == ly.L - 1) ? 16 : ((ly.out_dim[l] + 15) & ~15);
    c.nkb = (ly.in_dim[l] + 63) / 64;
    // last layer: only the sdf output is needed; with a rotated image (orot = 1) it sits at position out_dim - 1
    c.row0 = (l == ly.L - 1 && orot_last) ? ly.out_dim[l] - orot_last : 0;
    if (c.n_mma > 256 || c.nkb > 4 || (c.row0 & 7)) return -1;
  }
  // value-only: all L layers, sdf out.  Training forward (save_z given): layers 0..L-2 with stored pre-activations.
  const bool save = save_z != nullptr;
  a.Lrun = save ? ly.L - 1 : ly.L;
  a.save_u = save ? save_u : nullptr;
  a.ldz = ldz;
  for (int l = 0; l < VDN_MAX_LAYERS; ++l) a.save_z[l] = (save && l < ly.L - 1) ? save_z[l] : nullptr;
  static int num_sms = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sdf_chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sdf_chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long ntiles = (N + 127) / 128;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  double flops = 0.0;
  for (int l = 0; l < a.Lrun; ++l) flops += 2.0 * (double)N * a.layer[l].n_mma * a.layer[l].nkb * 64;
  double bytes = (double)N * (x ? 16.0 : 4.0);
  if (save) {
    bytes = (double)N * 12.0;
    for (int l = 0; l < a.Lrun; ++l) bytes += a.save_z[l] ? 4.0 * N * a.layer[l].out_dim : 0.0;
    if (a.save_u && skip >= 1) bytes += 4.0 * N * a.layer[skip - 1].out_dim;
  }
  prof_begin(PROF_CHAIN, st, flops, bytes);
  if (save) {
    VDN_LAUNCH(sdf_chain_tc_kernel<true>, grid, CH_THREADS, CH_SMEM, st, a, g_tc_fault, g_tc_dbg);
  } else {
    VDN_LAUNCH(sdf_chain_tc_kernel<false>, grid, CH_THREADS, CH_SMEM, st, a, g_tc_fault, g_tc_dbg);
  }
  prof_end(PROF_CHAIN, st);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

}  // namespace vdn

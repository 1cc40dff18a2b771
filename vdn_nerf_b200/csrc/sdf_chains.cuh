// SDF network on the chain engine (tensor-core mode): training forward, analytic normals and the two-phase backward
// of SURVEY.md Appendix A as four chain launches + one grouped weight-gradient launch, all saved tensors 16-bit.
//
//   reference: SDFNetwork.forward / .gradient (dpt_models/fields.py:72-108) and autograd's double backward through them.
//
// Saved tensors (16-bit ones TILE-BLOCKED, chain_engine.cuh; Npad = N rounded up to 128 rows; "A16" etc. are the names
// used in DESIGN.md).  Every 16-bit tensor is fp16 (11 significant bits; the stated <= 2e-3 tolerance on normals needs
// them).  Cotangents are ~1e-6 and would underflow fp16, so a backward call carries one power-of-two loss scale sigma
// (amax_sigma_kernel below: the largest incoming cotangent maps into [0.5, 1)); every cotangent tensor holds sigma times
// the true value, and the fp32 results (point gradient, weight gradients) are multiplied by 1 / sigma on the way out:
//   forward blob   E16 [Npad, 64]        kB2 * embedding (layer-0 input; kB2 = beta / ln2: base-2 softplus units)
//                  A16_l [Npad,256]      l = 0..L-2: a'_l = kB2 * softplus(z_l); the layer before the skip connection
//                                        holds [a' | kB2 * e] = sqrt2 * kB2 * (skip-layer input)
//   normals blob   D16_l [Npad,256]      delta_l = softplus'(z_l) * d sdf / d h_l  (softplus' = 1 - 2^-a')
//                  DE0, DES [Npad,48] fp32  d sdf / d e through layer 0 and through the skip connection
//   backward ws    Q16_0 [Npad,64], Q16_l [Npad,256]  sigma * q-bar_l (phase 1);  ZG16_l sigma * injected cotangents;
//                  ZB16_l  sigma * z-bar_l * (dsc_l / kB2)  (phase 2; pre-scaled so that ZB^T A16 = sigma z-bar^T u);
//                  FB16 [Npad,256] sigma * feature cotangent; SB / ONES [Npad,8] (column 0: sigma d_sdf / kB2, ones)
// Nothing else of a layer reaches HBM: softplus'(z) and softplus''(z) * a are recomputed from A16 and D16.
#pragma once
#include "chain_engine.cuh"
#include "wgrad16.cuh"
#include "pointwise.cuh"

namespace vdn {

extern int g_chain;    // api.cu: fused training chains enabled (tensor-core mode only)


inline long long pad128(long long n) { return (n + 127) / 128 * 128; }

// ---- pointwise producers of the 16-bit chain inputs -----------------------------------------------------
// All of them write TILE-BLOCKED tensors (chain_engine.cuh) with one 16-byte store per thread: thread i owns the eight
// columns of blocked position 8 i, consecutive threads are consecutive rows, so a warp writes 512 contiguous bytes.
__device__ __forceinline__ void store8_h(__half* dst, long long i, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(ce::pack_h2(v[0], v[1]), ce::pack_h2(v[2], v[3]), ce::pack_h2(v[4], v[5]),
                                                      ce::pack_h2(v[6], v[7]));
}
// saturating variant for scaled cotangents
__device__ __forceinline__ void store8_hs(__half* dst, long long i, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(ce::pack_h2_sat(v[0], v[1]), ce::pack_h2_sat(v[2], v[3]),
                                                      ce::pack_h2_sat(v[4], v[5]), ce::pack_h2_sat(v[6], v[7]));
}

// ---- loss scale of one backward call -------------------------------------------------------------------------
// sig[0] = sigma = 2^-e with the largest |cotangent * mul| over up to three fp32 sources in [2^(e-1), 2^e), sig[1] = 1 / sigma
// (sigma = 1 when every cotangent is zero or not finite).  sig[2] (max, as ordered bits) and sig[3] (block counter) must
// be zero on entry (a memset node precedes the launch).  A source of more than 2^21 elements is SAMPLED (every rstride-th
// row, its maximum taken times four): the scale only has to place the cotangents inside fp16's 40 binades - there are 16
// binades of headroom above the maximum and conversions saturate - and reading all of a [65536, 256] cotangent twice
// would cost more than the conversion it prepares.
struct AmaxSrc { const float* p; long long rows; int w, ld; float mul; int rstride; };
struct AmaxArgs { AmaxSrc s[3]; };
static __global__ void amax_sigma_kernel(const __grid_constant__ AmaxArgs a, float* __restrict__ sig) {
  float mx = 0.0f;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    const AmaxSrc& s = a.s[k];
    if (!s.p) continue;
    float m = 0.0f;
    if (s.rstride > 1) {
      const long long nr = s.rows / s.rstride;
      for (long long i = t0; i < nr * s.w; i += nt) m = fmaxf(m, fabsf(__ldg(s.p + (i / s.w) * s.rstride * s.ld + (i % s.w))));
      mx = fmaxf(mx, 4.0f * m * s.mul);
      continue;
    }
    const long long tot = s.rows * s.w;
    if (s.ld == s.w && (((uintptr_t)s.p) & 15) == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(s.p);
      for (long long i = t0; i < tot / 4; i += nt) {
        const float4 v = __ldg(p4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
      for (long long i = (tot & ~3LL) + t0; i < tot; i += nt) m = fmaxf(m, fabsf(s.p[i]));
    } else {
      for (long long i = t0; i < tot; i += nt) m = fmaxf(m, fabsf(s.p[(i / s.w) * s.ld + (i % s.w)]));
    }
    mx = fmaxf(mx, m * s.mul);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  unsigned* bits = reinterpret_cast<unsigned*>(sig + 2);
  if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(bits, __float_as_uint(mx));     // non-negative floats order like their bits
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(sig + 3), 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const float am = __uint_as_float(atomicMax(bits, 0u));
      int e = 0;
      if (am > 0.0f && am < 3.0e38f) {
        frexpf(am, &e);                 // am = f * 2^e, f in [0.5, 1)
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
      }
      sig[0] = ldexpf(1.0f, -e);
      sig[1] = ldexpf(1.0f, e);
    }
  }
}
static inline int launch_sigma(float* sig, cudaStream_t st, const float* p0, long long r0, int w0, int ld0, float m0,
                               const float* p1 = nullptr, long long r1 = 0, int w1 = 1, int ld1 = 1, float m1 = 1.0f,
                               const float* p2 = nullptr, long long r2 = 0, int w2 = 1, int ld2 = 1, float m2 = 1.0f) {
  cudaError_t e = cudaMemsetAsync(sig, 0, 4 * sizeof(float), st);
  if (e != cudaSuccess) return ce::trace_err((int)e, "loss-scale memset");
  AmaxArgs a;
  a.s[0] = {p0, r0, w0, ld0, m0, 1}; a.s[1] = {p1, r1, w1, ld1, m1, 1}; a.s[2] = {p2, r2, w2, ld2, m2, 1};
  long long tot = 0;
  for (int k = 0; k < 3; ++k) {
    const long long n = a.s[k].rows * a.s[k].w;
    if (n > (1LL << 21)) a.s[k].rstride = (int)(n >> 20);
    tot += n / a.s[k].rstride;
  }
  int blocks = (int)((tot / 4 + 255) / 256);
  blocks = blocks < 1 ? 1 : (blocks > 4 * ce::num_sms() ? 4 * ce::num_sms() : blocks);
  VDN_LAUNCH(amax_sigma_kernel, blocks, 256, 0, st, a, sig);
  return ce::trace_err((int)(cudaError_t)::vdn::take_launch_error(), "amax_sigma_kernel");
}

// column c of the positional embedding of a d-dimensional point x (embedder.py:15-36): [x | sin(2^k x) | cos(2^k x)]_k
__device__ __forceinline__ float embed_col(const float* x, int d, int L, int c, float scale) {
  if (c >= d * (1 + 2 * L)) return 0.0f;
  if (c < d) return x[c] * scale;
  const int k = (c - d) / (2 * d), rem = (c - d) - 2 * d * k, j = rem < d ? rem : rem - d;
  const float y = x[j] * scale * (float)(1 << k);
  return rem < d ? sinf(y) : cosf(y);
}

// E16 [Npad, 64] = fp16(kB2 * e(x * scale)), zero beyond d_e and beyond row N  (d = 3)
static __global__ void sdf_embed16_kernel(const float* __restrict__ x, long long N, long long Npad, int L, float scale,
                                          __half* __restrict__ e16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 8) return;
  long long m;
  int c;
  ce::blk_decode(i * 8, 64, &m, &c);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = m < N ? embed_col(x + m * 3, 3, L, c + j, scale) * ce::kB2 : 0.0f;
  store8_h(e16, i, v);
}

// delta of the last hidden layer: D16[m, c] = fp16((1 - 2^-A16[m, c]) * w[c]), w = first row of the last weight
static __global__ void sdf_delta_last_kernel(const __half* __restrict__ a16, const float* __restrict__ wrow,
                                             long long Npad, int width, __half* __restrict__ d16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 columns
  if (i >= Npad * 32) return;
  long long m;
  int c;
  ce::blk_decode(i * 8, 256, &m, &c);     // both tensors are tile-blocked [.., 256]: same position in each
  float a8[8], v[8];
  ce::unpack_h8(*reinterpret_cast<const uint4*>(a16 + i * 8), a8);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = c + j < width ? (1.0f - exp2f(-a8[j])) * wrow[c + j] : 0.0f;
  store8_h(d16, i, v);
}

// Q16_0[m, c] = fp16(sigma * (J_e n-bar)[c]): forward-mode product with the embedding Jacobian (pointwise.cuh embed_jvp_kernel)
static __global__ void sdf_qbar0_kernel(const float* __restrict__ x, long long N, long long Npad, int L, float scale,
                                        const float* __restrict__ nbar, const float* __restrict__ sigma,
                                        __half* __restrict__ q16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 8) return;
  long long m;
  int c0;
  ce::blk_decode(i * 8, 64, &m, &c0);
  const float sg = __ldg(sigma);
  float v[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int c = c0 + jj;
    float r = 0.0f;
    if (m < N && c < 3 + 6 * L) {
      if (c < 3) {
        r = nbar[m * 3 + c];
      } else {
        const int k = (c - 3) / 6, rem = (c - 3) - 6 * k, j = rem % 3;
        const float f = (float)(1 << k);
        const float y = x[m * 3 + j] * scale * f;
        r = (rem < 3 ? f * cosf(y) : -f * sinf(y)) * nbar[m * 3 + j];
      }
    }
    v[jj] = r * sg;
  }
  store8_hs(q16, i, v);
}

// dst[m, c] = fp16(src[m * lds + c] * mul * sigma) for c < w (zero beyond, zero rows beyond N; src null: zeros; sigma null: 1).
// W = width of the blocked tensor (multiple of 8).
static __global__ void rows_to_16_kernel(const float* __restrict__ src, int lds, int w, float mul_in,
                                         const float* __restrict__ sigma, long long N, long long Npad,
                                         __half* __restrict__ dst, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * (W / 8)) return;
  long long m;
  int c;
  ce::blk_decode(i * 8, W, &m, &c);
  float v[8];
  const float mul = sigma ? mul_in * __ldg(sigma) : mul_in;
  const float* r = src ? src + m * lds + c : nullptr;
  if (r && m < N && c + 8 <= w && ((((uintptr_t)r) & 15) == 0)) {
    const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
    v[0] = a.x * mul; v[1] = a.y * mul; v[2] = a.z * mul; v[3] = a.w * mul;
    v[4] = b.x * mul; v[5] = b.y * mul; v[6] = b.z * mul; v[7] = b.w * mul;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (r && m < N && c + j < w) ? r[j] * mul : 0.0f;
  }
  store8_hs(dst, i, v);
}
// dst[m, :] = fp16(sigma * [a[m, 0..wa) | b[m, 0..wb) | 0 ...]) up to W columns (null source: zeros), zero rows beyond N
static __global__ void gather2_16_kernel(const float* __restrict__ a, int lda, int wa, const float* __restrict__ b, int ldb,
                                         int wb, const float* __restrict__ sigma, long long N, long long Npad,
                                         __half* __restrict__ dst, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * (W / 8)) return;
  long long m;
  int c0;
  ce::blk_decode(i * 8, W, &m, &c0);
  const float sg = __ldg(sigma);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float r = 0.0f;
    if (m < N) {
      if (c < wa) { if (a) r = a[m * lda + c]; }
      else if (c < wa + wb) { if (b) r = b[m * ldb + (c - wa)]; }
    }
    v[j] = r * sg;
  }
  store8_hs(dst, i, v);
}
// dst[m, 0] = 1 for m < N, everything else zero: the "ones" operand that turns a column sum into a GEMM row
static __global__ void ones_col16_kernel(long long N, long long Npad, __half* __restrict__ dst, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * (W / 8)) return;
  long long m;
  int c;
  ce::blk_decode(i * 8, W, &m, &c);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m < N && c == 0) v[0] = 1.0f;
  store8_h(dst, i, v);
}
template <class K, class... A>
static inline int launch1d(K kern, long long total, cudaStream_t st, A... args) {
  if (total <= 0) return 0;
  VDN_LAUNCH(kern, (unsigned)((total + 255) / 256), 256, 0, st, args...);
  return ce::trace_err((int)(cudaError_t)::vdn::take_launch_error(), "pointwise producer");
}

// ---- layouts ----------------------------------------------------------------------------------------------
struct SdfChainBufs {
  long long Npad;
  __half* E16;
  __half* A16[VDN_MAX_LAYERS];
  __half* D16[VDN_MAX_LAYERS];
  float* DE0; float* DES;
  __half* Q16[VDN_MAX_LAYERS + 1];
  __half* ZG16[VDN_MAX_LAYERS];
  __half* ZB16[VDN_MAX_LAYERS];
  __half* FB16; __half* SB; __half* ONES;
  float* EB; float* ES;
  float* sig;      // {sigma, 1 / sigma, scratch, scratch} of the backward call
};
static inline long long sdf_chain_blob_floats(int L, long long N) { return pad128(N) * (32 + (long long)(L - 1) * 128); }
static inline long long sdf_chain_blobg_floats(int L, long long N) { return pad128(N) * ((long long)(L - 1) * 128 + 96); }
static inline long long sdf_chain_ws_floats(int L, long long N) {
  return pad128(N) * (32 + 3LL * (L - 1) * 128 + 128 + 4 + 4 + 96) + 32;
}
static inline void sdf_chain_carve(int L, long long N, float* blob, float* blobg, float* ws, SdfChainBufs* b) {
  const long long Np = pad128(N);
  b->Npad = Np;
  if (blob) {
    b->E16 = reinterpret_cast<__half*>(blob);
    for (int l = 0; l < L - 1; ++l) b->A16[l] = reinterpret_cast<__half*>(blob + Np * 32 + (long long)l * Np * 128);
  }
  if (blobg) {
    for (int l = 0; l < L - 1; ++l) b->D16[l] = reinterpret_cast<__half*>(blobg + (long long)l * Np * 128);
    b->DE0 = blobg + (long long)(L - 1) * Np * 128;
    b->DES = b->DE0 + Np * 48;
  }
  if (ws) {
    float* p = ws;
    b->Q16[0] = reinterpret_cast<__half*>(p); p += Np * 32;
    for (int l = 1; l <= L - 1; ++l) { b->Q16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    for (int l = 0; l < L - 1; ++l) { b->ZG16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    for (int l = 0; l < L - 1; ++l) { b->ZB16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    b->FB16 = reinterpret_cast<__half*>(p); p += Np * 128;
    b->SB = reinterpret_cast<__half*>(p); p += Np * 4;
    b->ONES = reinterpret_cast<__half*>(p); p += Np * 4;
    b->EB = p; p += Np * 48;
    b->ES = p; p += Np * 48;
    b->sig = p;
  }
}

// Shape of the SDF net as the chains need it (taken from SdfCfg by the caller).
struct SdfShape {
  int L, skip, d_e, multires;
  float scale;
  const MlpLayout* ly;
};
static inline float sdf_dsc(const SdfShape& s, int l) { return l == s.skip ? kInvSqrt2 : 1.0f; }

// ---- training forward: E16 -> [A16_l] -> sdf, feature -------------------------------------------------------
static inline int sdf_chain_forward(const SdfShape& s, const float* packed, const float* x, long long N, float* sdf, int lds,
                                    float* feat, int ldf, float out_mul, int save, const SdfChainBufs& b, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int L = s.L;
  int e = launch1d(sdf_embed16_kernel, b.Npad * 8, st, x, N, b.Npad, s.multires, s.scale, b.E16);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed;
  a.a0 = b.E16; a.a0_ld = 64; a.a0_w = 64;
  int P = 0;
  for (int l = 0; l < L - 1; ++l) {
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[l], ly.out_ld[l], 0, 0, ly.out_dim[l], ly.in_dim[l]);
    p.op = ce::OP_SOFTPLUS; p.width = ly.out_dim[l]; p.dsc = sdf_dsc(s, l);
    p.bias_off = ly.off_b[l]; p.bias_mul = ce::kB2;
    p.a_out = 1; p.a_wr = 256;
    if (save) { p.o16a = b.A16[l]; p.ldo16a = 256; }
    if (l + 1 == s.skip) { p.tail = b.E16; p.ldt = 64; p.tail_w = s.d_e; p.tail_mul = 1.0f; }
  }
  const int lo = L - 1;
  if (sdf) {   // image position out_dim-1 holds output 0 (orot = 1): one N = 16 MMA
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[lo], ly.out_ld[lo], ly.out_dim[lo] - 1, 0, 16, ly.in_dim[lo]);
    p.op = ce::OP_OUT32; p.width = 1; p.dsc = sdf_dsc(s, lo) * ce::kInvB2; p.bias_off = ly.off_b[lo];
    p.o32 = sdf; p.ldo32 = lds; p.o32_c0 = 0; p.o32_w = 1; p.o32_mul = out_mul / s.scale;
  }
  if (feat) {
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[lo], ly.out_ld[lo], 0, 0, ly.out_dim[lo] - 1, ly.in_dim[lo]);
    p.op = ce::OP_OUT32; p.width = ly.out_dim[lo] - 1; p.dsc = sdf_dsc(s, lo) * ce::kInvB2; p.bias_off = ly.off_b[lo] + 1;
    p.o32 = feat; p.ldo32 = ldf; p.o32_c0 = 0; p.o32_w = ly.out_dim[lo] - 1;
  }
  a.P = P;
  return ce::launch(a, st, PROF_CHAIN_TRAIN);
}

// ---- analytic normals: D16_{L-2} -> ... -> DE0 (+ DES), then J_e^T ----------------------------------------------
static inline int sdf_chain_normals(const SdfShape& s, const float* packed, const float* x, long long N,
                                    const SdfChainBufs& b, float* normals, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int L = s.L;
  int e = launch1d(sdf_delta_last_kernel, b.Npad * 32, st, (const __half*)b.A16[L - 2], packed + ly.off_w[L - 1], b.Npad,
                   ly.out_dim[L - 2], b.D16[L - 2]);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed;
  a.a0 = b.D16[L - 2]; a.a0_ld = 256; a.a0_w = 256;
  int P = 0;
  for (int l = L - 2; l >= 0; --l) {     // a_l = delta_l W_l ; epilogue -> delta_{l-1}
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[l], ly.in_ld[l], 0, 0, ly.in_dim[l], ly.out_dim[l]);
    p.dsc = sdf_dsc(s, l);
    if (l > 0) {
      p.op = ce::OP_NSTEP; p.width = ly.out_dim[l - 1];
      p.aux0 = b.A16[l - 1]; p.ld0 = 256;
      p.a_out = 1; p.a_wr = 256;
      p.o16a = b.D16[l - 1]; p.ldo16a = 256;
      if (l == s.skip) { p.o32 = b.DES; p.ldo32 = 48; p.o32_c0 = ly.out_dim[l - 1]; p.o32_w = s.d_e; }
    } else {
      p.op = ce::OP_OUT32; p.width = s.d_e;
      p.o32 = b.DE0; p.ldo32 = 48; p.o32_c0 = 0; p.o32_w = s.d_e;
    }
  }
  a.P = P;
  e = ce::launch(a, st, PROF_CHAIN_TRAIN);
  if (e) return e;
  const long long tot = N * 3;
  VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, 3, N, 3, s.multires, s.scale, b.DE0, 48,
             s.skip >= 0 ? b.DES : nullptr, 48, 1.0f, 1.0f, normals, 3, 0);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// ---- backward ------------------------------------------------------------------------------------------------
static inline int sdf_chain_backward(const SdfShape& s, const float* packed, const float* x, long long N,
                                     const SdfChainBufs& b, const float* d_sdf, int lds, const float* d_feat, int ldf,
                                     const float* d_normals, float* dpacked, float* d_x, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int L = s.L;
  const bool have_n = d_normals != nullptr;
  const int lo = L - 1;
  const int nf = ly.out_dim[lo] - 1;
  // one loss scale for the whole call: |J_e n-bar| <= 2^multires |n-bar|
  int e = launch_sigma(b.sig, st, d_feat, N, nf, ldf, 1.0f, d_sdf, d_sdf ? N : 0, 1, lds, 1.0f / s.scale, d_normals,
                       have_n ? N : 0, 3, 3, (float)(1 << s.multires));
  if (e) return e;
  // ---- phase 1: backward of the normals pass, l = 0 .. L-2 ----
  if (have_n) {
    e = launch1d(sdf_qbar0_kernel, b.Npad * 8, st, x, N, b.Npad, s.multires, s.scale, d_normals, b.sig, b.Q16[0]);
    if (e) return e;
    ce::Args a;
    ce::init_args(&a);
    a.N = N; a.packed = packed; a.sigma = b.sig;
    a.a0 = b.Q16[0]; a.a0_ld = 64; a.a0_w = 64;
    int P = 0;
    for (int l = 0; l <= L - 2; ++l) {    // delta-bar_l = q-bar_l W_l^T ; epilogue -> q-bar_{l+1}, z-bar^g_l
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      ce::set_mma(&p, ly.off_ih[l], ly.out_ld[l], 0, 0, ly.out_dim[l], ly.in_dim[l]);
      p.op = ce::OP_P1STEP; p.width = ly.out_dim[l];
      p.aux0 = b.A16[l]; p.ld0 = 256; p.aux1 = b.D16[l]; p.ld1 = 256;
      p.a_mul = sdf_dsc(s, l + 1);
      p.a_out = l < L - 2 ? 1 : 0; p.a_wr = 256;
      p.o16a = b.Q16[l + 1]; p.ldo16a = 256;
      p.o16b = b.ZG16[l]; p.ldo16b = 256;
      if (l + 1 == s.skip) { p.tail = b.Q16[0]; p.ldt = 64; p.tail_w = s.d_e; p.tail_mul = kInvSqrt2; }
    }
    a.P = P;
    e = ce::launch(a, st, PROF_CHAIN_TRAIN);
    if (e) return e;
  }
  // ---- phase 2: ordinary backward with the injected cotangents, l = L-1 .. 1 (.. 0 for the point gradient) ----
  e = launch1d(rows_to_16_kernel, b.Npad * 32, st, d_feat, ldf, nf, 1.0f, (const float*)b.sig, N, b.Npad, b.FB16, 256);
  if (e) return e;
  e = launch1d(rows_to_16_kernel, b.Npad, st, d_sdf, lds, 1, ce::kInvB2 / s.scale, (const float*)b.sig, N, b.Npad, b.SB, 8);
  if (e) return e;
  if (have_n) {
    e = launch1d(ones_col16_kernel, b.Npad, st, N, b.Npad, b.ONES, 8);
    if (e) return e;
  }
  {
    ce::Args a;
    ce::init_args(&a);
    a.N = N; a.packed = packed; a.sigma = b.sig;
    a.a0 = b.FB16; a.a0_ld = 256; a.a0_w = 256;
    a.row_off[0] = ly.off_w[lo]; a.row_len[0] = ly.in_dim[lo];
    int P = 0;
    for (int l = lo; l >= 1; --l) {      // u-bar = z-bar_l W_l ; epilogue -> z-bar_{l-1}
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      const int k = l == lo ? nf : ly.out_dim[l];
      ce::set_mma(&p, ly.off_iht[l], ly.in_ld[l], 0, 0, ly.in_dim[l], k);
      p.op = ce::OP_P2STEP; p.width = ly.out_dim[l - 1]; p.dsc = sdf_dsc(s, l);
      if (l == lo && d_sdf) { p.r1 = d_sdf; p.r1_stride = lds; p.r1_mul = sdf_dsc(s, l) / s.scale; p.r1_row = 0; p.r1_scaled = 1; }
      p.aux0 = b.A16[l - 1]; p.ld0 = 256;
      if (have_n) { p.aux1 = b.ZG16[l - 1]; p.ld1 = 256; }
      p.a_out = (l > 1 || d_x) ? 1 : 0; p.a_wr = 256;
      p.o16a = b.ZB16[l - 1]; p.ldo16a = 256; p.o16a_mul = sdf_dsc(s, l - 1) * ce::kInvB2;
      if (l == s.skip && d_x) { p.o32 = b.ES; p.ldo32 = 48; p.o32_c0 = ly.out_dim[l - 1]; p.o32_w = s.d_e; p.o32_unscale = 1; }
    }
    if (d_x) {
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      ce::set_mma(&p, ly.off_iht[0], ly.in_ld[0], 0, 0, ly.in_dim[0], ly.out_dim[0]);
      p.op = ce::OP_OUT32; p.width = s.d_e; p.dsc = sdf_dsc(s, 0);
      p.o32 = b.EB; p.ldo32 = 48; p.o32_c0 = 0; p.o32_w = s.d_e; p.o32_unscale = 1;
    }
    a.P = P;
    e = ce::launch(a, st, PROF_CHAIN_TRAIN);
    if (e) return e;
  }
  // ---- weight and bias gradients: one grouped launch ----
  {
    wg::Builder w(N, dpacked, b.sig);
    int mE = w.add_y(b.E16, 64, ly.in_dim[0]), mA[VDN_MAX_LAYERS], mD[VDN_MAX_LAYERS], mQ[VDN_MAX_LAYERS + 1], mZ[VDN_MAX_LAYERS];
    for (int l = 0; l < L - 1; ++l) {
      mA[l] = w.add_y(b.A16[l], 256, ly.in_dim[l + 1]);
      mZ[l] = w.add_x(b.ZB16[l], 256, ly.out_dim[l]);
      if (have_n) mD[l] = w.add_x(b.D16[l], 256, ly.out_dim[l]);
    }
    if (have_n) {
      mQ[0] = w.add_y(b.Q16[0], 64, ly.in_dim[0]);
      for (int l = 1; l <= L - 1; ++l) mQ[l] = w.add_y(b.Q16[l], 256, ly.in_dim[l]);
    }
    const int mF = w.add_x(b.FB16, 256, nf), mS = w.add_x(b.SB, 8, 1);
    const int mO = have_n ? w.add_x(b.ONES, 8, 1) : 0;
    for (int l = 0; l < L - 1; ++l) {     // W-bar_l = [z-bar_l ; delta_l]^T [u_l ; q-bar_l]
      wg::Job* j = w.add_job(ly.out_dim[l], ly.in_dim[l], ly.off_w[l], ly.in_ld[l], 1.0f, ly.off_b[l],
                             ce::kB2 / sdf_dsc(s, l));
      wg::Builder::add_seg(j, mZ[l], 0, l == 0 ? mE : mA[l - 1], 0);
      if (have_n) wg::Builder::add_seg(j, mD[l], 0, mQ[l], 0);
    }
    {   // last layer, feature rows 1 .. nf
      wg::Job* j = w.add_job(nf, ly.in_dim[lo], ly.off_w[lo] + ly.in_ld[lo], ly.in_ld[lo], sdf_dsc(s, lo) * ce::kInvB2,
                             ly.off_b[lo] + 1, 1.0f);
      wg::Builder::add_seg(j, mF, 0, mA[lo - 1], 0);
    }
    if (d_sdf || have_n) {   // last layer, row 0: sdf cotangent, and the column sum of q-bar_{L-1} (a_{L-1} IS that row)
      wg::Job* j = w.add_job(1, ly.in_dim[lo], ly.off_w[lo], ly.in_ld[lo], 1.0f, ly.off_b[lo], ce::kB2);
      wg::Builder::add_seg(j, mS, 0, mA[lo - 1], 0);
      if (have_n) wg::Builder::add_seg(j, mO, 0, mQ[lo], 0);
    }
    e = w.launch(st, PROF_WGRAD16);
    if (e) return e;
  }
  // ---- gradient w.r.t. the points (learnable poses) ----
  if (d_x) {
    const long long tot = N * 3;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, 3, N, 3, s.multires, s.scale, b.EB, 48,
               s.skip >= 0 ? b.ES : nullptr, 48, 1.0f, s.scale, d_x, 3, 0);
    if (have_n) {
      VDN_LAUNCH(embed_second_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, 3, N, 3, s.multires, s.scale, d_normals, 3,
                 b.DE0, 48, s.skip >= 0 ? b.DES : nullptr, 48, s.scale, d_x, 3);
    }
    e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
  }
  return 0;
}

}  // namespace vdn

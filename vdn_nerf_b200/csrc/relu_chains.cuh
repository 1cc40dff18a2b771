// ReLU networks on the chain engine (tensor-core mode): RenderingNetwork (colour / depth-feature head,
// dpt_models/fields.py:148-176) and the NeRF++ background field (fields.py:324-355), forward and backward, each pass
// ONE chain launch + one grouped weight-gradient launch.
//
// Everything 16-bit is fp16 (chain_engine.cuh): the forward passes save the post-ReLU activations H16_l (at once the
// ReLU mask of the backward chain and an operand of the weight gradient); the backward passes run cotangents scaled by
// the call's power-of-two loss scale sigma (sdf_chains.cuh).  Inputs wider than 256 columns (289-wide colour
// input, 340-wide NeRF skip layer, 283-wide view layer) are split: the narrow part (extras / embedding) runs first as a
// small MMA whose accumulator is parked in an fp16 scratch (L2 resident) and added in the epilogue of the main part.
#pragma once
#include "sdf_chains.cuh"

namespace vdn {

// ------------------------------------------------------------------------------------------------------------
// RenderingNetwork.  Packed layer 0 has its input columns rotated to [feature | extras] (extras = points, view
// embedding, normals in the reference's order).
// ------------------------------------------------------------------------------------------------------------
struct RnShape {
  int L, F, nextra, d_out, mode, multires_view, squeeze_out, ldIn;
  const MlpLayout* ly;
};

// X16 [Npad, 64]: the extras of the input row (fields.py:154), zero padded
static __global__ void rn_extras16_kernel(const float* __restrict__ pts, const float* __restrict__ nrm,
                                          const float* __restrict__ view, int L, int mode, long long N, long long Npad,
                                          __half* __restrict__ x16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 8) return;
  long long m;
  int c0;
  ce::blk_decode(i * 8, 64, &m, &c0);
  const int nview = (mode != 1) ? 3 * (1 + 2 * L) : 0;
  const int nnrm = (mode != 2) ? 3 : 0;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float r = 0.0f;
    if (m < N && c < 3 + nview + nnrm) {
      if (c < 3) r = pts[m * 3 + c];
      else if (c < 3 + nview) r = embed_col(view + m * 3, 3, L, c - 3, 1.0f);
      else r = nrm[m * 3 + (c - 3 - nview)];
    }
    v[j] = r;
  }
  store8_h(x16, i, v);
}

// ZL16[m, c] = fp16(sigma * d_out[m, c] * act'(out[m, c])) for c < w, zero padded to 128 columns.  kind 0: sigmoid, 1: relu
static __global__ void rn_zlast16_kernel(const float* __restrict__ d_out, const float* __restrict__ out, int w, int kind,
                                         const float* __restrict__ sigma, long long N, long long Npad,
                                         __half* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 16) return;
  long long m;
  int c0;
  ce::blk_decode(i * 8, 128, &m, &c0);
  const float sg = __ldg(sigma);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float r = 0.0f;
    if (m < N && c < w) {
      const float d = d_out[m * w + c], o = out[m * w + c];
      r = kind == 0 ? d * ((1.0f - o) * o) : (o > 0.0f ? d : 0.0f);
    }
    v[j] = r * sg;
  }
  store8_hs(dst, i, v);
}

struct RnChainBufs {
  long long Npad;
  __half* F16; __half* X16;
  __half* H16[VDN_MAX_LAYERS];
  uint16_t* stash;
  __half* ZL16; __half* ZB16[VDN_MAX_LAYERS];
  float* sig;
};
static inline long long rn_chain_blob_floats(int L, long long N) {
  return pad128(N) * (128 + 32 + (long long)(L - 1) * 128) + (long long)ce::stash_floats(N);
}
static inline long long rn_chain_ws_floats(int L, long long N) { return pad128(N) * (64 + (long long)(L - 1) * 128) + 32; }
static inline void rn_chain_carve(int L, long long N, float* blob, float* ws, RnChainBufs* b) {
  const long long Np = pad128(N);
  b->Npad = Np;
  if (blob) {
    float* p = blob;
    b->F16 = reinterpret_cast<__half*>(p); p += Np * 128;
    b->X16 = reinterpret_cast<__half*>(p); p += Np * 32;
    for (int l = 0; l < L - 1; ++l) { b->H16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    b->stash = reinterpret_cast<uint16_t*>(p);
  }
  if (ws) {
    float* p = ws;
    b->ZL16 = reinterpret_cast<__half*>(p); p += Np * 64;
    for (int l = 0; l < L - 1; ++l) { b->ZB16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    b->sig = p;
  }
}

static inline int rn_chain_forward(const RnShape& s, const float* packed, const float* points, const float* normals,
                                   const float* view_dirs, const float* feats, int ldf, long long N, float* out,
                                   const RnChainBufs& b, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int L = s.L;
  int e = launch1d(rn_extras16_kernel, b.Npad * 8, st, points, normals, view_dirs, s.multires_view, s.mode, N, b.Npad, b.X16);
  if (e) return e;
  e = launch1d(rows_to_16_kernel, b.Npad * 32, st, feats, ldf, s.F, 1.0f, (const float*)nullptr, N, b.Npad, b.F16, 256);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed; a.stash = b.stash;
  a.a0 = b.X16; a.a0_ld = 64; a.a0_w = 64;
  int P = 0;
  {   // extras part of layer 0 -> scratch; then the feature tile replaces the extras in the slot
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[0], ly.out_ld[0], 0, s.F / 64, ly.out_dim[0], s.nextra);
    p.op = ce::OP_STASH; p.width = ly.out_dim[0]; p.stash_w = 0;
    p.aload = b.F16; p.al_ld = 256; p.al_w = 256;
  }
  for (int l = 0; l < L - 1; ++l) {
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[l], ly.out_ld[l], 0, 0, ly.out_dim[l], l == 0 ? s.F : ly.in_dim[l]);
    p.op = ce::OP_RELU; p.width = ly.out_dim[l]; p.bias_off = ly.off_b[l];
    if (l == 0) p.stash_r = 0;
    p.a_out = 1; p.a_wr = 256;
    p.o16a = b.H16[l]; p.ldo16a = 256;
  }
  {
    const int lo = L - 1;
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[lo], ly.out_ld[lo], 0, 0, ly.out_dim[lo], ly.in_dim[lo]);
    p.op = ce::OP_OUT32; p.width = s.d_out; p.bias_off = ly.off_b[lo];
    p.o32 = out; p.ldo32 = s.d_out; p.o32_c0 = 0; p.o32_w = s.d_out; p.act = s.squeeze_out ? 1 : 2;
  }
  a.P = P;
  return ce::launch(a, st, PROF_CHAIN_TRAIN);
}

static inline int rn_chain_backward(const RnShape& s, const float* packed, long long N, const RnChainBufs& b,
                                    const float* out, const float* d_out, float* dpacked, float* d_cin, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int L = s.L, lo = L - 1;
  int e = launch_sigma(b.sig, st, d_out, N, s.d_out, s.d_out, 1.0f);
  if (e) return e;
  e = launch1d(rn_zlast16_kernel, b.Npad * 16, st, d_out, out, s.d_out, s.squeeze_out ? 0 : 1, (const float*)b.sig, N, b.Npad,
               b.ZL16);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed; a.sigma = b.sig;
  a.a0 = b.ZL16; a.a0_ld = 128; a.a0_w = 128;
  int P = 0;
  for (int l = lo; l >= 1; --l) {      // h-bar_{l-1} = z-bar_l W_l, masked by the sign of h_{l-1}
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[l], ly.in_ld[l], 0, 0, ly.in_dim[l], ly.out_dim[l]);
    p.op = ce::OP_MASK; p.width = ly.out_dim[l - 1];
    p.aux0 = b.H16[l - 1]; p.ld0 = 256;
    p.a_out = (l > 1 || d_cin) ? 1 : 0; p.a_wr = 256;
    p.o16a = b.ZB16[l - 1]; p.ldo16a = 256;
  }
  if (d_cin) {     // cotangent of the input row, [feature | extras] like the packed layer 0
    {
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      ce::set_mma(&p, ly.off_iht[0], ly.in_ld[0], 0, 0, s.F, ly.out_dim[0]);
      p.op = ce::OP_OUT32; p.width = s.F;
      p.o32 = d_cin; p.ldo32 = s.ldIn; p.o32_c0 = 0; p.o32_w = s.F; p.o32_unscale = 1;
    }
    {
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      ce::set_mma(&p, ly.off_iht[0], ly.in_ld[0], s.F, 0, s.nextra, ly.out_dim[0]);
      p.op = ce::OP_OUT32; p.width = s.nextra;
      p.o32 = d_cin + s.F; p.ldo32 = s.ldIn; p.o32_c0 = 0; p.o32_w = s.nextra; p.o32_unscale = 1;
    }
  }
  a.P = P;
  e = ce::launch(a, st, PROF_CHAIN_TRAIN);
  if (e) return e;
  wg::Builder w(N, dpacked, b.sig);
  const int mZL = w.add_x(b.ZL16, 128, s.d_out), mF = w.add_y(b.F16, 256, s.F), mX = w.add_y(b.X16, 64, s.nextra);
  int mH[VDN_MAX_LAYERS], mZ[VDN_MAX_LAYERS];
  for (int l = 0; l < L - 1; ++l) {
    mH[l] = w.add_y(b.H16[l], 256, ly.in_dim[l + 1]);
    mZ[l] = w.add_x(b.ZB16[l], 256, ly.out_dim[l]);
  }
  {
    wg::Job* j = w.add_job(s.d_out, ly.in_dim[lo], ly.off_w[lo], ly.in_ld[lo], 1.0f, ly.off_b[lo], 1.0f);
    wg::Builder::add_seg(j, mZL, 0, mH[lo - 1], 0);
  }
  for (int l = 1; l < L - 1; ++l) {
    wg::Job* j = w.add_job(ly.out_dim[l], ly.in_dim[l], ly.off_w[l], ly.in_ld[l], 1.0f, ly.off_b[l], 1.0f);
    wg::Builder::add_seg(j, mZ[l], 0, mH[l - 1], 0);
  }
  {
    wg::Job* j = w.add_job(ly.out_dim[0], s.F, ly.off_w[0], ly.in_ld[0], 1.0f, ly.off_b[0], 1.0f);
    wg::Builder::add_seg(j, mZ[0], 0, mF, 0);
    wg::Job* j2 = w.add_job(ly.out_dim[0], s.nextra, ly.off_w[0] + s.F, ly.in_ld[0], 1.0f, -1, 1.0f);
    wg::Builder::add_seg(j2, mZ[0], 0, mX, 0);
  }
  return w.launch(st, PROF_WGRAD16);
}

// ------------------------------------------------------------------------------------------------------------
// NeRF++ background field.  Packed layers: 0..D-1 pts_linears (layer skip+1 takes [hidden | embedding]),
// D = [alpha ; feature] (weight images rotated: features first), D+1 = views_linears.0 on [feature | view
// embedding], D+2 = [rgb ; dpt].
// ------------------------------------------------------------------------------------------------------------
struct NerfShape {
  int D, W, d_in, multires, multires_view, skip, rgb_dims, dpt_dim, d_e, d_ev;
  const MlpLayout* ly;
};

// EV16 [Npad, 128]: columns 0 .. d_e-1 the point embedding (zero to 95), 96 .. 96+d_ev-1 the view embedding
static __global__ void nerf_in16_kernel(const float* __restrict__ pts, const float* __restrict__ views, int d_in, int L,
                                        int Lv, long long N, long long Npad, __half* __restrict__ ev16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad * 16) return;
  long long m;
  int c0;
  ce::blk_decode(i * 8, 128, &m, &c0);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float r = 0.0f;
    if (m < N) {
      if (c < 96) r = embed_col(pts + m * d_in, d_in, L, c, 1.0f);
      else r = embed_col(views + m * 3, 3, Lv, c - 96, 1.0f);
    }
    v[j] = r;
  }
  store8_h(ev16, i, v);
}

struct NerfChainBufs {
  long long Npad;
  __half* EV16;
  __half* H16[VDN_MAX_LAYERS];
  __half* FT16; __half* HV16;
  uint16_t* stash;
  __half* ZO16; __half* SG16; __half* ZV16; __half* ZF16;
  __half* ZB16[VDN_MAX_LAYERS];
  float* DVE; float* EE0; float* EE1;
  float* sig;
};
static inline long long nerf_chain_blob_floats(int D, long long N) {
  return pad128(N) * (64 + (long long)D * 128 + 128 + 64) + (long long)ce::stash_floats(N);
}
static inline long long nerf_chain_ws_floats(int D, long long N) {
  return pad128(N) * (64 + 4 + 64 + 128 + (long long)D * 128 + 32 + 96 + 96) + 32;
}
static inline void nerf_chain_carve(int D, long long N, float* blob, float* ws, NerfChainBufs* b) {
  const long long Np = pad128(N);
  b->Npad = Np;
  if (blob) {
    float* p = blob;
    b->EV16 = reinterpret_cast<__half*>(p); p += Np * 64;
    for (int l = 0; l < D; ++l) { b->H16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    b->FT16 = reinterpret_cast<__half*>(p); p += Np * 128;
    b->HV16 = reinterpret_cast<__half*>(p); p += Np * 64;
    b->stash = reinterpret_cast<uint16_t*>(p);
  }
  if (ws) {
    float* p = ws;
    b->ZO16 = reinterpret_cast<__half*>(p); p += Np * 64;
    b->SG16 = reinterpret_cast<__half*>(p); p += Np * 4;
    b->ZV16 = reinterpret_cast<__half*>(p); p += Np * 64;
    b->ZF16 = reinterpret_cast<__half*>(p); p += Np * 128;
    for (int l = 0; l < D; ++l) { b->ZB16[l] = reinterpret_cast<__half*>(p); p += Np * 128; }
    b->DVE = p; p += Np * 32;
    b->EE0 = p; p += Np * 96;
    b->EE1 = p; p += Np * 96;
    b->sig = p;
  }
}

static inline int nerf_chain_forward(const NerfShape& s, const float* packed, const float* pts, const float* views,
                                     long long N, float* sigma, float* rgb, float* dpt, const NerfChainBufs& b,
                                     cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int D = s.D;
  int e = launch1d(nerf_in16_kernel, b.Npad * 16, st, pts, views, s.d_in, s.multires, s.multires_view, N, b.Npad, b.EV16);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed; a.stash = b.stash;
  a.a0 = b.EV16; a.a0_ld = 128; a.a0_w = 128;
  int P = 0;
  if (s.skip >= 0) {     // embedding part of the layer after the skip concat -> scratch 0
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[s.skip + 1], ly.out_ld[s.skip + 1], 0, s.W / 64, s.W, s.d_e, 0);
    p.op = ce::OP_STASH; p.width = s.W; p.stash_w = 0;
  }
  {                      // view-embedding part of the view layer -> scratch 1
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[D + 1], ly.out_ld[D + 1], 0, s.W / 64, s.W / 2, s.d_ev, 48);
    p.op = ce::OP_STASH; p.width = s.W / 2; p.stash_w = 1;
  }
  for (int l = 0; l < D; ++l) {
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[l], ly.out_ld[l], 0, 0, s.W, l == 0 ? s.d_e : s.W, 0);
    p.op = ce::OP_RELU; p.width = s.W; p.bias_off = ly.off_b[l];
    if (l == s.skip + 1 && s.skip >= 0) p.stash_r = 0;
    p.a_out = 1; p.a_wr = 256;
    p.o16a = b.H16[l]; p.ldo16a = 256;
  }
  {   // alpha head: image position W holds output 0 (orot = 1)
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[D], ly.out_ld[D], s.W, 0, 16, s.W);
    p.op = ce::OP_OUT32; p.width = 1; p.bias_off = ly.off_b[D];
    p.o32 = sigma; p.ldo32 = 1; p.o32_c0 = 0; p.o32_w = 1;
  }
  {   // feature head (no activation) -> next operand
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[D], ly.out_ld[D], 0, 0, s.W, s.W);
    p.op = ce::OP_LINEAR; p.width = s.W; p.bias_off = ly.off_b[D] + 1;
    p.a_out = 1; p.a_wr = 256;
    p.o16a = b.FT16; p.ldo16a = 256;
  }
  {   // view layer
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[D + 1], ly.out_ld[D + 1], 0, 0, s.W / 2, s.W);
    p.op = ce::OP_RELU; p.width = s.W / 2; p.bias_off = ly.off_b[D + 1]; p.stash_r = 1;
    p.a_out = 1; p.a_wr = 128;
    p.o16a = b.HV16; p.ldo16a = 128;
  }
  {   // [rgb ; dpt]
    const int no = s.rgb_dims + s.dpt_dim;
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_ih[D + 2], ly.out_ld[D + 2], 0, 0, no, s.W / 2);
    p.op = ce::OP_OUT32; p.width = no; p.bias_off = ly.off_b[D + 2];
    p.o32 = rgb; p.ldo32 = s.rgb_dims; p.o32_c0 = 0; p.o32_w = s.rgb_dims;
    if (s.dpt_dim > 0 && dpt) { p.o32b = dpt; p.ldo32b = s.dpt_dim; p.o32b_c0 = s.rgb_dims; p.o32b_w = s.dpt_dim; }
  }
  a.P = P;
  return ce::launch(a, st, PROF_CHAIN_TRAIN);
}

static inline int nerf_chain_backward(const NerfShape& s, const float* packed, const float* pts, const float* views,
                                      long long N, const NerfChainBufs& b, const float* d_sigma, const float* d_rgb,
                                      const float* d_dpt, float* dpacked, float* d_pts, float* d_views, cudaStream_t st) {
  const MlpLayout& ly = *s.ly;
  const int D = s.D;
  const int no = s.rgb_dims + s.dpt_dim;
  const float* dd = s.dpt_dim > 0 ? d_dpt : nullptr;
  int e = launch_sigma(b.sig, st, d_rgb, d_rgb ? N : 0, s.rgb_dims, s.rgb_dims, 1.0f, dd, dd ? N : 0, s.dpt_dim > 0 ? s.dpt_dim : 1,
                       s.dpt_dim > 0 ? s.dpt_dim : 1, 1.0f, d_sigma, d_sigma ? N : 0, 1, 1, 1.0f);
  if (e) return e;
  e = launch1d(gather2_16_kernel, b.Npad * 16, st, d_rgb, s.rgb_dims, s.rgb_dims, dd, s.dpt_dim, s.dpt_dim, (const float*)b.sig, N,
               b.Npad, b.ZO16, 128);
  if (e) return e;
  e = launch1d(rows_to_16_kernel, b.Npad, st, d_sigma, 1, 1, 1.0f, (const float*)b.sig, N, b.Npad, b.SG16, 8);
  if (e) return e;
  ce::Args a;
  ce::init_args(&a);
  a.N = N; a.packed = packed; a.sigma = b.sig;
  a.a0 = b.ZO16; a.a0_ld = 128; a.a0_w = 128;
  a.row_off[0] = ly.off_w[D]; a.row_len[0] = s.W;       // alpha row of the stacked head
  int P = 0;
  {   // [rgb ; dpt] -> view layer
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[D + 2], ly.in_ld[D + 2], 0, 0, s.W / 2, no);
    p.op = ce::OP_MASK; p.width = s.W / 2;
    p.aux0 = b.HV16; p.ld0 = 128;
    p.a_out = 1; p.a_wr = 128;
    p.o16a = b.ZV16; p.ldo16a = 128;
  }
  if (d_views) {   // view-embedding tail of the view layer's input cotangent
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[D + 1], ly.in_ld[D + 1], s.W, 0, s.d_ev, s.W / 2);
    p.op = ce::OP_OUT32; p.width = s.d_ev;
    p.o32 = b.DVE; p.ldo32 = 32; p.o32_c0 = 0; p.o32_w = s.d_ev; p.o32_unscale = 1;
  }
  {   // view layer -> feature (linear head: no mask)
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[D + 1], ly.in_ld[D + 1], 0, 0, s.W, s.W / 2);
    p.op = ce::OP_MASK; p.width = s.W;
    p.a_out = 1; p.a_wr = 256;
    p.o16a = b.ZF16; p.ldo16a = 256;
  }
  {   // stacked head -> h_{D-1}: features through the MMA, the alpha row as a rank-1 term
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[D], ly.in_ld[D], 0, 0, s.W, s.W);
    p.op = ce::OP_MASK; p.width = s.W;
    if (d_sigma) { p.r1 = d_sigma; p.r1_stride = 1; p.r1_mul = 1.0f; p.r1_row = 0; p.r1_scaled = 1; }
    p.aux0 = b.H16[D - 1]; p.ld0 = 256;
    p.a_out = 1; p.a_wr = 256;
    p.o16a = b.ZB16[D - 1]; p.ldo16a = 256;
  }
  for (int l = D - 1; l >= 1; --l) {
    if (d_pts && s.skip >= 0 && l == s.skip + 1) {   // embedding tail of the skip concat
      ce::Phase& p = a.ph[P++];
      p = ce::make_phase();
      ce::set_mma(&p, ly.off_iht[l], ly.in_ld[l], s.W, 0, s.d_e, s.W);
      p.op = ce::OP_OUT32; p.width = s.d_e;
      p.o32 = b.EE1; p.ldo32 = 96; p.o32_c0 = 0; p.o32_w = s.d_e; p.o32_unscale = 1;
    }
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[l], ly.in_ld[l], 0, 0, s.W, s.W);
    p.op = ce::OP_MASK; p.width = s.W;
    p.aux0 = b.H16[l - 1]; p.ld0 = 256;
    p.a_out = (l > 1 || d_pts) ? 1 : 0; p.a_wr = 256;
    p.o16a = b.ZB16[l - 1]; p.ldo16a = 256;
  }
  if (d_pts) {
    ce::Phase& p = a.ph[P++];
    p = ce::make_phase();
    ce::set_mma(&p, ly.off_iht[0], ly.in_ld[0], 0, 0, s.d_e, s.W);
    p.op = ce::OP_OUT32; p.width = s.d_e;
    p.o32 = b.EE0; p.ldo32 = 96; p.o32_c0 = 0; p.o32_w = s.d_e; p.o32_unscale = 1;
  }
  a.P = P;
  e = ce::launch(a, st, PROF_CHAIN_TRAIN);
  if (e) return e;
  // ---- weight / bias gradients ----
  wg::Builder w(N, dpacked, b.sig);
  const int mZO = w.add_x(b.ZO16, 128, no), mSG = w.add_x(b.SG16, 8, 1), mZV = w.add_x(b.ZV16, 128, s.W / 2);
  const int mZF = w.add_x(b.ZF16, 256, s.W), mFT = w.add_y(b.FT16, 256, s.W), mHV = w.add_y(b.HV16, 128, s.W / 2);
  const int mEVe = w.add_y(b.EV16, 128, s.d_e), mEVv = w.add_y(b.EV16, 128, s.d_ev);   // point / view embedding parts
  int mH[VDN_MAX_LAYERS], mZ[VDN_MAX_LAYERS];
  for (int l = 0; l < D; ++l) { mH[l] = w.add_y(b.H16[l], 256, s.W); mZ[l] = w.add_x(b.ZB16[l], 256, s.W); }
  {   // [rgb ; dpt]
    wg::Job* j = w.add_job(no, s.W / 2, ly.off_w[D + 2], ly.in_ld[D + 2], 1.0f, ly.off_b[D + 2], 1.0f);
    wg::Builder::add_seg(j, mZO, 0, mHV, 0);
  }
  {   // view layer: feature columns, view-embedding columns
    wg::Job* j = w.add_job(s.W / 2, s.W, ly.off_w[D + 1], ly.in_ld[D + 1], 1.0f, ly.off_b[D + 1], 1.0f);
    wg::Builder::add_seg(j, mZV, 0, mFT, 0);
    wg::Job* j2 = w.add_job(s.W / 2, s.d_ev, ly.off_w[D + 1] + s.W, ly.in_ld[D + 1], 1.0f, -1, 1.0f);
    wg::Builder::add_seg(j2, mZV, 0, mEVv, 96);
  }
  {   // stacked head: feature rows 1 .. W, alpha row 0
    wg::Job* j = w.add_job(s.W, s.W, ly.off_w[D] + ly.in_ld[D], ly.in_ld[D], 1.0f, ly.off_b[D] + 1, 1.0f);
    wg::Builder::add_seg(j, mZF, 0, mH[D - 1], 0);
    if (d_sigma) {
      wg::Job* j2 = w.add_job(1, s.W, ly.off_w[D], ly.in_ld[D], 1.0f, ly.off_b[D], 1.0f);
      wg::Builder::add_seg(j2, mSG, 0, mH[D - 1], 0);
    }
  }
  for (int l = D - 1; l >= 0; --l) {
    if (l == 0) {
      wg::Job* j = w.add_job(s.W, s.d_e, ly.off_w[0], ly.in_ld[0], 1.0f, ly.off_b[0], 1.0f);
      wg::Builder::add_seg(j, mZ[0], 0, mEVe, 0);
    } else {
      wg::Job* j = w.add_job(s.W, s.W, ly.off_w[l], ly.in_ld[l], 1.0f, ly.off_b[l], 1.0f);
      wg::Builder::add_seg(j, mZ[l], 0, mH[l - 1], 0);
      if (s.skip >= 0 && l == s.skip + 1) {
        wg::Job* j2 = w.add_job(s.W, s.d_e, ly.off_w[l] + s.W, ly.in_ld[l], 1.0f, -1, 1.0f);
        wg::Builder::add_seg(j2, mZ[l], 0, mEVe, 0);
      }
    }
  }
  e = w.launch(st, PROF_WGRAD16);
  if (e) return e;
  if (d_pts) {
    const long long tot = N * s.d_in;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, pts, s.d_in, N, s.d_in, s.multires, 1.0f, b.EE0,
               96, s.skip >= 0 ? b.EE1 : nullptr, 96, 1.0f, 1.0f, d_pts, s.d_in, 0);
  }
  if (d_views) {
    const long long tot = N * 3;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, views, 3, N, 3, s.multires_view, 1.0f, b.DVE, 32,
               nullptr, 0, 0.0f, 1.0f, d_views, 3, 0);
  }
  return (int)(cudaError_t)::vdn::take_launch_error();
}

}  // namespace vdn

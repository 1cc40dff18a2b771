// ReLU MLPs of the path, exact-fp32 layer-wise mode:
//   * RenderingNetwork (colour head d_out=3, depth-feature head d_out=96)   reference fields.py:112-176
//   * NeRF++ background field                                             reference fields.py:264-355
// Post-activation tensors H_l are stored (ReLU backward only needs the sign), heads that share an input are
// packed as one stacked weight so they run as one GEMM with a splitting epilogue.
#include "gemm_tn_tc.cuh"
#include "relu_chains.cuh"
#include "pointwise.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {

// ------------------------------------------------------------------------------------------------
// RenderingNetwork
// ------------------------------------------------------------------------------------------------
struct RnCfg {
  int d_feature, mode, d_out, d_hidden, n_layers, multires_view, squeeze_out;
  int L, in0, ldIn, ldH;
  MlpLayout ly;
};

static int parse_rn_cfg(const int* cfg, RnCfg* c) {
  c->d_feature = cfg[0]; c->mode = cfg[1]; c->d_out = cfg[2]; c->d_hidden = cfg[3]; c->n_layers = cfg[4];
  c->multires_view = cfg[5]; c->squeeze_out = cfg[6];
  c->L = c->n_layers + 1;
  if (c->L < 2 || c->L > VDN_MAX_LAYERS || c->mode < 0 || c->mode > 2 || c->multires_view < 0) return 1;
  int nview = (c->mode != 1) ? 3 * (1 + 2 * c->multires_view) : 0;
  int nnrm = (c->mode != 2) ? 3 : 0;
  c->in0 = 3 + nview + nnrm + c->d_feature;
  int in_dims[VDN_MAX_LAYERS], out_dims[VDN_MAX_LAYERS];
  for (int l = 0; l < c->L; ++l) {
    in_dims[l] = l == 0 ? c->in0 : c->d_hidden;
    out_dims[l] = l == c->L - 1 ? c->d_out : c->d_hidden;
  }
  c->ldIn = round_up(c->in0, 16);
  c->ldH = round_up(c->d_hidden, 16);
  return make_layout(c->L, in_dims, out_dims, &c->ly);
}

// Fused chains (relu_chains.cuh) for the shipped shape: 256 feature inputs, 256-wide hidden layers, at most 48 extra
// input columns, at most 128 outputs.
static bool rn_chain_ok(const RnCfg& c) {
  if (g_mode != 1 || !g_chain) return false;
  const int nextra = c.in0 - c.d_feature;
  if (c.d_feature != 256 || c.d_hidden != 256 || nextra < 1 || nextra > 48 || c.d_out > 128 || c.L < 2 || c.L > 10) return false;
  return wg::encode_fn() != nullptr;
}
static RnShape rn_shape(const RnCfg& c) {
  return RnShape{c.L, c.d_feature, c.in0 - c.d_feature, c.d_out, c.mode, c.multires_view, c.squeeze_out, c.ldIn, &c.ly};
}
}  // namespace vdn
using namespace vdn;

extern "C" int vdn_rendernet_layer_dims(const int* cfg, int* in_dims, int* out_dims) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return -1;
  for (int l = 0; l < c.L; ++l) { in_dims[l] = c.ly.in_dim[l]; out_dims[l] = c.ly.out_dim[l]; }
  return c.L;
}

// saved blob: CIN [N, ldIn] | H_0 .. H_{L-2} [N, ldH]
extern "C" long long vdn_rendernet_blob_floats(const int* cfg, long long N) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return -1;
  return rn_chain_ok(c) ? rn_chain_blob_floats(c.L, N) : N * c.ldIn + (long long)(c.L - 1) * N * c.ldH;
}

// Input-column rotation of the packed layers (layer 0 is stored as [feature | extras]); returns the number of layers.
extern "C" int vdn_rendernet_layer_rot(const int* cfg, int* rot) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return -1;
  for (int l = 0; l < c.L; ++l) rot[l] = 0;
  rot[0] = c.in0 - c.d_feature;
  return c.L;
}

extern "C" int vdn_rendernet_forward(const int* cfg, const float* packed, const float* points, const float* normals,
                                     const float* view_dirs, const float* feats, int ldf, long long N, float* out,
                                     float* blob, void* stream) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rn_chain_ok(c)) {
    RnChainBufs cb;
    rn_chain_carve(c.L, N, blob, nullptr, &cb);
    return rn_chain_forward(rn_shape(c), packed, points, normals, view_dirs, feats, ldf, N, out, cb, st);
  }
  float* CIN = blob;
  float* H = blob + N * c.ldIn;
  long long threads = N * 32;
  VDN_LAUNCH(rendernet_input_kernel, (unsigned)((threads + 255) / 256), 256, 0, st, points, normals, view_dirs, feats, ldf,
                                                                           c.d_feature, c.multires_view, c.mode, N,
                                                                           CIN, c.ldIn);
  int e = (int)(cudaError_t)::vdn::take_launch_error();
  if (e) return e;
  for (int l = 0; l < c.L; ++l) {
    Operand A = (l == 0) ? make_operand(CIN, c.ldIn, c.ldIn, c.in0)
                         : make_operand(H + (long long)(l - 1) * N * c.ldH, c.ldH, c.ly.in_ld[l], c.ly.in_dim[l]);
    if (l > 0) A.rounded = 1;
    Epilogue E;
    if (l == c.L - 1)
      E = make_epilogue(c.squeeze_out ? EPI_SIGMOID : EPI_RELU, packed + c.ly.off_b[l], out, c.d_out);
    else
      E = make_epilogue(EPI_RELU, packed + c.ly.off_b[l], H + (long long)l * N * c.ldH, c.ldH), E.round_c = 1;
    e = launch_gemm_nt((int)N, c.ly.out_dim[l], c.ly.in_ld[l], A, wref(c.ly, packed, l), E, st);
    if (e) return e;
  }
  return 0;
}

extern "C" long long vdn_rendernet_bwd_ws_floats(const int* cfg, long long N) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return -1;
  long long S = wgrad_max_splits(N);
  long long maxw = 0, maxo = 0;
  for (int l = 0; l < c.L; ++l) {
    long long w = (long long)c.ly.out_dim[l] * ((c.ly.in_dim[l] + 3) & ~3);
    if (w > maxw) maxw = w;
    if (c.ly.out_ld[l] > maxo) maxo = c.ly.out_ld[l];
  }
  return rn_chain_ok(c) ? rn_chain_ws_floats(c.L, N) : 2 * N * c.ldH + N * c.ly.out_ld[c.L - 1] + S * maxw + 256 * maxo + 64;
}

// d_out: [N, d_out] contiguous cotangent of the network output; `out` is the forward output.
// d_cin (nullable): [N, ldIn] cotangent of the assembled input row (caller slices points/normals/feature/view).
extern "C" int vdn_rendernet_backward(const int* cfg, const float* packed, long long N, const float* blob,
                                      const float* out, const float* d_out, float* dpacked, float* d_cin, float* ws,
                                      void* stream) {
  RnCfg c;
  if (parse_rn_cfg(cfg, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rn_chain_ok(c)) {
    RnChainBufs cb;
    rn_chain_carve(c.L, N, const_cast<float*>(blob), ws, &cb);
    return rn_chain_backward(rn_shape(c), packed, N, cb, out, d_out, dpacked, d_cin, st);
  }
  const MlpLayout& ly = c.ly;
  const int L = c.L, M = (int)N;
  const float* CIN = blob;
  const float* H = blob + N * c.ldIn;
  float* ZB[2] = {ws, ws + N * c.ldH};
  float* ZL = ws + 2 * N * c.ldH;
  float* partials = ZL + N * ly.out_ld[L - 1];
  // zbar_last = d_out * act'(out), padded to out_ld
  {
    long long tot = N * ly.out_ld[L - 1];
    VDN_LAUNCH(act_backward_pad_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, d_out, out, c.d_out, N, ZL,
                                                                          ly.out_ld[L - 1], c.squeeze_out ? 0 : 1);
    int e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
  }
  for (int l = L - 1; l >= 0; --l) {
    Operand zbar = (l == L - 1) ? make_operand(ZL, ly.out_ld[l], ly.out_ld[l], ly.out_dim[l])
                                : make_operand(ZB[l & 1], c.ldH, ly.out_ld[l], ly.out_dim[l]);
    Operand u = (l == 0) ? make_operand(CIN, c.ldIn, c.ldIn, c.in0)
                         : make_operand(H + (long long)(l - 1) * N * c.ldH, c.ldH, ly.in_ld[l], ly.in_dim[l]);
    if (l < L - 1) zbar.rounded = 1;
    if (l > 0) u.rounded = 1;
    int e = launch_wgrad_any(M, ly.out_dim[l], ly.in_dim[l], zbar, u, partials, dpacked + ly.off_w[l], ly.in_ld[l], 1,
                             dpacked + ly.off_b[l], st);
    if (e) return e;
    if (l > 0) {
      Epilogue E = make_epilogue(EPI_RELU_MASK, nullptr, ZB[(l - 1) & 1], c.ldH);
      E.round_c = 1;
      E.aux = H + (long long)(l - 1) * N * c.ldH; E.ldaux = c.ldH; E.split = 0;
      e = launch_gemm_nt(M, ly.in_dim[l], ly.out_ld[l], zbar, wtref(ly, packed, l), E, st);
      if (e) return e;
    } else if (d_cin) {
      Epilogue E = make_epilogue(EPI_STORE, nullptr, d_cin, c.ldIn);
      e = launch_gemm_nt(M, c.in0, ly.out_ld[0], zbar, wtref(ly, packed, 0), E, st);
      if (e) return e;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NeRF++ background field.  Packed layer list:
//   0..D-1 : pts_linears          D   : [alpha_linear ; feature_linear] stacked (1 + W rows)
//   D+1    : views_linears.0      D+2 : [rgb_linear ; dpt_linear] stacked (3 [+ dpt_dim] rows)
// ------------------------------------------------------------------------------------------------
namespace vdn {
struct NerfCfg {
  int D, W, d_in, d_in_view, multires, multires_view, skip, rgb_dims, dpt_dim;
  int L, d_e, d_ev, ldE, ldH, ldU, ldV, ldHV, vin;
  MlpLayout ly;
};

static int parse_nerf_cfg(const int* cfg, NerfCfg* c) {
  c->D = cfg[0]; c->W = cfg[1]; c->d_in = cfg[2]; c->d_in_view = cfg[3]; c->multires = cfg[4];
  c->multires_view = cfg[5]; c->skip = cfg[6]; c->rgb_dims = cfg[7]; c->dpt_dim = cfg[8];
  c->L = c->D + 3;
  if (c->D < 2 || c->L > VDN_MAX_LAYERS || c->d_in < 1 || c->d_in > 4 || c->d_in_view != 3) return 1;
  if (c->skip >= 0 && (c->skip < 0 || c->skip > c->D - 2)) return 1;
  c->d_e = c->d_in * (1 + 2 * c->multires);
  c->d_ev = c->d_in_view * (1 + 2 * c->multires_view);
  int in_dims[VDN_MAX_LAYERS], out_dims[VDN_MAX_LAYERS];
  for (int i = 0; i < c->D; ++i) {
    in_dims[i] = (i == 0) ? c->d_e : (i - 1 == c->skip ? c->W + c->d_e : c->W);
    out_dims[i] = c->W;
  }
  in_dims[c->D] = c->W; out_dims[c->D] = 1 + c->W;
  c->vin = c->W + c->d_ev;
  in_dims[c->D + 1] = c->vin; out_dims[c->D + 1] = c->W / 2;
  in_dims[c->D + 2] = c->W / 2; out_dims[c->D + 2] = c->rgb_dims + c->dpt_dim;
  c->ldE = round_up(c->d_e, 16);
  c->ldH = round_up(c->W, 16);
  c->ldU = round_up(c->W + c->d_e, 16);
  c->ldV = round_up(c->vin, 16);
  c->ldHV = round_up(c->W / 2, 16);
  return make_layout(c->L, in_dims, out_dims, &c->ly);
}

// saved blob: E [N,ldE] | VIN [N,ldV] | U [N,ldU] (skip only) | H_0..H_{D-1} [N,ldH] | HV [N,ldHV]
struct NerfBlob {
  float* E; float* VIN; float* U; float* H[VDN_MAX_LAYERS]; float* HV;
};
static long long nerf_blob_floats(const NerfCfg& c, long long N) {
  return N * c.ldE + N * c.ldV + (c.skip >= 0 ? N * c.ldU : 0) + (long long)c.D * N * c.ldH + N * c.ldHV;
}
static void carve_nerf(const NerfCfg& c, long long N, float* p, NerfBlob* b) {
  b->E = p; p += N * c.ldE;
  b->VIN = p; p += N * c.ldV;
  b->U = nullptr;
  if (c.skip >= 0) { b->U = p; p += N * c.ldU; }
  for (int i = 0; i < c.D; ++i) { b->H[i] = p; p += N * c.ldH; }
  b->HV = p;
}
static int nerf_orot_head(const NerfCfg& c) { return (c.W & 7) == 0 ? 1 : 0; }
static bool nerf_chain_ok(const NerfCfg& c) {
  if (g_mode != 1 || !g_chain) return false;
  if (c.W != 256 || c.d_e > 96 || c.d_ev > 32 || c.D < 2 || c.D > 8 || c.rgb_dims + c.dpt_dim > 128 || c.rgb_dims < 1) return false;
  if (c.skip >= 0 && c.skip > c.D - 2) return false;
  return wg::encode_fn() != nullptr;
}
static NerfShape nerf_shape(const NerfCfg& c) {
  return NerfShape{c.D, c.W, c.d_in, c.multires, c.multires_view, c.skip, c.rgb_dims, c.dpt_dim, c.d_e, c.d_ev, &c.ly};
}
// The input operand of pts layer i and where its post-ReLU output lives (buffer, ld, column offset).
static Operand nerf_input(const NerfCfg& c, const NerfBlob& b, int i) {
  if (i == 0) return make_operand(b.E, c.ldE, c.ldE, c.d_e);
  if (i - 1 == c.skip) return make_operand(b.U, c.ldU, c.ldU, c.W + c.d_e);
  Operand o = make_operand(b.H[i - 1], c.ldH, c.ly.in_ld[i], c.W);
  o.rounded = 1;   // written by an EPI_RELU epilogue with round_c
  return o;
}
}  // namespace vdn

extern "C" int vdn_nerf_layer_dims(const int* cfg, int* in_dims, int* out_dims) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return -1;
  for (int l = 0; l < c.L; ++l) { in_dims[l] = c.ly.in_dim[l]; out_dims[l] = c.ly.out_dim[l]; }
  return c.L;
}

extern "C" long long vdn_nerf_blob_floats(const int* cfg, long long N) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return -1;
  return nerf_chain_ok(c) ? nerf_chain_blob_floats(c.D, N) : nerf_blob_floats(c, N);
}

// Output rotation per packed layer for vdn_mlp_pack (the stacked [alpha ; feature] head presents its features first).
extern "C" int vdn_nerf_layer_orot(const int* cfg, int* orot) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return -1;
  for (int l = 0; l < c.L; ++l) orot[l] = 0;
  orot[c.D] = nerf_orot_head(c);
  return c.L;
}

extern "C" int vdn_nerf_forward(const int* cfg, const float* packed, const float* pts, const float* views,
                                long long N, float* sigma, float* rgb, float* dpt, float* blob, void* stream) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (nerf_chain_ok(c)) {
    NerfChainBufs cb;
    nerf_chain_carve(c.D, N, blob, nullptr, &cb);
    return nerf_chain_forward(nerf_shape(c), packed, pts, views, N, sigma, rgb, dpt, cb, st);
  }
  const MlpLayout& ly = c.ly;
  NerfBlob b;
  carve_nerf(c, N, blob, &b);
  const int M = (int)N, D = c.D;
  // pts embedding -> E and the tail of the skip buffer U.  The reference concatenates [input_pts, h]
  // (fields.py:334-335); the packed weight of the next layer has its columns rotated so U is [h | input_pts].
  int e = launch_embed_rows(pts, c.d_in, N, c.d_in, c.multires, 1.0f, b.E, c.ldE, b.U, c.ldU, c.W, 1.0f, c.ldU, st);
  if (e) return e;
  // view embedding -> tail of VIN (reference fields.py:340: cat[feature, input_views])
  e = launch_embed_rows(views, c.d_in_view, N, c.d_in_view, c.multires_view, 1.0f, nullptr, 0, b.VIN, c.ldV, c.W, 1.0f,
                        c.ldV, st);
  if (e) return e;
  for (int i = 0; i < D; ++i) {
    Operand A = nerf_input(c, b, i);
    Epilogue E = make_epilogue(EPI_RELU, packed + ly.off_b[i], b.H[i], c.ldH);
    E.round_c = 1;
    if (i == c.skip) { E.c = b.U; E.ldc = c.ldU; E.coff = 0; }
    e = launch_gemm_nt(M, c.W, ly.in_ld[i], A, wref(ly, packed, i), E, st);
    if (e) return e;
  }
  // heads on h_{D-1}: sigma (row 0) and the feature (rows 1..W) into VIN[:, :W]
  {
    Operand A = nerf_input(c, b, D);
    Epilogue E = make_epilogue(EPI_SPLIT, packed + ly.off_b[D], b.VIN, c.ldV);
    E.c2 = sigma; E.ldc2 = 1; E.split = 1; E.scale = 1.0f;
    e = launch_gemm_nt(M, 1 + c.W, ly.in_ld[D], A, wref(ly, packed, D), E, st);
    if (e) return e;
  }
  {
    Operand A = make_operand(b.VIN, c.ldV, c.ldV, c.vin);
    Epilogue E = make_epilogue(EPI_RELU, packed + ly.off_b[D + 1], b.HV, c.ldHV);
    E.round_c = 1;
    e = launch_gemm_nt(M, c.W / 2, ly.in_ld[D + 1], A, wref(ly, packed, D + 1), E, st);
    if (e) return e;
  }
  {
    Operand A = make_operand(b.HV, c.ldHV, c.ldHV, c.W / 2);
    A.rounded = 1;
    Epilogue E = make_epilogue(EPI_SPLIT, packed + ly.off_b[D + 2], dpt, c.dpt_dim);
    E.c2 = rgb; E.ldc2 = c.rgb_dims; E.split = c.rgb_dims; E.scale = 1.0f;
    e = launch_gemm_nt(M, c.rgb_dims + c.dpt_dim, ly.in_ld[D + 2], A, wref(ly, packed, D + 2), E, st);
    if (e) return e;
  }
  return 0;
}

extern "C" long long vdn_nerf_bwd_ws_floats(const int* cfg, long long N) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return -1;
  long long S = wgrad_max_splits(N);
  long long maxw = 0, maxo = 0;
  for (int l = 0; l < c.L; ++l) {
    long long w = (long long)c.ly.out_dim[l] * ((c.ly.in_dim[l] + 3) & ~3);
    if (w > maxw) maxw = w;
    if (c.ly.out_ld[l] > maxo) maxo = c.ly.out_ld[l];
  }
  if (nerf_chain_ok(c)) return nerf_chain_ws_floats(c.D, N);
  return 2 * N * c.ldH + N * c.ly.out_ld[c.D] + N * c.ldHV + N * c.ly.out_ld[c.D + 2] + N * c.ldV + 2 * N * c.ldE + S * maxw +
         256 * maxo + 64;
}

// d_pts (nullable): [N, d_in]; d_views (nullable): [N, 3].
extern "C" int vdn_nerf_backward(const int* cfg, const float* packed, const float* pts, const float* views,
                                 long long N, const float* blob, const float* d_sigma, const float* d_rgb,
                                 const float* d_dpt, float* dpacked, float* d_pts, float* d_views, float* ws,
                                 void* stream) {
  NerfCfg c;
  if (parse_nerf_cfg(cfg, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (nerf_chain_ok(c)) {
    NerfChainBufs cb;
    nerf_chain_carve(c.D, N, const_cast<float*>(blob), ws, &cb);
    return nerf_chain_backward(nerf_shape(c), packed, pts, views, N, cb, d_sigma, d_rgb, d_dpt, dpacked, d_pts, d_views, st);
  }
  const MlpLayout& ly = c.ly;
  NerfBlob b;
  carve_nerf(c, N, const_cast<float*>(blob), &b);
  const int M = (int)N, D = c.D;
  float* p = ws;
  float* ZB[2] = {p, p + N * c.ldH}; p += 2 * N * c.ldH;
  float* ZHEAD = p; p += N * ly.out_ld[D];        // [d_sigma | d_feature]
  float* ZV = p; p += N * c.ldHV;                  // zbar of the view layer
  float* ZO = p; p += N * ly.out_ld[D + 2];        // [d_rgb | d_dpt]
  float* DVIN = p; p += N * c.ldV;                 // cotangent of the view-embedding tail (only for d_views)
  float* EE0 = p; p += N * c.ldE;
  float* EE1 = p; p += N * c.ldE;
  float* partials = p;
  int e;
  auto wg = [&](int l, const Operand& zbar, const Operand& u) -> int {
    return launch_wgrad_any(M, ly.out_dim[l], ly.in_dim[l], zbar, u, partials, dpacked + ly.off_w[l], ly.in_ld[l], 1,
                            dpacked + ly.off_b[l], st);
  };
  // output heads
  {
    int ldo = ly.out_ld[D + 2];
    e = launch_gather2_rows(d_rgb, c.rgb_dims, c.rgb_dims, 1.0f, c.dpt_dim > 0 ? d_dpt : nullptr, c.dpt_dim, c.dpt_dim, 1.0f,
                            N, ZO, ldo, st);
    if (e) return e;
    Operand zo = make_operand(ZO, ldo, ldo, ly.out_dim[D + 2]);
    Operand hv = make_operand(b.HV, c.ldHV, c.ldHV, c.W / 2);
    hv.rounded = 1;
    e = wg(D + 2, zo, hv);
    if (e) return e;
    Epilogue E = make_epilogue(EPI_RELU_MASK, nullptr, ZV, c.ldHV);
    E.round_c = 1;
    E.aux = b.HV; E.ldaux = c.ldHV;
    e = launch_gemm_nt(M, c.W / 2, ldo, zo, wtref(ly, packed, D + 2), E, st);
    if (e) return e;
  }
  // view layer
  {
    Operand zv = make_operand(ZV, c.ldHV, c.ldHV, c.W / 2);
    zv.rounded = 1;
    Operand vin = make_operand(b.VIN, c.ldV, c.ldV, c.vin);
    e = wg(D + 1, zv, vin);
    if (e) return e;
    // cotangent of the feature -> columns 1..W of ZHEAD; column 0 = d_sigma
    int ldh = ly.out_ld[D];
    {
      const long long tot = N * (1 + ldh - (1 + c.W));
      VDN_LAUNCH(head_edges_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, d_sigma, N, ZHEAD, ldh, 1 + c.W);
    }
    e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
    Epilogue E = make_epilogue(EPI_STORE, nullptr, ZHEAD, ldh);
    E.coff = 1;
    e = launch_gemm_nt(M, c.W, c.ldHV, zv, wtref(ly, packed, D + 1), E, st);
    if (e) return e;
    if (d_views) {
      Epilogue E2 = make_epilogue(EPI_STORE, nullptr, DVIN, c.ldV);
      e = launch_gemm_nt(M, c.d_ev, c.ldHV, zv, wtref(ly, packed, D + 1, c.W), E2, st);
      if (e) return e;
    }
  }
  // stacked alpha / feature head -> zbar_{D-1}
  {
    int ldh = ly.out_ld[D];
    Operand zh = make_operand(ZHEAD, ldh, ldh, ly.out_dim[D]);
    e = wg(D, zh, nerf_input(c, b, D));
    if (e) return e;
    Epilogue E = make_epilogue(EPI_RELU_MASK, nullptr, ZB[(D - 1) & 1], c.ldH);
    E.round_c = 1;
    E.aux = b.H[D - 1]; E.ldaux = c.ldH;
    e = launch_gemm_nt(M, c.W, ldh, zh, wtref(ly, packed, D), E, st);
    if (e) return e;
  }
  for (int i = D - 1; i >= 0; --i) {
    Operand zbar = make_operand(ZB[i & 1], c.ldH, c.ldH, c.W);
    zbar.rounded = 1;
    e = wg(i, zbar, nerf_input(c, b, i));
    if (e) return e;
    if (i > 0) {
      // hidden part of the input cotangent, masked by the sign of h_{i-1}
      const bool after_skip = (i - 1 == c.skip);
      Epilogue E = make_epilogue(EPI_RELU_MASK, nullptr, ZB[(i - 1) & 1], c.ldH);
      E.round_c = 1;
      if (after_skip) { E.aux = b.U; E.ldaux = c.ldU; E.split = 0; }
      else { E.aux = b.H[i - 1]; E.ldaux = c.ldH; E.split = 0; }
      e = launch_gemm_nt(M, c.W, c.ldH, zbar, wtref(ly, packed, i), E, st);
      if (e) return e;
      if (after_skip && d_pts) {
        Epilogue E2 = make_epilogue(EPI_STORE, nullptr, EE1, c.ldE);
        e = launch_gemm_nt(M, c.d_e, c.ldH, zbar, wtref(ly, packed, i, c.W), E2, st);
        if (e) return e;
      }
    } else if (d_pts) {
      Epilogue E2 = make_epilogue(EPI_STORE, nullptr, EE0, c.ldE);
      e = launch_gemm_nt(M, c.d_e, c.ldH, zbar, wtref(ly, packed, 0), E2, st);
      if (e) return e;
    }
  }
  if (d_pts) {
    long long tot = N * c.d_in;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, pts, c.d_in, N, c.d_in, c.multires, 1.0f, EE0,
                                                                   c.ldE, c.skip >= 0 ? EE1 : nullptr, c.ldE, 1.0f,
                                                                   1.0f, d_pts, c.d_in, 0);
  }
  if (d_views) {
    long long tot = N * c.d_in_view;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, views, c.d_in_view, N, c.d_in_view,
                                                                   c.multires_view, 1.0f, DVIN, c.ldV, nullptr, 0,
                                                                   0.0f, 1.0f, d_views, c.d_in_view, 0);
  }
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// SDF network (8x256 softplus(beta=100) MLP with skip connection, reference dpt_models/fields.py:9-108).
// Three passes, each a sequence of GEMM launches with fused prologues / epilogues (FFMA kernels in the exact-fp32
// mode, tcgen05 kernels in the tensor-core mode):
//
//   vdn_sdf_forward   value (+ feature) forward                           fields.py:72-92
//   vdn_sdf_normals   analytic reverse-mode input gradient d sdf / d x    fields.py:97-108 (replaces autograd.grad)
//   vdn_sdf_backward  hand-derived backward of (sdf, feature, normals)    replaces autograd double backward
//
// In the tensor-core mode the forward runs through the fused chain kernel of sdf_chain_tc.cuh: all layers for a
// value-only query (and the grid query), layers 0..L-2 with stored pre-activations for the training forward, whose
// multi-head last layer stays a layer-wise launch.
//
// Math: SURVEY.md Appendix A.  Storage: one pre-activation tensor Z_l per layer (activations are recomputed
// in the consumer's operand prologue), plus G_l = d sdf / d(input of layer l) from the normals pass.
#include "gemm_tn_tc.cuh"
#include "sdf_chain_tc.cuh"
#include "sdf_chains.cuh"
#include "pointwise.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {

struct SdfCfg {
  int d_in, multires, d_hidden, n_layers, d_out, skip;
  float scale;
  int L, d_e, ldE, ldH;
  MlpLayout ly;
};

static int parse_sdf_cfg(const int* cfg, float scale, SdfCfg* c) {
  c->d_in = cfg[0]; c->multires = cfg[1]; c->d_hidden = cfg[2]; c->n_layers = cfg[3]; c->d_out = cfg[4];
  c->skip = cfg[5];
  c->scale = scale;
  c->L = c->n_layers + 1;
  if (c->d_in < 1 || c->d_in > 4 || c->multires < 0 || c->multires > 16 || c->L < 2 || c->L > VDN_MAX_LAYERS)
    return 1;
  c->d_e = c->d_in * (1 + 2 * c->multires);
  if (c->skip >= 0 && (c->skip < 1 || c->skip > c->L - 2)) return 1;
  if (c->skip >= 0 && c->d_hidden <= c->d_e) return 1;
  int in_dims[VDN_MAX_LAYERS], out_dims[VDN_MAX_LAYERS];
  for (int l = 0; l < c->L; ++l) {
    in_dims[l] = (l == 0) ? c->d_e : c->d_hidden;
    int o = (l == c->L - 1) ? c->d_out : c->d_hidden;
    if (l + 1 == c->skip) o -= c->d_e;
    out_dims[l] = o;
  }
  c->ldE = round_up(c->d_e, 16);
  c->ldH = round_up(c->d_hidden, 16);
  return make_layout(c->L, in_dims, out_dims, &c->ly);
}

// Output rotation of the fp16 weight images (mlp_layout.cuh): the stacked [sdf ; feature] last layer presents its
// features first when the feature count is a multiple of 8 (so that the scalar row starts an 8-row swizzle atom).
static int sdf_orot_last(const SdfCfg& c) { return (c.d_out > 1 && ((c.d_out - 1) & 7) == 0) ? 1 : 0; }

// The fused training chains (sdf_chains.cuh) cover the shipped shape: 3-D points, 256-wide hidden layers, an embedding
// of at most 64 columns, 256 feature outputs, one optional skip connection; anything else runs layer-wise.
static bool sdf_chain_ok(const SdfCfg& c) {
  if (g_mode != 1 || !g_chain) return false;
  if (c.d_in != 3 || c.d_hidden != 256 || c.d_e > 48 || c.d_out != 257 || c.L < 3 || c.L > 12) return false;
  if (c.skip >= 0 && c.ly.out_dim[c.skip - 1] + c.d_e != 256) return false;
  return wg::encode_fn() != nullptr;
}
static SdfShape sdf_shape(const SdfCfg& c) { return SdfShape{c.L, c.skip, c.d_e, c.multires, c.scale, &c.ly}; }

struct SdfBlob {
  float* E;
  float* U;
  float* Z[VDN_MAX_LAYERS];
};

static long long sdf_blob_floats(const SdfCfg& c, long long N, int save) {
  int nz = save ? c.L - 1 : (c.L - 1 < 2 ? c.L - 1 : 2);
  return N * c.ldE + (c.skip >= 0 ? N * c.ldH : 0) + (long long)nz * N * c.ldH;
}
static void carve_blob(const SdfCfg& c, long long N, int save, float* blob, SdfBlob* b) {
  b->E = blob;
  blob += N * c.ldE;
  b->U = nullptr;
  if (c.skip >= 0) {
    b->U = blob;
    blob += N * c.ldH;
  }
  for (int l = 0; l < c.L - 1; ++l) b->Z[l] = blob + (long long)(save ? l : (l & 1)) * N * c.ldH;
}

struct SdfBlobG {
  float* G[VDN_MAX_LAYERS];  // G[l], 1 <= l <= L-2
  float* DE;
};
static long long sdf_blobg_floats(const SdfCfg& c, long long N) {
  return (long long)(c.L - 2) * N * c.ldH + N * c.ldE;
}
static void carve_blobg(const SdfCfg& c, long long N, float* blob, SdfBlobG* g) {
  for (int l = 1; l <= c.L - 2; ++l) g->G[l] = blob + (long long)(l - 1) * N * c.ldH;
  g->DE = blob + (long long)(c.L - 2) * N * c.ldH;
}

// d sdf / d h_l (the post-activation output of layer l), as (pointer, ld, scale): either the broadcast
// first row of the last layer's weight or the stored G_{l+1} (scaled by 1/sqrt2 through the skip concat).
struct GinRef {
  const float* p;
  int ld;
  float scale;
};
static GinRef gin_of(const SdfCfg& c, const float* packed, const SdfBlobG& g, int l) {
  GinRef r;
  if (l == c.L - 2) {
    r.p = packed + c.ly.off_w[c.L - 1];
    r.ld = 0;
    r.scale = 1.0f;
  } else {
    r.p = g.G[l + 1];
    r.ld = c.ldH;
    r.scale = (l + 1 == c.skip) ? kInvSqrt2 : 1.0f;
  }
  return r;
}

// Input operand u_l of layer l for the forward pass and for the ordinary weight gradient.
static Operand input_operand(const SdfCfg& c, const SdfBlob& b, int l) {
  if (l == 0) return make_operand(b.E, c.ldE, c.ldE, c.d_e);
  if (l == c.skip) return make_operand(b.U, c.ldH, c.ly.in_ld[l], c.ly.in_dim[l]);
  return make_operand(b.Z[l - 1], c.ldH, c.ly.in_ld[l], c.ly.in_dim[l], PRO_SOFTPLUS);
}

static int launch_embed(const SdfCfg& c, const float* x, long long N, const SdfBlob& b, cudaStream_t st) {
  int ucol = 0;
  if (c.skip >= 0) ucol = c.ly.in_dim[c.skip] - c.d_e;
  return launch_embed_rows(x, c.d_in, N, c.d_in, c.multires, c.scale, b.E, c.ldE, b.U, c.ldH, ucol, kInvSqrt2, c.ldH, st);
}

}  // namespace vdn

using namespace vdn;

extern "C" int vdn_sdf_layer_dims(const int* cfg, int* in_dims, int* out_dims) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, 1.0f, &c)) return -1;
  for (int l = 0; l < c.L; ++l) {
    in_dims[l] = c.ly.in_dim[l];
    out_dims[l] = c.ly.out_dim[l];
  }
  return c.L;
}

extern "C" long long vdn_sdf_blob_floats(const int* cfg, long long N, int save) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, 1.0f, &c)) return -1;
  return sdf_chain_ok(c) ? sdf_chain_blob_floats(c.L, N) : sdf_blob_floats(c, N, save);   // layout of the path the calls take
}

extern "C" int vdn_sdf_layer_orot(const int* cfg, int* orot) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, 1.0f, &c)) return -1;
  for (int l = 0; l < c.L; ++l) orot[l] = 0;
  orot[c.L - 1] = sdf_orot_last(c);
  return c.L;
}

extern "C" long long vdn_sdf_blobg_floats(const int* cfg, long long N) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, 1.0f, &c)) return -1;
  return sdf_chain_ok(c) ? sdf_chain_blobg_floats(c.L, N) : sdf_blobg_floats(c, N);
}

static int sdf_forward_impl(const SdfCfg& c, const float* packed, const float* x, long long N, float* sdf, int lds,
                            float* feat, int ldf, float* blob, int save, float out_mul, cudaStream_t st) {
  if (N <= 0) return 0;
  if (N > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  if (g_mode == 1 && !feat && !save) {   // tensor-core mode, value only: fused chain kernel
    int r = launch_sdf_chain(c.ly, c.d_in, c.multires, c.d_hidden, c.skip, c.scale, packed, x, nullptr, nullptr, nullptr,
                             0, 0, 0, N, sdf, lds, out_mul, st, sdf_orot_last(c));
    if (r >= 0) return r;
  }
  if (sdf_chain_ok(c)) {                 // training / feature forward on the chain engine, 16-bit saved activations
    SdfChainBufs b;
    sdf_chain_carve(c.L, N, blob, nullptr, nullptr, &b);
    return sdf_chain_forward(sdf_shape(c), packed, x, N, sdf, lds, feat, ldf, out_mul, save, b, st);
  }
  SdfBlob b;
  carve_blob(c, N, save, blob, &b);
  int e = launch_embed(c, x, N, b, st);
  if (e) return e;
  int l0 = 0;
  if (g_mode == 1) {
    // tensor-core mode: layers 0..L-2 in the fused chain kernel, which stores the pre-activations the gradient passes
    // need (all of them when saving, else only the last layer's input) and the head of the skip layer's input
    float* zs[VDN_MAX_LAYERS];
    for (int l = 0; l < c.L - 1; ++l) zs[l] = (save || l == c.L - 2) ? b.Z[l] : nullptr;
    int r = launch_sdf_chain(c.ly, c.d_in, c.multires, c.d_hidden, c.skip, c.scale, packed, x, nullptr, nullptr, nullptr,
                             0, 0, 0, N, nullptr, 0, 1.0f, st, sdf_orot_last(c), zs, save ? b.U : nullptr, c.ldH);
    if (r > 0) return r;
    if (r == 0) l0 = c.L - 1;
  }
  for (int l = l0; l < c.L; ++l) {
    Operand A = input_operand(c, b, l);
    const float* bias = packed + c.ly.off_b[l];
    Epilogue E;
    int ngemm = c.ly.out_dim[l];
    if (l == c.L - 1) {
      E = make_epilogue(EPI_SPLIT, bias, feat, ldf);
      E.c2 = sdf; E.ldc2 = lds; E.split = 1; E.scale = out_mul / c.scale;
      if (!feat) ngemm = 1;
    } else if (l + 1 == c.skip) {
      E = make_epilogue(EPI_SDF_SKIP, bias, b.Z[l], c.ldH);
      E.c2 = b.U; E.ldc2 = c.ldH; E.scale = kInvSqrt2;
    } else {
      E = make_epilogue(EPI_STORE, bias, b.Z[l], c.ldH);
    }
    e = launch_gemm_nt((int)N, ngemm, c.ly.in_ld[l], A, wref(c.ly, packed, l), E, st);
    if (e) return e;
  }
  return 0;
}

extern "C" int vdn_sdf_forward(const int* cfg, float scale, const float* packed, const float* x, long long N,
                               float* sdf, int lds, float* feat, int ldf, float* blob, int save, void* stream) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, scale, &c)) return (int)cudaErrorInvalidValue;
  return sdf_forward_impl(c, packed, x, N, sdf, lds, feat, ldf, blob, save, 1.0f, (cudaStream_t)stream);
}

namespace vdn {
// Lattice points of an x-slab [i0, i1) of the extract_fields grid (reference renderer.py:10-30): meshgrid with
// ij indexing of the host-generated linspace coordinates, flattened x-major.
__global__ void grid_points_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                   const float* __restrict__ zs, int i0, int ny, int nz, long long count,
                                   float* __restrict__ pts) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  int k = (int)(idx % nz);
  long long t = idx / nz;
  int j = (int)(t % ny);
  int i = i0 + (int)(t / ny);
  pts[idx * 3] = xs[i];
  pts[idx * 3 + 1] = ys[j];
  pts[idx * 3 + 2] = zs[k];
}
}  // namespace vdn

// u[i0:i1, :, :] = out_mul * sdf(grid points); `pts` is workspace for (i1-i0)*ny*nz*3 floats and `blob` a
// non-saving forward blob for that many points.
extern "C" int vdn_grid_sdf(const int* cfg, float scale, const float* packed, const float* xs, const float* ys,
                            const float* zs, int ny, int nz, int i0, int i1, float out_mul, float* u_slab,
                            float* pts, float* blob, void* stream) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, scale, &c)) return (int)cudaErrorInvalidValue;
  if (c.d_in != 3 || i1 <= i0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  long long count = (long long)(i1 - i0) * ny * nz;
  if (g_mode == 1) {   // tensor-core mode: lattice points are generated inside the fused chain kernel
    int r = launch_sdf_chain(c.ly, c.d_in, c.multires, c.d_hidden, c.skip, c.scale, packed, nullptr, xs, ys, zs, ny, nz,
                             i0, count, u_slab, 1, out_mul, st, sdf_orot_last(c));
    if (r >= 0) return r;
  }
  VDN_LAUNCH(grid_points_kernel, (unsigned)((count + 255) / 256), 256, 0, st, xs, ys, zs, i0, ny, nz, count, pts);
  int e = (int)(cudaError_t)::vdn::take_launch_error();
  if (e) return e;
  return sdf_forward_impl(c, packed, pts, count, u_slab, 1, nullptr, 0, blob, 0, out_mul, st);
}

extern "C" int vdn_sdf_normals(const int* cfg, float scale, const float* packed, const float* x, long long N,
                               const float* blob, float* blobg, float* normals, void* stream) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, scale, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (sdf_chain_ok(c)) {
    SdfChainBufs cb;
    sdf_chain_carve(c.L, N, const_cast<float*>(blob), blobg, nullptr, &cb);
    return sdf_chain_normals(sdf_shape(c), packed, x, N, cb, normals, st);
  }
  SdfBlob b;
  carve_blob(c, N, 1, const_cast<float*>(blob), &b);
  SdfBlobG g;
  carve_blobg(c, N, blobg, &g);
  for (int l = c.L - 2; l >= 0; --l) {
    GinRef gi = gin_of(c, packed, g, l);
    Operand A = make_operand(gi.p, gi.ld, c.ly.out_ld[l], c.ly.out_dim[l], PRO_DSIG, b.Z[l], c.ldH, gi.scale);
    Epilogue E;
    if (l > 0) {
      E = make_epilogue(EPI_STORE, nullptr, g.G[l], c.ldH);
    } else if (c.skip >= 0) {
      E = make_epilogue(EPI_ADD_SCALED, nullptr, g.DE, c.ldE);
      GinRef tail = gin_of(c, packed, g, c.skip - 1);  // G_skip (raw)
      E.aux = tail.p; E.ldaux = tail.ld; E.split = c.ly.in_dim[c.skip] - c.d_e; E.scale = kInvSqrt2;
    } else {
      E = make_epilogue(EPI_STORE, nullptr, g.DE, c.ldE);
    }
    int e = launch_gemm_nt((int)N, c.ly.in_dim[l], c.ly.out_ld[l], A, wtref(c.ly, packed, l), E, st);
    if (e) return e;
  }
  long long tot = N * c.d_in;
  VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, c.d_in, N, c.d_in, c.multires, c.scale, g.DE,
                                                                 c.ldE, nullptr, 0, 0.0f, 1.0f, normals, c.d_in, 0);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" long long vdn_sdf_bwd_ws_floats(const int* cfg, long long N) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, 1.0f, &c)) return -1;
  long long S = wgrad_max_splits(N);
  long long maxw = 0, maxo = 0;
  for (int l = 0; l < c.L; ++l) {
    long long w = (long long)c.ly.out_dim[l] * ((c.ly.in_dim[l] + 3) & ~3);
    if (w > maxw) maxw = w;
    if (c.ly.out_ld[l] > maxo) maxo = c.ly.out_ld[l];
  }
  if (sdf_chain_ok(c)) return sdf_chain_ws_floats(c.L, N);
  return 3 * N * c.ldH + (long long)(c.L - 1) * N * c.ldH + N * c.ly.out_ld[c.L - 1] + 3 * N * c.ldE + S * maxw + 256 * maxo +
         64;
}

extern "C" int vdn_sdf_backward(const int* cfg, float scale, const float* packed, const float* x, long long N,
                                const float* blob, const float* blobg, const float* d_sdf, int lds,
                                const float* d_feat, int ldf, const float* d_normals, float* dpacked, float* d_x,
                                float* ws,
                                void* stream) {
  SdfCfg c;
  if (parse_sdf_cfg(cfg, scale, &c)) return (int)cudaErrorInvalidValue;
  if (N <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const MlpLayout& ly = c.ly;
  const int L = c.L;
  SdfBlob b;
  carve_blob(c, N, 1, const_cast<float*>(blob), &b);
  SdfBlobG g;
  if (blobg) carve_blobg(c, N, const_cast<float*>(blobg), &g);
  const bool have_n = (d_normals != nullptr);
  if (have_n && !blobg) return (int)cudaErrorInvalidValue;
  if (sdf_chain_ok(c)) {
    SdfChainBufs cb;
    sdf_chain_carve(c.L, N, const_cast<float*>(blob), const_cast<float*>(blobg), ws, &cb);
    return sdf_chain_backward(sdf_shape(c), packed, x, N, cb, d_sdf, lds, d_feat, ldf, d_normals, dpacked, d_x, st);
  }

  // workspace carve
  float* Q[2] = {ws, ws + N * c.ldH};
  float* QS = ws + 2 * N * c.ldH;
  float* ZG[VDN_MAX_LAYERS];
  float* p = ws + 3 * N * c.ldH;
  for (int l = 0; l < L - 1; ++l) { ZG[l] = p; p += N * c.ldH; }
  float* ZL = p; p += N * ly.out_ld[L - 1];
  float* DEB = p; p += N * c.ldE;
  float* ES = p; p += N * c.ldE;
  float* EB = p; p += N * c.ldE;
  float* partials = p;
  int e;
  const int M = (int)N;

  // ---- phase 1: backward of the normals pass (walks l = 0 .. L-2) ---------------------------------
  if (have_n) {
    int qcol = (c.skip >= 0) ? ly.in_dim[c.skip] - c.d_e : 0;
    VDN_LAUNCH(embed_jvp_kernel, (unsigned)((N + 127) / 128), 128, 0, st, x, c.d_in, N, c.d_in, c.multires, c.scale,
                                                                 d_normals, c.d_in, DEB, c.ldE,
                                                                 c.skip >= 0 ? QS : nullptr, c.ldH, qcol, kInvSqrt2,
                                                                 c.ldH);
    e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
    for (int l = 0; l <= L - 2; ++l) {
      Operand qbar = (l == 0) ? make_operand(DEB, c.ldE, c.ldE, c.d_e)
                              : make_operand(l == c.skip ? QS : Q[l & 1], c.ldH, ly.in_ld[l], ly.in_dim[l]);
      if (l > 0 && l != c.skip) qbar.rounded = 1;   // written by the previous layer's EPI_GRAD_DUAL with round_c
      GinRef gi = gin_of(c, packed, g, l);
      // dbar_l = qbar_l * W_l^T ; epilogue splits it into qbar_{l+1} and the injected pre-activation cotangent
      Epilogue E = make_epilogue(EPI_GRAD_DUAL, nullptr, (l + 1 == c.skip) ? QS : Q[(l + 1) & 1], c.ldH);
      E.round_c = 1;
      E.scale = (l + 1 == c.skip) ? kInvSqrt2 : 1.0f;
      E.c2 = ZG[l]; E.ldc2 = c.ldH;
      E.aux = b.Z[l]; E.ldaux = c.ldH;
      E.aux2 = gi.p; E.ldaux2 = gi.ld; E.scale2 = gi.scale;
      e = launch_gemm_nt(M, ly.out_dim[l], ly.in_ld[l], qbar, wref(ly, packed, l), E, st);
      if (e) return e;
      // Wbar_l += delta_l^T qbar_l with delta_l = softplus'(z_l) * Gin_l recomputed on the fly
      Operand delta = make_operand(gi.p, gi.ld, ly.out_ld[l], ly.out_dim[l], PRO_DSIG, b.Z[l], c.ldH, gi.scale);
      e = launch_wgrad_any(M, ly.out_dim[l], ly.in_dim[l], delta, qbar, partials, dpacked + ly.off_w[l], ly.in_ld[l], 1,
                           nullptr, st);
      if (e) return e;
    }
    // a_{L-1} is the first row of the last weight: its cotangent is the column sum of Gin-bar_{L-2}
    Operand gb = make_operand(Q[(L - 1) & 1], c.ldH, ly.out_ld[L - 2], ly.out_dim[L - 2]);
    e = launch_colsum(M, ly.out_dim[L - 2], gb, partials, dpacked + ly.off_w[L - 1], 1, st);
    if (e) return e;
  }

  // ---- phase 2: ordinary backward with the injected cotangents ------------------------------------
  {
    // cotangent of the last layer's stacked [sdf | feature] output
    e = launch_gather2_rows(d_sdf, lds, 1, 1.0f / c.scale, d_feat, ldf, c.d_out - 1, 1.0f, N, ZL, ly.out_ld[L - 1], st);
    if (e) return e;
  }
  for (int l = L - 1; l >= 0; --l) {
    Operand zbar = (l == L - 1) ? make_operand(ZL, ly.out_ld[l], ly.out_ld[l], ly.out_dim[l])
                                : make_operand(ZG[l], c.ldH, ly.out_ld[l], ly.out_dim[l]);
    if (l < L - 1) zbar.rounded = 1;   // written by EPI_BWD_INJECT with round_c
    Operand u = input_operand(c, b, l);
    e = launch_wgrad_any(M, ly.out_dim[l], ly.in_dim[l], zbar, u, partials, dpacked + ly.off_w[l], ly.in_ld[l], 1,
                         dpacked + ly.off_b[l], st);
    if (e) return e;
    if (l > 0) {
      // ubar = zbar_l W_l restricted to the hidden part; epilogue -> zbar_{l-1}
      Epilogue E = make_epilogue(EPI_BWD_INJECT, nullptr, ZG[l - 1], c.ldH);
      E.round_c = 1;
      E.aux = b.Z[l - 1]; E.ldaux = c.ldH;
      E.aux2 = have_n ? ZG[l - 1] : nullptr; E.ldaux2 = c.ldH;
      E.scale = (l == c.skip) ? kInvSqrt2 : 1.0f;
      e = launch_gemm_nt(M, ly.out_dim[l - 1], ly.out_ld[l], zbar, wtref(ly, packed, l), E, st);
      if (e) return e;
      if (l == c.skip && d_x) {  // embedding tail of the skip concat
        int off = ly.in_dim[l] - c.d_e;
        Epilogue E2 = make_epilogue(EPI_STORE, nullptr, ES, c.ldE);
        e = launch_gemm_nt(M, c.d_e, ly.out_ld[l], zbar, wtref(ly, packed, l, off), E2, st);
        if (e) return e;
      }
    } else if (d_x) {
      Epilogue E2 = make_epilogue(EPI_STORE, nullptr, EB, c.ldE);
      e = launch_gemm_nt(M, c.d_e, ly.out_ld[0], zbar, wtref(ly, packed, 0), E2, st);
      if (e) return e;
    }
  }
  if (d_x) {
    long long tot = N * c.d_in;
    VDN_LAUNCH(embed_vjp_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, c.d_in, N, c.d_in, c.multires, c.scale, EB,
                                                                   c.ldE, c.skip >= 0 ? ES : nullptr, c.ldE,
                                                                   kInvSqrt2, c.scale, d_x, c.d_in, 0);
    if (have_n) {
      VDN_LAUNCH(embed_second_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, x, c.d_in, N, c.d_in, c.multires, c.scale,
                                                                        d_normals, c.d_in, g.DE, c.ldE, nullptr, 0,
                                                                        c.scale, d_x, c.d_in);
    }
    e = (int)(cudaError_t)::vdn::take_launch_error();
    if (e) return e;
  }
  return 0;
}

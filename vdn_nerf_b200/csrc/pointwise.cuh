// Small pointwise kernels around the MLP chains: positional encoding (forward, J e, J^T), input assembly
// for the rendering / background networks, and cotangent packing.  All fp32, no fast-math.
#pragma once
#include "common.cuh"

namespace vdn {

// e(y) = [y | sin(2^0 y) | cos(2^0 y) | ... ]  with y = x * scale  (reference embedder.py:15-36).
// Writes row m of `e` (ld lde; columns >= d_e zeroed up to lde) and, when u != nullptr, the same values
// times uscale into u[m, ucol ... ucol+d_e) (zero up to u_pad_to): the skip-connection tail of the SDF net
// (fields.py:82-83) or of the NeRF field.  One thread per (row, frequency, coordinate): a single sincosf feeds the
// sin and the cos column of both destinations; further threads of the row write the identity and padding columns.
static __global__ void embed_rows_kernel(const float* __restrict__ x, int ldx, long long N, int d, int L, float scale,
                                  float* __restrict__ e, int lde, float* __restrict__ u, int ldu, int ucol,
                                  float uscale, int u_pad_to) {
  const int d_e = d * (1 + 2 * L);
  const int pe = e ? lde - d_e : 0;                   // padding columns of e
  const int pu = u ? u_pad_to - ucol - d_e : 0;       // padding columns of u
  const int T = d + d * L + pe + pu;                  // work items per row
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * T) return;
  long long m;
  int t;
  if (N * T < 0x7fffffffLL) {
    const unsigned mi = (unsigned)idx / (unsigned)T;
    m = mi;
    t = (int)((unsigned)idx - mi * (unsigned)T);
  } else {
    m = idx / T;
    t = (int)(idx - m * T);
  }
  float* er = e ? e + m * lde : nullptr;
  float* ur = u ? u + m * ldu + ucol : nullptr;
  if (t < d) {
    const float y = x[m * ldx + t] * scale;
    if (er) er[t] = y;
    if (ur) ur[t] = y * uscale;
  } else if (t < d + d * L) {
    const int q = t - d, k = q / d, j = q - k * d;
    const float y = x[m * ldx + j] * scale;
    float sn, cs;
    sincosf(y * (float)(1 << k), &sn, &cs);
    const int c0 = d + 2 * k * d + j, c1 = c0 + d;
    if (er) { er[c0] = sn; er[c1] = cs; }
    if (ur) { ur[c0] = sn * uscale; ur[c1] = cs * uscale; }
  } else {
    const int p = t - d - d * L;
    if (p < pe) er[d_e + p] = 0.0f;
    else ur[d_e + (p - pe)] = 0.0f;
  }
}

inline int launch_embed_rows(const float* x, int ldx, long long N, int d, int L, float scale, float* e, int lde, float* u,
                             int ldu, int ucol, float uscale, int u_pad_to, cudaStream_t st) {
  const int d_e = d * (1 + 2 * L);
  const long long T = d + d * L + (e ? lde - d_e : 0) + (u ? u_pad_to - ucol - d_e : 0);
  const long long total = N * T;
  if (total <= 0) return 0;
  VDN_LAUNCH(embed_rows_kernel, (unsigned)((total + 255) / 256), 256, 0, st, x, ldx, N, d, L, scale, e, lde, u, ldu, ucol,
             uscale, u_pad_to);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// out[m, j] (+)= oscale * sum_c de[m,c] * d e_c / d y_j  (= J_e^T de), j < d.   de rows have ld ldde.
// Optional second cotangent de2 (added with weight w2) lets the caller fold the skip path in.
static __global__ void embed_vjp_kernel(const float* __restrict__ x, int ldx, long long N, int d, int L, float scale,
                                 const float* __restrict__ de, int ldde, const float* __restrict__ de2, int ldde2,
                                 float w2, float oscale, float* __restrict__ out, int ldo, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * d) return;
  long long m = idx / d;
  int j = (int)(idx - m * d);
  float y = x[m * ldx + j] * scale;
  const float* r = de + m * ldde;
  const float* r2 = de2 ? de2 + m * ldde2 : nullptr;
  auto val = [&](int c) { return r[c] + (r2 ? w2 * r2[c] : 0.0f); };
  float acc = val(j);
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s, c;
    sincosf(y * f, &s, &c);
    int cs = d + (2 * k) * d + j, cc = cs + d;
    acc += f * (c * val(cs) - s * val(cc));
    f *= 2.0f;
  }
  acc *= oscale;
  float* o = out + m * ldo + j;
  *o = accumulate ? (*o + acc) : acc;
}

// Forward-mode product with the embedding Jacobian: t[m, c] = (J_e n)[c] = (d e_c / d y_j) * n[m, j].
// Writes t (ld ldt, zero padded to ldt) and optionally the scaled copy into q[m, qcol + c] (skip tail).
static __global__ void embed_jvp_kernel(const float* __restrict__ x, int ldx, long long N, int d, int L, float scale,
                                 const float* __restrict__ nbar, int ldn, float* __restrict__ t, int ldt,
                                 float* __restrict__ q, int ldq, int qcol, float qscale, int q_pad_to) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= N) return;
  const int d_e = d * (1 + 2 * L);
  float* tr = t + m * ldt;
  float* qr = q ? q + m * ldq + qcol : nullptr;
  for (int j = 0; j < d; ++j) {
    float y = x[m * ldx + j] * scale;
    float nb = nbar[m * ldn + j];
    tr[j] = nb;
    if (qr) qr[j] = nb * qscale;
    float f = 1.0f;
    for (int k = 0; k < L; ++k) {
      float s, c;
      sincosf(y * f, &s, &c);
      int cs = d + (2 * k) * d + j, cc = cs + d;
      float vs = f * c * nb, vc = -f * s * nb;
      tr[cs] = vs; tr[cc] = vc;
      if (qr) { qr[cs] = vs * qscale; qr[cc] = vc * qscale; }
      f *= 2.0f;
    }
  }
  for (int j = d_e; j < ldt; ++j) tr[j] = 0.0f;
  if (qr) for (int j = qcol + d_e; j < q_pad_to; ++j) q[m * ldq + j] = 0.0f;
}

// Second-order term of the normal w.r.t. the point: xbar[m,j] += oscale * nbar[m,j] *
//   sum_k f_k^2 ( -sin(f_k y_j) de_sin[k,j] - cos(f_k y_j) de_cos[k,j] )       
static __global__ void embed_second_kernel(const float* __restrict__ x, int ldx, long long N, int d, int L, float scale,
                                    const float* __restrict__ nbar, int ldn, const float* __restrict__ de, int ldde,
                                    const float* __restrict__ de2, int ldde2, float oscale, float* __restrict__ out,
                                    int ldo) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * d) return;
  long long m = idx / d;
  int j = (int)(idx - m * d);
  float y = x[m * ldx + j] * scale;
  const float* r = de + m * ldde;
  const float* r2 = de2 ? de2 + m * ldde2 : nullptr;       // optional second part of d sdf / d e (skip connection)
  float acc = 0.0f, f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s, c;
    sincosf(y * f, &s, &c);
    int cs = d + (2 * k) * d + j, cc = cs + d;
    const float vs = r[cs] + (r2 ? r2[cs] : 0.0f), vc = r[cc] + (r2 ? r2[cc] : 0.0f);
    acc += f * f * (-s * vs - c * vc);
    f *= 2.0f;
  }
  out[m * ldo + j] += oscale * nbar[m * ldn + j] * acc;
}

// dst[m, 0..w) = src[m, 0..w) * scale, zero padded to ldd; generic strided row copy.
static __global__ void copy_pad_rows_kernel(const float* __restrict__ src, int lds, int w, long long N,
                                     float* __restrict__ dst, int ldd, int dcol, int pad_to, float scale) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int span = pad_to - dcol;
  if (idx >= N * span) return;
  long long m = idx / span;
  int c = (int)(idx - m * span);
  dst[m * ldd + dcol + c] = (c < w && src) ? src[m * lds + c] * scale : 0.0f;
}

// dst[m, :] = [ a[m, 0..wa) * sa | b[m, 0..wb) * sb | 0 ... ] up to ldd columns (null source = zeros): the cotangent of
// a stacked pair of heads assembled in ONE pass (it used to take a zero fill plus one strided copy per head).
static __global__ void gather2_rows_kernel(const float* __restrict__ a, int lda, int wa, float sa, const float* __restrict__ b,
                                           int ldb, int wb, float sb, long long N, float* __restrict__ dst, int ldd) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * ldd) return;
  long long m;
  int c;
  if (N * ldd < 0x7fffffffLL) {
    const unsigned mi = (unsigned)idx / (unsigned)ldd;
    m = mi;
    c = (int)((unsigned)idx - mi * (unsigned)ldd);
  } else {
    m = idx / ldd;
    c = (int)(idx - m * ldd);
  }
  float v = 0.0f;
  if (c < wa) { if (a) v = a[m * lda + c] * sa; }
  else if (c < wa + wb) { if (b) v = b[m * ldb + (c - wa)] * sb; }
  dst[idx] = v;
}
inline int launch_gather2_rows(const float* a, int lda, int wa, float sa, const float* b, int ldb, int wb, float sb,
                               long long N, float* dst, int ldd, cudaStream_t st) {
  const long long tot = N * ldd;
  if (tot <= 0) return 0;
  VDN_LAUNCH(gather2_rows_kernel, (unsigned)((tot + 255) / 256), 256, 0, st, a, lda, wa, sa, b, ldb, wb, sb, N, dst, ldd);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

// dst[m, 0] = a ? a[m] : 0 and dst[m, wreal..ldd) = 0: the columns of a stacked-head cotangent that the following GEMM
// epilogue (which fills columns 1..wreal-1) does not write.
static __global__ void head_edges_kernel(const float* __restrict__ a, long long N, float* __restrict__ dst, int ldd,
                                         int wreal) {
  const int ne = 1 + (ldd - wreal);
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * ne) return;
  const long long m = idx / ne;
  const int j = (int)(idx - m * ne);
  if (j == 0) dst[m * ldd] = a ? a[m] : 0.0f;
  else dst[m * ldd + wreal + j - 1] = 0.0f;
}

// Rendering-network input row (reference fields.py:148-158), stored ROTATED as [feats(F) | extras] like the packed
// first layer (rot = in0 - F), so that the fused chains can treat the 256-wide feature as the main operand:
//   idr extras:   [points(3) | PE_L(view)(3+6L) | normals(3)]
//   no_view_dir:  [points | normals]         no_normal: [points | PE(view)]
// One warp per row: lanes 0..2 write the point / view-embedding / normal columns of their coordinate,
// all lanes copy the feature columns with coalesced accesses.
static __global__ void rendernet_input_kernel(const float* __restrict__ pts, const float* __restrict__ nrm,
                                       const float* __restrict__ view, const float* __restrict__ feat, int ldf,
                                       int F, int L, int mode, long long N, float* __restrict__ cin, int ldc) {
  long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= N) return;
  float* r = cin + m * ldc;
  const int d = 3;
  const int nview = (mode != 1) ? d * (1 + 2 * L) : 0;
  const int nnrm = (mode != 2) ? 3 : 0;
  float* x = r + F;                     // extras follow the feature columns
  if (lane < 3) {
    const int j = lane;
    x[j] = pts[m * 3 + j];
    if (mode != 1) {
      float v = view[m * 3 + j];
      x[3 + j] = v;
      float f = 1.0f;
      for (int k = 0; k < L; ++k) {
        float s, co;
        sincosf(v * f, &s, &co);
        x[3 + d + 2 * k * d + j] = s;
        x[3 + d + 2 * k * d + d + j] = co;
        f *= 2.0f;
      }
    }
    if (mode != 2) x[3 + nview + j] = nrm[m * 3 + j];
  }
  const int c0 = 3 + nview + nnrm;
  for (int j = lane; j < F; j += 32) r[j] = feat[m * ldf + j];
  for (int c = c0 + F + lane; c < ldc; c += 32) r[c] = 0.0f;
}

// dst[m, c] = d_out[m,c] * act'(out[m,c]) for c < w, zero for w <= c < ldd.  kind 0: sigmoid, 1: relu.
static __global__ void act_backward_pad_kernel(const float* __restrict__ d_out, const float* __restrict__ out, int w,
                                        long long N, float* __restrict__ dst, int ldd, int kind) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * ldd) return;
  long long m = idx / ldd;
  int c = (int)(idx - m * ldd);
  float v = 0.0f;
  if (c < w) {
    float d = d_out[m * w + c], o = out[m * w + c];
    v = (kind == 0) ? d * ((1.0f - o) * o) : (o > 0.0f ? d : 0.0f);
  }
  dst[idx] = v;
}

}  // namespace vdn

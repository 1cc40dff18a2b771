// Warp-per-ray kernels of the volume-rendering path (reference dpt_models/renderer.py):
//   vdn_ray_points      p = o + d * z                                     renderer.py:150, 196, 369
//   vdn_upsample_step   up_sample + sample_pdf(det) + the sort of cat_z_vals, fused      :44-74, 147-207
//   vdn_fine_prep       section lengths / mid points / fine sample points                 :228-237
//   vdn_bg_prep         merge with the outside samples, inverted-sphere reparametrisation :100-120, 388-391
//   vdn_composite_fwd   sigmoid-CDF alpha, background blend, transmittance scan, compositing, Eikonal sums :262-315
//   vdn_composite_bwd   closed-form backward of the above
//
// Per-ray arrays (<= 256 samples) are staged in shared memory; global accesses are lane-contiguous.
// This file is compiled with -fmad=false: the up-sampling arithmetic reproduces the reference's separately
// rounded fp32 elementwise ops, and its scans use fp64 accumulators like ATen's CPU cumsum/cumprod, so the
// sample indices are bit-exact against the CPU reference except at exact ties of u against a CDF knot.
#include "common.cuh"
#include "../../include/vdn_b200.h"

namespace vdn {

constexpr int RAY_MAXN = 256;        // max samples per ray handled in shared memory
constexpr int RAY_WARPS = 4;         // rays per CTA

__device__ __forceinline__ float sigmoid_ref(float x) {
  // 1 / (1 + exp(-x)) with a correctly rounded fp32 exp
  float e = (float)exp((double)(-x));
  return 1.0f / (1.0f + e);
}
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

__global__ void ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                  const float* __restrict__ z, long long B, int n, float* __restrict__ pts) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  long long b = idx / n;
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[idx * 3 + c] = o[b * 3 + c] + d[b * 3 + c] * z[idx];
}

// ------------------------------------------------------------------------------------------------
// One hierarchical up-sampling iteration.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RAY_WARPS * 32)
upsample_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z_in, int n,
                const float* __restrict__ sdf_prev, int n_prev, const float* __restrict__ sdf_new, int n_new_prev,
                const unsigned char* __restrict__ perm_prev, float inv_s, int n_imp, long long B,
                float* __restrict__ z_out, float* __restrict__ sdf_out, unsigned char* __restrict__ perm_out,
                float* __restrict__ new_z, float* __restrict__ new_pts, long long* __restrict__ inds_out) {
  __shared__ float sz[RAY_WARPS][RAY_MAXN], ss[RAY_WARPS][RAY_MAXN], sw[RAY_WARPS][RAY_MAXN],
      sc[RAY_WARPS][RAY_MAXN], sn[RAY_WARPS][32];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * RAY_WARPS + wid;
  if (b >= B) return;
  float* z = sz[wid];
  float* s = ss[wid];
  float* w = sw[wid];
  float* cdf = sc[wid];
  float* zn = sn[wid];
  const float ox = o[b * 3], oy = o[b * 3 + 1], oz = o[b * 3 + 2];
  const float dx = d[b * 3], dy = d[b * 3 + 1], dz = d[b * 3 + 2];

  // (a) load z, gather the merged sdf of the previous iteration, radii into cdf[] (scratch)
  for (int k = lane; k < n; k += 32) {
    float zk = z_in[b * n + k];
    z[k] = zk;
    float sk;
    if (perm_prev) {
      int src = perm_prev[b * n + k];
      sk = (src < n_prev) ? sdf_prev[b * n_prev + src] : sdf_new[b * n_new_prev + (src - n_prev)];
    } else {
      sk = sdf_prev[b * n + k];
    }
    s[k] = sk;
    if (sdf_out) sdf_out[b * n + k] = sk;
    cdf[k] = norm3(ox + dx * zk, oy + dy * zk, oz + dz * zk);
  }
  __syncwarp();
  // (b) per-interval alpha (renderer.py:153-186)
  for (int k = lane; k < n - 1; k += 32) {
    const bool inside = (cdf[k] < 1.0f) | (cdf[k + 1] < 1.0f);
    const float dist = z[k + 1] - z[k];
    float cosv = (s[k + 1] - s[k]) / (dist + 1e-5f);
    float prevc = 0.0f;
    if (k > 0) prevc = (s[k] - s[k - 1]) / ((z[k] - z[k - 1]) + 1e-5f);
    float c = fminf(prevc, cosv);
    c = fminf(fmaxf(c, -1e3f), 0.0f) * (inside ? 1.0f : 0.0f);
    const float mid = (s[k] + s[k + 1]) * 0.5f;
    const float pe = mid - c * dist * 0.5f;
    const float ne = mid + c * dist * 0.5f;
    const float P = sigmoid_ref(pe * inv_s);
    const float Nn = sigmoid_ref(ne * inv_s);
    w[k] = (P - Nn + 1e-5f) / (P + 1e-5f);
  }
  __syncwarp();
  // (c) sequential scans with fp64 accumulators (ATen CPU cumprod / cumsum semantics)
  if (lane == 0) {
    double T = 1.0, sum = 0.0;
    for (int k = 0; k < n - 1; ++k) {
      const float a = w[k];
      float wk = a * (float)T;
      T *= (double)(1.0f - a + 1e-7f);
      wk = wk + 1e-5f;
      w[k] = wk;
      sum += (double)wk;
    }
    const float sumf = (float)sum;
    double acc = 0.0;
    cdf[0] = 0.0f;
    for (int k = 0; k < n - 1; ++k) {
      acc += (double)(w[k] / sumf);
      cdf[k + 1] = (float)acc;
    }
  }
  __syncwarp();
  // (d) invert the CDF at the deterministic u (renderer.py:52-72)
  if (lane < n_imp) {
    const float start = 0.5f / n_imp, end = 1.0f - 0.5f / n_imp;
    const float step = (n_imp > 1) ? (end - start) / (float)(n_imp - 1) : 0.0f;
    const float u = (lane < n_imp / 2) ? start + step * (float)lane : end - step * (float)(n_imp - 1 - lane);
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(ind, n - 1);
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-5f) denom = 1.0f;
    const float t = (u - cdf[below]) / denom;
    const float znew = z[below] + t * (z[above] - z[below]);
    zn[lane] = znew;
    if (inds_out) inds_out[b * n_imp + lane] = ind;
    new_z[b * n_imp + lane] = znew;
    if (new_pts) {
      float* p = new_pts + (b * n_imp + lane) * 3;
      p[0] = ox + dx * znew; p[1] = oy + dy * znew; p[2] = oz + dz * znew;
    }
  }
  __syncwarp();
  // (e) stable merge of the sorted old samples with the new ones (cat + sort of renderer.py:197-198)
  const int nt = n + n_imp;
  for (int k = lane; k < n; k += 32) {
    const float zk = z[k];
    int cnt = 0;
    for (int j = 0; j < n_imp; ++j) cnt += (zn[j] < zk) ? 1 : 0;
    z_out[b * nt + k + cnt] = zk;
    perm_out[b * nt + k + cnt] = (unsigned char)k;
  }
  if (lane < n_imp) {
    const float v = zn[lane];
    int lo = 0, hi = n;  // number of old samples <= v
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (z[mid] <= v) lo = mid + 1; else hi = mid;
    }
    int cnt = lo;
    for (int j = 0; j < n_imp; ++j) cnt += (zn[j] < v || (zn[j] == v && j < lane)) ? 1 : 0;
    z_out[b * nt + cnt] = v;
    perm_out[b * nt + cnt] = (unsigned char)(n + lane);
  }
}

// Stable merge of sorted za[B,n] with zb[B,m] (cat + sort of renderer.py:197-198): z_out[B,n+m] and the source
// index of every output sample (< n: from za, >= n: from zb).
__global__ void __launch_bounds__(RAY_WARPS * 32)
merge_sorted_kernel(const float* __restrict__ za, int n, const float* __restrict__ zb, int m, long long B,
                    float* __restrict__ z_out, unsigned char* __restrict__ perm) {
  __shared__ float sa[RAY_WARPS][RAY_MAXN], sb[RAY_WARPS][RAY_MAXN];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * RAY_WARPS + wid;
  if (b >= B) return;
  float* a = sa[wid];
  float* c = sb[wid];
  for (int k = lane; k < n; k += 32) a[k] = za[b * n + k];
  for (int j = lane; j < m; j += 32) c[j] = zb[b * m + j];
  __syncwarp();
  const int nt = n + m;
  for (int k = lane; k < n; k += 32) {
    const float v = a[k];
    int cnt = 0;
    for (int j = 0; j < m; ++j) cnt += (c[j] < v) ? 1 : 0;
    z_out[b * nt + k + cnt] = v;
    perm[b * nt + k + cnt] = (unsigned char)k;
  }
  for (int j = lane; j < m; j += 32) {
    const float v = c[j];
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    int cnt = lo;
    for (int j2 = 0; j2 < m; ++j2) cnt += (c[j2] < v || (c[j2] == v && j2 < j)) ? 1 : 0;
    z_out[b * nt + cnt] = v;
    perm[b * nt + cnt] = (unsigned char)(n + j);
  }
}

// ------------------------------------------------------------------------------------------------
// Sample placement for render_core / render_core_outside
// ------------------------------------------------------------------------------------------------
__global__ void fine_prep_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                 const float* __restrict__ z, float sample_dist, long long B, int S,
                                 float* __restrict__ dists, float* __restrict__ mid_z, float* __restrict__ pts) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  long long b = idx / S;
  int k = (int)(idx - b * S);
  float zk = z[idx];
  float dist = (k + 1 < S) ? z[idx + 1] - zk : sample_dist;
  float mid = zk + dist * 0.5f;
  dists[idx] = dist;
  mid_z[idx] = mid;
#pragma unroll
  for (int c = 0; c < 3; ++c) pts[idx * 3 + c] = o[b * 3 + c] + d[b * 3 + c] * mid;
}

__global__ void __launch_bounds__(RAY_WARPS * 32)
bg_prep_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z_fine, int S,
               const float* __restrict__ z_outside, int NO, float sample_dist, long long B,
               float* __restrict__ dists, float* __restrict__ mid_z, float* __restrict__ pts4) {
  __shared__ float sz[RAY_WARPS][RAY_MAXN], sm[RAY_WARPS][RAY_MAXN];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * RAY_WARPS + wid;
  if (b >= B) return;
  float* zf = sz[wid];   // fine z, later reused
  float* zm = sm[wid];   // merged
  const int nt = S + NO;
  for (int k = lane; k < S; k += 32) zf[k] = z_fine[b * S + k];
  __syncwarp();
  // stable merge: fine samples are sorted; outside samples are ranked individually
  for (int k = lane; k < S; k += 32) {
    const float zk = zf[k];
    int cnt = 0;
    for (int j = 0; j < NO; ++j) cnt += (z_outside[b * NO + j] < zk) ? 1 : 0;
    zm[k + cnt] = zk;
  }
  for (int j = lane; j < NO; j += 32) {
    const float v = z_outside[b * NO + j];
    int lo = 0, hi = S;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (zf[mid] <= v) lo = mid + 1; else hi = mid;
    }
    int cnt = lo;
    for (int j2 = 0; j2 < NO; ++j2) {
      const float v2 = z_outside[b * NO + j2];
      cnt += (v2 < v || (v2 == v && j2 < j)) ? 1 : 0;
    }
    zm[cnt] = v;
  }
  __syncwarp();
  const float ox = o[b * 3], oy = o[b * 3 + 1], oz = o[b * 3 + 2];
  const float dx = d[b * 3], dy = d[b * 3 + 1], dz = d[b * 3 + 2];
  for (int k = lane; k < nt; k += 32) {
    const float zk = zm[k];
    const float dist = (k + 1 < nt) ? zm[k + 1] - zk : sample_dist;
    const float mid = zk + dist * 0.5f;
    dists[b * nt + k] = dist;
    mid_z[b * nt + k] = mid;
    const float px = ox + dx * mid, py = oy + dy * mid, pz = oz + dz * mid;
    const float r = fminf(fmaxf(norm3(px, py, pz), 1.0f), 1e10f);
    float4 q = make_float4(px / r, py / r, pz / r, 1.0f / r);
    *reinterpret_cast<float4*>(pts4 + (b * nt + k) * 4) = q;
  }
}

// ------------------------------------------------------------------------------------------------
// Compositing
// ------------------------------------------------------------------------------------------------
struct CompArgs {
  long long B;
  int S, NB, F;                 // fine samples, background samples (0 = none; else >= S), feature width
  const float *o, *d, *mid_z, *dists;
  const float *sdf, *nrm, *col, *feat;
  const float *sigma_bg, *rgb_bg, *feat_bg, *dists_bg;
  const float* variance;
  const float* bg_rgb;          // [3] or null
  float cos_anneal;
};

struct FineEval {
  float P, Nn, raw, alpha, tc, a1, ic, ep, en, inside, relax, gnorm;
};

__device__ __forceinline__ float inv_s_of(const float* variance) {
  return fminf(fmaxf(expf(variance[0] * 10.0f), 1e-6f), 1e6f);
}

__device__ __forceinline__ FineEval eval_fine(const CompArgs& a, long long b, int k, float inv_s, float dx, float dy,
                                              float dz, float ox, float oy, float oz) {
  FineEval f;
  const long long i = b * a.S + k;
  const float gx = a.nrm[i * 3], gy = a.nrm[i * 3 + 1], gz = a.nrm[i * 3 + 2];
  const float sdf = a.sdf[i], dist = a.dists[i], mid = a.mid_z[i];
  f.tc = dx * gx + dy * gy + dz * gz;
  f.a1 = -f.tc * 0.5f + 0.5f;
  f.ic = -(fmaxf(f.a1, 0.0f) * (1.0f - a.cos_anneal) + fmaxf(-f.tc, 0.0f) * a.cos_anneal);
  f.en = sdf + f.ic * dist * 0.5f;
  f.ep = sdf - f.ic * dist * 0.5f;
  f.P = sigmoidf_(f.ep * inv_s);
  f.Nn = sigmoidf_(f.en * inv_s);
  f.raw = (f.P - f.Nn + 1e-5f) / (f.P + 1e-5f);
  f.alpha = fminf(fmaxf(f.raw, 0.0f), 1.0f);
  const float pn = norm3(ox + dx * mid, oy + dy * mid, oz + dz * mid);
  f.inside = pn < 1.0f ? 1.0f : 0.0f;
  f.relax = pn < 1.2f ? 1.0f : 0.0f;
  f.gnorm = norm3(gx, gy, gz);
  return f;
}

__device__ __forceinline__ float softplus1(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// Background alpha: 1 - exp(-softplus(sigma) * dist) (renderer.py:124), or sigma_bg taken as an already
// computed alpha when no section lengths are given (the stand-alone render_core API of renderer.py:209).
__device__ __forceinline__ float bg_alpha(const CompArgs& a, long long j) {
  if (!a.dists_bg) return a.sigma_bg[j];
  return 1.0f - expf(-softplus1(a.sigma_bg[j]) * a.dists_bg[j]);
}

// blocked exclusive product scan over NW values held in shared memory: T[k] = prod_{j<k} (1 - alpha[j] + 1e-7)
__device__ __forceinline__ void transmittance_scan(const float* alpha, float* T, int NW, int lane) {
  const int per = (NW + 31) / 32;
  const int k0 = lane * per;
  float local = 1.0f;
  for (int i = 0; i < per; ++i) {
    int k = k0 + i;
    if (k < NW) local *= (1.0f - alpha[k] + 1e-7f);
  }
  float incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl *= v;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  float run = excl;
  for (int i = 0; i < per; ++i) {
    int k = k0 + i;
    if (k < NW) {
      T[k] = run;
      run *= (1.0f - alpha[k] + 1e-7f);
    }
  }
}

__global__ void __launch_bounds__(RAY_WARPS * 32)
composite_fwd_kernel(CompArgs a, float* __restrict__ weights, float* __restrict__ cdf_out,
                     float* __restrict__ inside_out, float* __restrict__ color, float* __restrict__ dfeat,
                     float* __restrict__ eik_num, float* __restrict__ eik_den) {
  __shared__ float sa[RAY_WARPS][RAY_MAXN], sT[RAY_WARPS][RAY_MAXN], sin_[RAY_WARPS][RAY_MAXN];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * RAY_WARPS + wid;
  if (b >= a.B) return;
  float* alpha = sa[wid];
  float* T = sT[wid];
  float* ins = sin_[wid];
  const int S = a.S, NB = a.NB, NW = NB > 0 ? NB : S;
  const float inv_s = inv_s_of(a.variance);
  const float ox = a.o[b * 3], oy = a.o[b * 3 + 1], oz = a.o[b * 3 + 2];
  const float dx = a.d[b * 3], dy = a.d[b * 3 + 1], dz = a.d[b * 3 + 2];
  float en = 0.0f, ed = 0.0f;
  for (int k = lane; k < NW; k += 32) {
    float al, in = 0.0f;
    if (k < S) {
      FineEval f = eval_fine(a, b, k, inv_s, dx, dy, dz, ox, oy, oz);
      al = f.alpha;
      in = f.inside;
      cdf_out[b * S + k] = f.P;
      inside_out[b * S + k] = f.inside;
      const float ge = (f.gnorm - 1.0f) * (f.gnorm - 1.0f);
      en += f.relax * ge;
      ed += f.relax;
      if (NB > 0) {
        const float abg = bg_alpha(a, b * NB + k);
        al = al * in + abg * (1.0f - in);
      }
    } else {
      al = bg_alpha(a, b * NB + k);
    }
    alpha[k] = al;
    ins[k] = in;
  }
  en = warp_sum(en);
  ed = warp_sum(ed);
  if (lane == 0) { eik_num[b] = en; eik_den[b] = ed; }
  __syncwarp();
  transmittance_scan(alpha, T, NW, lane);
  __syncwarp();
  float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, wsum = 0.0f;
  for (int k = lane; k < NW; k += 32) {
    const float wk = alpha[k] * T[k];
    T[k] = wk;  // T now holds the weights
    weights[b * NW + k] = wk;
    wsum += wk;
    float r0, r1, r2;
    if (k < S) {
      const long long i = b * S + k;
      r0 = a.col[i * 3]; r1 = a.col[i * 3 + 1]; r2 = a.col[i * 3 + 2];
      if (NB > 0) {
        const long long j = b * NB + k;
        const float in = ins[k], om = 1.0f - in;
        r0 = r0 * in + a.rgb_bg[j * 3] * om;
        r1 = r1 * in + a.rgb_bg[j * 3 + 1] * om;
        r2 = r2 * in + a.rgb_bg[j * 3 + 2] * om;
      }
    } else {
      const long long j = b * NB + k;
      r0 = a.rgb_bg[j * 3]; r1 = a.rgb_bg[j * 3 + 1]; r2 = a.rgb_bg[j * 3 + 2];
    }
    c0 += r0 * wk; c1 += r1 * wk; c2 += r2 * wk;
  }
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); wsum = warp_sum(wsum);
  if (lane == 0) {
    if (a.bg_rgb) {
      c0 += a.bg_rgb[0] * (1.0f - wsum); c1 += a.bg_rgb[1] * (1.0f - wsum); c2 += a.bg_rgb[2] * (1.0f - wsum);
    }
    color[b * 3] = c0; color[b * 3 + 1] = c1; color[b * 3 + 2] = c2;
  }
  if (a.F > 0) {
    __syncwarp();
    for (int f0 = 0; f0 < a.F; f0 += 32) {
      const int f = f0 + lane;
      float acc = 0.0f;
      if (f < a.F) {
        for (int k = 0; k < NW; ++k) {
          float v;
          if (k < S) {
            v = a.feat[(b * S + k) * a.F + f];
            if (NB > 0) v = v * ins[k] + a.feat_bg[(b * NB + k) * a.F + f] * (1.0f - ins[k]);
          } else {
            v = a.feat_bg[(b * NB + k) * a.F + f];
          }
          acc += v * T[k];
        }
        dfeat[b * a.F + f] = acc;
      }
    }
  }
}

struct CompGrads {
  const float *d_color, *d_weights, *d_cdf, *d_dfeat, *d_eik_num;   // cotangents (nullable except d_color)
  float *d_sdf, *d_nrm, *d_col, *d_feat;                             // fine-sample gradients
  float *d_sigma_bg, *d_rgb_bg, *d_feat_bg, *d_dists_bg;             // background gradients (d_dists_bg nullable)
  float *d_var_partial;                                              // [B]
  float *d_dirs;                                                     // [B,3] nullable
};

__global__ void __launch_bounds__(RAY_WARPS * 32)
composite_bwd_kernel(CompArgs a, CompGrads g) {
  __shared__ float sa[RAY_WARPS][RAY_MAXN], sT[RAY_WARPS][RAY_MAXN], sin_[RAY_WARPS][RAY_MAXN],
      swb[RAY_WARPS][RAY_MAXN];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * RAY_WARPS + wid;
  if (b >= a.B) return;
  float* alpha = sa[wid];
  float* T = sT[wid];
  float* ins = sin_[wid];
  float* wb = swb[wid];
  const int S = a.S, NB = a.NB, NW = NB > 0 ? NB : S, F = a.F;
  const float inv_s = inv_s_of(a.variance);
  const float ox = a.o[b * 3], oy = a.o[b * 3 + 1], oz = a.o[b * 3 + 2];
  const float dx = a.d[b * 3], dy = a.d[b * 3 + 1], dz = a.d[b * 3 + 2];
  // recompute alpha' and the scan
  for (int k = lane; k < NW; k += 32) {
    float al, in = 0.0f;
    if (k < S) {
      FineEval f = eval_fine(a, b, k, inv_s, dx, dy, dz, ox, oy, oz);
      al = f.alpha;
      in = f.inside;
      if (NB > 0) {
        const float abg = bg_alpha(a, b * NB + k);
        al = al * in + abg * (1.0f - in);
      }
    } else {
      al = bg_alpha(a, b * NB + k);
    }
    alpha[k] = al;
    ins[k] = in;
  }
  __syncwarp();
  transmittance_scan(alpha, T, NW, lane);
  __syncwarp();
  const float C0 = g.d_color[b * 3], C1 = g.d_color[b * 3 + 1], C2 = g.d_color[b * 3 + 2];
  float bgdot = 0.0f;
  if (a.bg_rgb) bgdot = C0 * a.bg_rgb[0] + C1 * a.bg_rgb[1] + C2 * a.bg_rgb[2];
  // wbar_k (colour part) and the colour gradients
  for (int k = lane; k < NW; k += 32) {
    const float wk = alpha[k] * T[k];
    float r0, r1, r2;
    const float in = ins[k], om = 1.0f - in;
    if (k < S) {
      const long long i = b * S + k;
      r0 = a.col[i * 3]; r1 = a.col[i * 3 + 1]; r2 = a.col[i * 3 + 2];
      const float fin = (NB > 0) ? in : 1.0f;
      g.d_col[i * 3] = wk * C0 * fin; g.d_col[i * 3 + 1] = wk * C1 * fin; g.d_col[i * 3 + 2] = wk * C2 * fin;
      if (NB > 0) {
        const long long j = b * NB + k;
        r0 = r0 * in + a.rgb_bg[j * 3] * om;
        r1 = r1 * in + a.rgb_bg[j * 3 + 1] * om;
        r2 = r2 * in + a.rgb_bg[j * 3 + 2] * om;
        g.d_rgb_bg[j * 3] = wk * C0 * om; g.d_rgb_bg[j * 3 + 1] = wk * C1 * om; g.d_rgb_bg[j * 3 + 2] = wk * C2 * om;
      }
    } else {
      const long long j = b * NB + k;
      r0 = a.rgb_bg[j * 3]; r1 = a.rgb_bg[j * 3 + 1]; r2 = a.rgb_bg[j * 3 + 2];
      g.d_rgb_bg[j * 3] = wk * C0; g.d_rgb_bg[j * 3 + 1] = wk * C1; g.d_rgb_bg[j * 3 + 2] = wk * C2;
    }
    float v = C0 * r0 + C1 * r1 + C2 * r2 - bgdot;
    if (g.d_weights) v += g.d_weights[b * NW + k];
    wb[k] = v;
  }
  __syncwarp();
  // feature part of wbar and the feature gradients
  if (F > 0 && g.d_dfeat) {
    for (int k = 0; k < NW; ++k) {
      const float wk = alpha[k] * T[k];
      const float in = ins[k], om = 1.0f - in;
      float dot = 0.0f;
      for (int f = lane; f < F; f += 32) {
        const float D = g.d_dfeat[b * F + f];
        float v;
        if (k < S) {
          v = a.feat[(b * S + k) * F + f];
          const float fin = (NB > 0) ? in : 1.0f;
          g.d_feat[(b * S + k) * F + f] = wk * D * fin;
          if (NB > 0) {
            v = v * in + a.feat_bg[(b * NB + k) * F + f] * om;
            g.d_feat_bg[(b * NB + k) * F + f] = wk * D * om;
          }
        } else {
          v = a.feat_bg[(b * NB + k) * F + f];
          g.d_feat_bg[(b * NB + k) * F + f] = wk * D;
        }
        dot += D * v;
      }
      dot = warp_sum(dot);
      if (lane == 0) wb[k] += dot;
    }
    __syncwarp();
  } else if (F > 0) {
    for (int k = 0; k < NW; ++k)
      for (int f = lane; f < F; f += 32) {
        if (k < S) g.d_feat[(b * S + k) * F + f] = 0.0f;
        if (NB > 0) g.d_feat_bg[(b * NB + k) * F + f] = 0.0f;
      }
  }
  // suffix sums R_k = sum_{m>k} wbar_m w_m  (blocked reverse scan), then alpha'-bar into wb[]
  {
    const int per = (NW + 31) / 32;
    const int k0 = lane * per;
    float local = 0.0f;
    for (int i = 0; i < per; ++i) {
      int k = k0 + i;
      if (k < NW) local += wb[k] * alpha[k] * T[k];
    }
    float incl = local;  // inclusive suffix over lanes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      float v = __shfl_down_sync(0xffffffffu, incl, off);
      if (lane + off < 32) incl += v;
    }
    float run = incl - local;  // sum over lanes > this lane
    float ab[8];
    for (int i = per - 1; i >= 0; --i) {
      int k = k0 + i;
      if (k < NW) {
        const float wk = alpha[k] * T[k];
        ab[i] = wb[k] * T[k] - run / (1.0f - alpha[k] + 1e-7f);
        run += wb[k] * wk;
      }
    }
    __syncwarp();
    for (int i = 0; i < per; ++i) {
      int k = k0 + i;
      if (k < NW) wb[k] = ab[i];
    }
  }
  __syncwarp();
  // chain through the fine alpha and the background alpha
  const float Enum = g.d_eik_num ? g.d_eik_num[b] : 0.0f;
  float dvar = 0.0f, dd0 = 0.0f, dd1 = 0.0f, dd2 = 0.0f;
  for (int k = lane; k < NW; k += 32) {
    const float abar = wb[k];
    const float in = ins[k];
    float abar_bg = abar;
    if (k < S) {
      const long long i = b * S + k;
      FineEval f = eval_fine(a, b, k, inv_s, dx, dy, dz, ox, oy, oz);
      const float abar_f = (NB > 0) ? abar * in : abar;
      abar_bg = abar * (1.0f - in);
      const float rawbar = (f.raw >= 0.0f && f.raw <= 1.0f) ? abar_f : 0.0f;
      const float den = f.P + 1e-5f;
      float Pbar = rawbar * f.Nn / (den * den);
      const float Nbar = -rawbar / den;
      if (g.d_cdf) Pbar += g.d_cdf[i];
      const float tp = Pbar * f.P * (1.0f - f.P), tn = Nbar * f.Nn * (1.0f - f.Nn);
      const float epb = tp * inv_s, enb = tn * inv_s;
      dvar += tp * f.ep + tn * f.en;
      g.d_sdf[i] = epb + enb;
      const float icb = (enb - epb) * a.dists[i] * 0.5f;
      const float tcb = icb * (0.5f * (1.0f - a.cos_anneal) * (f.a1 > 0.0f ? 1.0f : 0.0f) +
                               a.cos_anneal * (-f.tc > 0.0f ? 1.0f : 0.0f));
      const float gx = a.nrm[i * 3], gy = a.nrm[i * 3 + 1], gz = a.nrm[i * 3 + 2];
      float ek = 0.0f;
      if (f.gnorm > 0.0f) ek = Enum * f.relax * 2.0f * (f.gnorm - 1.0f) / f.gnorm;
      g.d_nrm[i * 3] = tcb * dx + ek * gx;
      g.d_nrm[i * 3 + 1] = tcb * dy + ek * gy;
      g.d_nrm[i * 3 + 2] = tcb * dz + ek * gz;
      dd0 += tcb * gx; dd1 += tcb * gy; dd2 += tcb * gz;
    }
    if (NB > 0) {
      const long long j = b * NB + k;
      if (!a.dists_bg) {
        g.d_sigma_bg[j] = abar_bg;
      } else {
        const float sg = a.sigma_bg[j], dist = a.dists_bg[j];
        const float sp = softplus1(sg);
        const float ex = expf(-sp * dist);
        const float dsp = sg > 20.0f ? 1.0f : sigmoidf_(sg);
        g.d_sigma_bg[j] = abar_bg * ex * dist * dsp;
        if (g.d_dists_bg) g.d_dists_bg[j] = abar_bg * ex * sp;
      }
    }
  }
  dvar = warp_sum(dvar);
  dd0 = warp_sum(dd0); dd1 = warp_sum(dd1); dd2 = warp_sum(dd2);
  if (lane == 0) {
    // inv_s = clip(exp(10 v)); d inv_s / d v = 10 inv_s inside the clip
    const float raw = expf(a.variance[0] * 10.0f);
    const float pass = (raw >= 1e-6f && raw <= 1e6f) ? 1.0f : 0.0f;
    g.d_var_partial[b] = dvar * 10.0f * inv_s * pass;
    if (g.d_dirs) { g.d_dirs[b * 3] = dd0; g.d_dirs[b * 3 + 1] = dd1; g.d_dirs[b * 3 + 2] = dd2; }
  }
}

}  // namespace vdn
using namespace vdn;

extern "C" int vdn_ray_points(const float* o, const float* d, const float* z, long long B, int n, float* pts,
                              void* stream) {
  long long tot = B * n;
  if (tot <= 0) return 0;
  VDN_LAUNCH(ray_points_kernel, (unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream, o, d, z, B, n, pts);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_upsample_step(const float* o, const float* d, const float* z_in, int n, const float* sdf_prev,
                                 int n_prev, const float* sdf_new, int n_new_prev, const unsigned char* perm_prev,
                                 float inv_s, int n_imp, long long B, float* z_out, float* sdf_out,
                                 unsigned char* perm_out, float* new_z, float* new_pts, long long* inds_out,
                                 void* stream) {
  if (B <= 0) return 0;
  if (n < 2 || n_imp < 1 || n_imp > 32 || n + n_imp > RAY_MAXN) return (int)cudaErrorInvalidValue;
  if (perm_prev && (!sdf_new || n_prev + n_new_prev != n)) return (int)cudaErrorInvalidValue;
  if (!perm_prev && n_prev != n) return (int)cudaErrorInvalidValue;
  unsigned blocks = (unsigned)((B + RAY_WARPS - 1) / RAY_WARPS);
  VDN_LAUNCH(upsample_kernel, blocks, RAY_WARPS * 32, 0, (cudaStream_t)stream, o, d, z_in, n, sdf_prev, n_prev, sdf_new,
                                                                      n_new_prev, perm_prev, inv_s, n_imp, B, z_out,
                                                                      sdf_out, perm_out, new_z, new_pts, inds_out);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_merge_sorted(const float* za, int n, const float* zb, int m, long long B, float* z_out,
                                unsigned char* perm, void* stream) {
  if (B <= 0) return 0;
  if (n < 0 || m < 0 || n + m > RAY_MAXN || n + m < 1) return (int)cudaErrorInvalidValue;
  unsigned blocks = (unsigned)((B + RAY_WARPS - 1) / RAY_WARPS);
  VDN_LAUNCH(merge_sorted_kernel, blocks, RAY_WARPS * 32, 0, (cudaStream_t)stream, za, n, zb, m, B, z_out, perm);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_fine_prep(const float* o, const float* d, const float* z, float sample_dist, long long B, int S,
                             float* dists, float* mid_z, float* pts, void* stream) {
  long long tot = B * S;
  if (tot <= 0) return 0;
  VDN_LAUNCH(fine_prep_kernel, (unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream, o, d, z, sample_dist, B, S, dists,
                                                                                  mid_z, pts);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_bg_prep(const float* o, const float* d, const float* z_fine, int S, const float* z_outside, int NO,
                           float sample_dist, long long B, float* dists, float* mid_z, float* pts4, void* stream) {
  if (B <= 0) return 0;
  if (S < 1 || NO < 0 || S + NO > RAY_MAXN) return (int)cudaErrorInvalidValue;
  unsigned blocks = (unsigned)((B + RAY_WARPS - 1) / RAY_WARPS);
  VDN_LAUNCH(bg_prep_kernel, blocks, RAY_WARPS * 32, 0, (cudaStream_t)stream, o, d, z_fine, S, z_outside, NO, sample_dist, B,
                                                                     dists, mid_z, pts4);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

static int fill_comp_args(CompArgs* a, long long B, int S, int NB, int F, const float* o, const float* d,
                          const float* mid_z, const float* dists, const float* sdf, const float* nrm,
                          const float* col, const float* feat, const float* sigma_bg, const float* rgb_bg,
                          const float* feat_bg, const float* dists_bg, const float* variance, const float* bg_rgb,
                          float cos_anneal) {
  if (S < 1 || S > RAY_MAXN || (NB != 0 && (NB < S || NB > RAY_MAXN)) || F < 0) return 1;
  if ((NB + 31) / 32 > 8 || (S + 31) / 32 > 8) return 1;
  a->B = B; a->S = S; a->NB = NB; a->F = F;
  a->o = o; a->d = d; a->mid_z = mid_z; a->dists = dists;
  a->sdf = sdf; a->nrm = nrm; a->col = col; a->feat = feat;
  a->sigma_bg = sigma_bg; a->rgb_bg = rgb_bg; a->feat_bg = feat_bg; a->dists_bg = dists_bg;
  a->variance = variance; a->bg_rgb = bg_rgb; a->cos_anneal = cos_anneal;
  return 0;
}

extern "C" int vdn_composite_fwd(long long B, int S, int NB, int F, const float* o, const float* d,
                                 const float* mid_z, const float* dists, const float* sdf, const float* nrm,
                                 const float* col, const float* feat, const float* sigma_bg, const float* rgb_bg,
                                 const float* feat_bg, const float* dists_bg, const float* variance,
                                 const float* bg_rgb, float cos_anneal, float* weights, float* cdf, float* inside,
                                 float* color, float* dfeat, float* eik_num, float* eik_den, void* stream) {
  if (B <= 0) return 0;
  CompArgs a;
  if (fill_comp_args(&a, B, S, NB, F, o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg,
                     variance, bg_rgb, cos_anneal))
    return (int)cudaErrorInvalidValue;
  unsigned blocks = (unsigned)((B + RAY_WARPS - 1) / RAY_WARPS);
  VDN_LAUNCH(composite_fwd_kernel, blocks, RAY_WARPS * 32, 0, (cudaStream_t)stream, a, weights, cdf, inside, color, dfeat,
                                                                           eik_num, eik_den);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

extern "C" int vdn_composite_bwd(long long B, int S, int NB, int F, const float* o, const float* d,
                                 const float* mid_z, const float* dists, const float* sdf, const float* nrm,
                                 const float* col, const float* feat, const float* sigma_bg, const float* rgb_bg,
                                 const float* feat_bg, const float* dists_bg, const float* variance,
                                 const float* bg_rgb, float cos_anneal, const float* d_color, const float* d_weights,
                                 const float* d_cdf, const float* d_dfeat, const float* d_eik_num, float* d_sdf,
                                 float* d_nrm, float* d_col, float* d_feat, float* d_sigma_bg, float* d_rgb_bg,
                                 float* d_feat_bg, float* d_dists_bg, float* d_var_partial, float* d_dirs,
                                 void* stream) {
  if (B <= 0) return 0;
  CompArgs a;
  if (fill_comp_args(&a, B, S, NB, F, o, d, mid_z, dists, sdf, nrm, col, feat, sigma_bg, rgb_bg, feat_bg, dists_bg,
                     variance, bg_rgb, cos_anneal))
    return (int)cudaErrorInvalidValue;
  CompGrads g;
  g.d_color = d_color; g.d_weights = d_weights; g.d_cdf = d_cdf; g.d_dfeat = d_dfeat; g.d_eik_num = d_eik_num;
  g.d_sdf = d_sdf; g.d_nrm = d_nrm; g.d_col = d_col; g.d_feat = d_feat;
  g.d_sigma_bg = d_sigma_bg; g.d_rgb_bg = d_rgb_bg; g.d_feat_bg = d_feat_bg; g.d_dists_bg = d_dists_bg;
  g.d_var_partial = d_var_partial; g.d_dirs = d_dirs;
  unsigned blocks = (unsigned)((B + RAY_WARPS - 1) / RAY_WARPS);
  VDN_LAUNCH(composite_bwd_kernel, blocks, RAY_WARPS * 32, 0, (cudaStream_t)stream, a, g);
  return (int)(cudaError_t)::vdn::take_launch_error();
}

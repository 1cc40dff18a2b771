/* vdn_b200.h -- C ABI of libvdn_b200.so, the sm_100a kernels behind the VDN-NeRF neural-SDF
 * volume-rendering hot path.
 *
 * The reference (BoifZ/VDN-NeRF) is pure Python/PyTorch and has no FFI layer; its
 * boundary is the Python class surface of dpt_models/{embedder,fields,renderer}.py.  This library is what the
 * host-side mirror of those classes (vdn_nerf_b200/*.py) binds with ctypes; each entry point names the
 * reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data owned by the caller (torch tensors) unless noted
 *     "host"; the library never allocates device memory and never synchronises;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - return value: 0 on success, otherwise a cudaError_t value (invalid arguments -> cudaErrorInvalidValue);
 *   - N, B are counts of points / rays; row-major everywhere; "ld" = leading dimension in floats.
 *
 * Packed parameters.  A network's effective weights live in one fp32 buffer laid out by vdn_mlp_layout():
 * per layer W[out_ld,in_ld], W^T[in_ld,out_ld], bias[out_ld] with in_ld/out_ld rounded up to 16 and zero
 * padding, followed by tf32 SWIZZLE_128B tile images of W and W^T for the tcgen05 kernels (csrc/mlp_layout.cuh).  A packed gradient buffer has the same layout (W and bias regions are accumulated into).
 */
#ifndef VDN_B200_H
#define VDN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------- */
/* ABI version of this header; bumped on any signature change. */
int vdn_abi_version(void);
/* Number of this library's kernel launches enqueued by the calling process so far (for bench.py's
 * gpu_launches claim). */
long long vdn_launch_count(void);
/* cudaGetErrorString for the codes this library returns. */
const char* vdn_error_string(int code);

/* Arithmetic mode of the MLP contractions: 0 = exact fp32 (FFMA kernels; parity <= 1e-5), 1 = tensor cores (tcgen05.mma
 * with fp32 accumulation in TMEM: fp16 operands in the forward / normals chains, bf16 in the backward chains, tf32 in the
 * layer-wise kernels; parity <= 2e-3 on colour and normals).  Process-wide.
 * A tcgen05 kernel raises a 4-byte device flag if one of its bounded barrier waits times out.  The flag is memory of the
 * CALLER (vdn_set_fault_flag; the library allocates nothing); vdn_tc_fault() reads it synchronously,
 * vdn_tc_fault_async() enqueues a copy to (pinned) host memory on a stream. */
int vdn_set_mode(int mode);
int vdn_get_mode(void);
int vdn_set_fault_flag(int* device_flag);
int vdn_tc_fault(void);
int vdn_tc_fault_async(int* host_dst /*host, pinned*/, void* stream);
/* Tensor-core mode only: 1 (default) runs the training passes of the three networks as fused layer chains
 * (csrc/chain_engine.cuh: activations resident in tensor memory, 16-bit saved tensors, grouped TMA weight gradient),
 * 0 runs them layer by layer (csrc/gemm_tc.cuh).  Must not change between a forward and its backward. */
int vdn_set_chain(int on);
int vdn_get_chain(void);

/* Debug aid: when device_buf (8192 int64 of device memory) is non-null, the tcgen05 kernels record time stamps there:
 * CTA 0 of a layer-wise GEMM launch its pipeline events as clock64() in entries [0, 512) and every CTA (first 1500)
 * its %globaltimer start / end and SM id in entries [1024, 7024) (csrc/gemm_tc.cuh; the environment variable
 * VDN_DBG_LAUNCH=n restricts the recording to the n-th launch after the buffer was installed); CTA 0 of the fused
 * chain kernel its per-phase events in entries [0, 512) (csrc/sdf_chain_tc.cuh).  Pass null to switch it off.
 * Readers: tools/diag_timeline.py, tools/diag_ctas.py, tools/diag_chain.py. */
int vdn_debug_timeline(long long* device_buf);

/* Measurement aid for bench.py: when enabled, CUDA events bracket every launch of a kernel family on its
 * stream (0 = FFMA gemm_nt, 1 = weight-gradient gemm_tn (+ reduce), 2 = layer-wise tcgen05 gemm_nt, 3 = fused SDF chain
 * kernel); vdn_prof_read sums the recorded durations (ms), the number of spans and the executed FLOPs,
 * vdn_prof_read_bytes the algorithmic HBM bytes (every operand / output element once).  Enabling/disabling clears
 * the record. */
int vdn_prof_enable(int on);
int vdn_prof_read(int family, double* ms /*host*/, long long* spans /*host*/, double* flops /*host*/);
int vdn_prof_read_bytes(int family, double* bytes /*host*/);

/* ---- packed parameters (weight_norm: fields.py:65-66, 141-142; nn.Linear: fields.py:303-318) ------------ */
long long vdn_mlp_layout(int L, const int* in_dims /*host*/, const int* out_dims /*host*/, long long* off_w /*host*/,
                         long long* off_wt /*host*/, long long* off_b /*host*/);
/* v/g/b/rows are host arrays of 2*L entries: each packed layer stacks up to two parameter sets by rows
 * (second entry null / 0 rows when unused).  g == null means a plain (not weight-normed) weight. */
int vdn_mlp_pack(int L, const int* in_dims, const int* out_dims, const float* const* v, const float* const* g,
                 const float* const* b, const int* rows, const int* rot /*host, nullable: per-layer input-column
                 rotation*/, const int* orot /*host, nullable: per-layer output rotation of the fp16 tile images
                 (vdn_sdf_layer_orot / vdn_nerf_layer_orot)*/, float* packed, void* stream);
/* Weight-norm backward + un-padding: packed gradient -> d weight_v, d weight_g, d bias (overwritten). */
int vdn_mlp_unpack_grads(int L, const int* in_dims, const int* out_dims, const float* const* v,
                         const float* const* g, const int* rows, const int* rot, const float* dpacked,
                         float* const* dv, float* const* dg, float* const* db, void* stream);

/* ---- SDFNetwork (fields.py:9-108).  cfg (host) = {d_in, multires, d_hidden, n_layers, d_out, skip_layer|-1} */
int vdn_sdf_layer_dims(const int* cfg, int* in_dims, int* out_dims); /* returns number of linear layers */
/* Output rotation per layer that vdn_mlp_pack must be given for this network (the stacked [sdf ; feature] head presents
 * its features first in the fp16 tile images); returns the number of layers. */
int vdn_sdf_layer_orot(const int* cfg, int* orot);
long long vdn_sdf_blob_floats(const int* cfg, long long N, int save);
long long vdn_sdf_blobg_floats(const int* cfg, long long N);
long long vdn_sdf_bwd_ws_floats(const int* cfg, long long N);
/* SDFNetwork.forward / .sdf (fields.py:72-92): sdf[N] (stride lds) and, when feat != null, feat[N, d_out-1]
 * (ld ldf).
 * blob: scratch of vdn_sdf_blob_floats(cfg, N, save) floats; save=1 keeps every pre-activation for
 * vdn_sdf_normals / vdn_sdf_backward. */
int vdn_sdf_forward(const int* cfg, float scale, const float* packed, const float* x, long long N, float* sdf,
                    int lds, float* feat, int ldf, float* blob, int save, void* stream);
/* SDFNetwork.gradient (fields.py:97-108) without autograd: normals[N, d_in] = d sdf / d x.
 * blob: the save=1 blob of the forward on the same x; blobg: vdn_sdf_blobg_floats floats (kept for backward). */
int vdn_sdf_normals(const int* cfg, float scale, const float* packed, const float* x, long long N, const float* blob,
                    float* blobg, float* normals, void* stream);
/* Backward of (sdf, feat, normals) w.r.t. the packed parameters (accumulated into dpacked) and, when
 * d_x != null, the points (overwritten).  Replaces autograd's double backward through fields.py:97-108.
 * Cotangents may be null.  ws: vdn_sdf_bwd_ws_floats floats. */
int vdn_sdf_backward(const int* cfg, float scale, const float* packed, const float* x, long long N, const float* blob,
                     const float* blobg, const float* d_sdf, int lds, const float* d_feat, int ldf,
                     const float* d_normals, float* dpacked, float* d_x, float* ws, void* stream);
/* extract_fields (renderer.py:10-30) for the x-slab [i0, i1): u_slab[(i-i0), j, k] = out_mul * sdf(xs[i], ys[j],
 * zs[k]).  pts: (i1-i0)*ny*nz*3 floats scratch; blob: save=0 blob for that many points. */
int vdn_grid_sdf(const int* cfg, float scale, const float* packed, const float* xs, const float* ys, const float* zs,
                 int ny, int nz, int i0, int i1, float out_mul, float* u_slab, float* pts, float* blob, void* stream);

/* ---- RenderingNetwork (fields.py:112-176).
 * cfg (host) = {d_feature, mode(0 idr,1 no_view_dir,2 no_normal), d_out, d_hidden, n_layers, multires_view,
 *               squeeze_out} */
int vdn_rendernet_layer_dims(const int* cfg, int* in_dims, int* out_dims);
/* Input-column rotation per layer for vdn_mlp_pack / vdn_mlp_unpack_grads: layer 0 is packed as [feature | extras], and
 * that is also the column order of the saved input row and of d_cin below. */
int vdn_rendernet_layer_rot(const int* cfg, int* rot);
long long vdn_rendernet_blob_floats(const int* cfg, long long N);
long long vdn_rendernet_bwd_ws_floats(const int* cfg, long long N);
int vdn_rendernet_forward(const int* cfg, const float* packed, const float* points, const float* normals,
                          const float* view_dirs, const float* feats, int ldf, long long N, float* out, float* blob,
                          void* stream);
/* d_cin (nullable): [N, round_up(in0,16)] cotangent of the input row of fields.py:154 in the ROTATED column order
 * [feature | points | view embedding | normals] (vdn_rendernet_layer_rot). */
int vdn_rendernet_backward(const int* cfg, const float* packed, long long N, const float* blob, const float* out,
                           const float* d_out, float* dpacked, float* d_cin, float* ws, void* stream);

/* ---- NeRF background field (fields.py:264-355).
 * cfg (host) = {D, W, d_in, d_in_view, multires, multires_view, skip|-1, rgb_dims, dpt_dim(0 = no depth head)}
 * Packed layers: pts_linears[0..D-1], [alpha_linear;feature_linear], views_linears[0], [rgb_linear;dpt_linear]. */
int vdn_nerf_layer_dims(const int* cfg, int* in_dims, int* out_dims);
/* Output rotation per layer for vdn_mlp_pack (stacked [alpha ; feature] head: features first in the 16-bit images). */
int vdn_nerf_layer_orot(const int* cfg, int* orot);
long long vdn_nerf_blob_floats(const int* cfg, long long N);
long long vdn_nerf_bwd_ws_floats(const int* cfg, long long N);
int vdn_nerf_forward(const int* cfg, const float* packed, const float* pts, const float* views, long long N,
                     float* sigma, float* rgb, float* dpt, float* blob, void* stream);
int vdn_nerf_backward(const int* cfg, const float* packed, const float* pts, const float* views, long long N,
                      const float* blob, const float* d_sigma, const float* d_rgb, const float* d_dpt, float* dpacked,
                      float* d_pts, float* d_views, float* ws, void* stream);

/* ---- embedder (embedder.py:11-36) -------------------------------------------------------------------- */
int vdn_embed_fwd(const float* x, long long N, int d, int multires, float* out /*[N, d*(1+2*multires)]*/,
                  void* stream);
int vdn_embed_bwd(const float* x, long long N, int d, int multires, const float* d_out, float* d_x, void* stream);

/* ---- per-ray kernels (renderer.py) ------------------------------------------------------------------- */
/* pts[b,k,:] = o[b] + d[b] * z[b,k]   (renderer.py:150, 196, 369) */
int vdn_ray_points(const float* o, const float* d, const float* z, long long B, int n, float* pts, void* stream);
/* One iteration of the up-sampling loop (renderer.py:372-384): gathers the merged sdf of the previous
 * iteration (perm_prev/sdf_new null on the first), runs up_sample + sample_pdf(det=True) (renderer.py:44-74,
 * 147-191) and the cat+sort of cat_z_vals (193-198).  Outputs: z_out[B,n+n_imp] sorted, sdf_out[B,n] (nullable),
 * perm_out[B,n+n_imp] (source index of each sorted sample; >= n means new sample), new_z[B,n_imp],
 * new_pts[B*n_imp,3] (nullable), inds_out[B,n_imp] (int64 searchsorted result, nullable). */
int vdn_upsample_step(const float* o, const float* d, const float* z_in, int n, const float* sdf_prev, int n_prev,
                      const float* sdf_new, int n_new_prev, const unsigned char* perm_prev, float inv_s, int n_imp,
                      long long B, float* z_out, float* sdf_out, unsigned char* perm_out, float* new_z, float* new_pts,
                      long long* inds_out, void* stream);
/* cat + sort of cat_z_vals (renderer.py:197-198) as a stable merge: za[B,n] sorted, zb[B,m] arbitrary;
 * perm[B,n+m] = source index of each output sample (< n from za, >= n from zb). */
int vdn_merge_sorted(const float* za, int n, const float* zb, int m, long long B, float* z_out, unsigned char* perm,
                     void* stream);
/* Section lengths, mid points and fine sample points (renderer.py:228-237). */
int vdn_fine_prep(const float* o, const float* d, const float* z, float sample_dist, long long B, int S, float* dists,
                  float* mid_z, float* pts, void* stream);
/* Merge fine and outside samples (renderer.py:389-391) and build the inverted-sphere inputs (102-120). */
int vdn_bg_prep(const float* o, const float* d, const float* z_fine, int S, const float* z_outside, int NO,
                float sample_dist, long long B, float* dists, float* mid_z, float* pts4, void* stream);
/* Sigmoid-CDF alpha, background blend, transmittance scan, compositing and Eikonal sums (renderer.py:262-315).
 * NB = 0 (no background) or the merged sample count (>= S); F = feature width (0 = none). */
int vdn_composite_fwd(long long B, int S, int NB, int F, const float* o, const float* d, const float* mid_z,
                      const float* dists, const float* sdf, const float* nrm, const float* col, const float* feat,
                      const float* sigma_bg, const float* rgb_bg, const float* feat_bg, const float* dists_bg,
                      const float* variance, const float* bg_rgb, float cos_anneal, float* weights, float* cdf,
                      float* inside, float* color, float* dfeat, float* eik_num, float* eik_den, void* stream);
int vdn_composite_bwd(long long B, int S, int NB, int F, const float* o, const float* d, const float* mid_z,
                      const float* dists, const float* sdf, const float* nrm, const float* col, const float* feat,
                      const float* sigma_bg, const float* rgb_bg, const float* feat_bg, const float* dists_bg,
                      const float* variance, const float* bg_rgb, float cos_anneal, const float* d_color,
                      const float* d_weights, const float* d_cdf, const float* d_dfeat, const float* d_eik_num,
                      float* d_sdf, float* d_nrm, float* d_col, float* d_feat, float* d_sigma_bg, float* d_rgb_bg,
                      float* d_feat_bg, float* d_dists_bg, float* d_var_partial, float* d_dirs, void* stream);

/* ---- callers on either side of the path (SURVEY.md 8(f) "next" rows) ------------------------------------ */
/* torch.optim.Adam's update (dpt_runner.py:88, 251-253; no weight decay, no amsgrad) for n_tensors parameter tensors in one
 * launch.  params / grads / exp_avg / exp_avg_sq / numel are HOST arrays of n_tensors entries (<= 96) holding device
 * pointers and element counts (numel 0 skips a tensor); hyper_dev is a DEVICE array {lr, beta1, beta2, eps, 1 - beta1,
 * 1 - beta2, then for every tensor i: 1 - beta1^t_i, 1 - beta2^t_i} (torch keeps one step counter per parameter). */
int vdn_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const int* numel, const float* hyper_dev, void* stream);
/* Masked L1 colour loss of the driver (dpt_runner.py:228-232): sums[3] = {sum |c - t| m, sum ((c - t) m)^2, sum m},
 * d_color[B,3] = sign(c - t) m (mask nullable: ones). */
int vdn_color_loss(const float* color, const float* true_rgb, const float* mask, long long B, float* sums, float* d_color,
                   void* stream);
/* Rays of one camera (dpt_models/poses.py:189-212): rays_d = R normalize(Kinv [px, py, 1]), rays_o = t with pose = [R | t]
 * (3 x 4, device), Kinv 3 x 3 (device); the backward returns the cotangent of the pose (12 floats, overwritten). */
int vdn_raygen_fwd(const float* px, const float* py, long long B, const float* kinv, const float* pose, float* rays_o,
                   float* rays_d, void* stream);
int vdn_raygen_bwd(const float* px, const float* py, long long B, const float* kinv, const float* d_rays_o,
                   const float* d_rays_d, float* d_pose, void* stream);

/* Marching cubes on the device-resident field u[nx, ny, nz] (replaces the host library call at renderer.py:36; inside =
 * u > threshold).  vdn_mc_count writes the triangle count of every cell ((nx-1)(ny-1)(nz-1) ints); the caller forms the
 * exclusive prefix sum `offsets`; vdn_mc_emit writes, per triangle corner, the key of the cut grid edge (3 * linear index
 * of its lower end point + axis) and the interpolated position in grid-index coordinates.  tri_count[256],
 * tri_table[256*16], edge_corner[12], edge_axis[12]: the generated case table (vdn_nerf_b200/mcubes_table.py), device. */
int vdn_mc_count(const float* u, int nx, int ny, int nz, float threshold, const int* tri_count, int* counts, void* stream);
int vdn_mc_emit(const float* u, int nx, int ny, int nz, float threshold, const int* tri_count, const int* tri_table,
                const int* edge_corner, const int* edge_axis, const long long* offsets, long long* keys, float* pos,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDN_B200_H */

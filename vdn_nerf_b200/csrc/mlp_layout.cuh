// Packed parameter layout shared by every MLP kernel family.
//
// One contiguous fp32 buffer per network and per optimiser step holds, for every linear layer l:
//   W_l    [out_ld, in_ld]    effective weight (weight-norm applied: g * v / ||v||_row), zero padded
//   WT_l   [in_ld, out_ld]    its transpose (the dgrad / input-gradient passes read it as the "B" operand)
//   b_l    [out_ld]
//   IW_l   [ceil(in_dim/32)][out_ld][32]   W_l  as tcgen05 operand tiles: per 32-column K block, the rows in the
//   IWT_l  [ceil(out_dim/32)][in_ld][32]   W_l^T  SWIZZLE_128B K-major shared-memory image (tc_common.cuh), values
//                                          rounded to tf32 - one linear cp.async.bulk per tile, no tensor map
//   IH_l   [ceil(in_dim/64)][out_ld][64 halfs]   W_l rounded to fp16 (10-bit mantissa like tf32), same K-major
//                                          SWIZZLE_128B image with 64-element K blocks, for the kind::f16 chain kernels
//   IHT_l  [ceil(out_dim/64)][in_ld][64 halfs]  W_l^T rounded to fp16, same image (rows = input index, K = output index):
//                                          the B operand of the fused normals / backward chains
// The fp16 images may rotate the OUTPUT index (orot): image position j holds output (j + orot) mod out_dim, so a
// 257-wide stacked head [scalar ; 256 features] presents the features as positions 0..255 (one N = 256 MMA, K blocks
// 0..3 of the transposed image) and the scalar as position 256.
// with in_ld = round_up(in_dim, 16), out_ld = round_up(out_dim, 16).  The gradient buffer of a network uses
// the W and b regions of the same layout, so weight-norm backward is one kernel over the whole network.
// A layer may rotate its input columns (rot): packed column (c - rot) mod in_dim holds source column c; the
// NeRF skip layer uses it to store [hidden | embedding] instead of the reference's [embedding | hidden].
#pragma once
#include "common.cuh"

#define VDN_MAX_LAYERS 16

namespace vdn {

struct MlpLayout {
  int L;
  int in_dim[VDN_MAX_LAYERS], out_dim[VDN_MAX_LAYERS];
  int in_ld[VDN_MAX_LAYERS], out_ld[VDN_MAX_LAYERS];
  long long off_w[VDN_MAX_LAYERS], off_wt[VDN_MAX_LAYERS], off_b[VDN_MAX_LAYERS];
  long long off_iw[VDN_MAX_LAYERS], off_iwt[VDN_MAX_LAYERS], off_ih[VDN_MAX_LAYERS], off_iht[VDN_MAX_LAYERS];
  long long total;  // floats
};

inline int make_layout(int L, const int* in_dims, const int* out_dims, MlpLayout* ly) {
  if (L < 1 || L > VDN_MAX_LAYERS) return 1;
  ly->L = L;
  long long off = 0;
  for (int l = 0; l < L; ++l) {
    if (in_dims[l] < 1 || out_dims[l] < 1) return 1;
    ly->in_dim[l] = in_dims[l];
    ly->out_dim[l] = out_dims[l];
    ly->in_ld[l] = round_up(in_dims[l], 16);
    ly->out_ld[l] = round_up(out_dims[l], 16);
    ly->off_w[l] = off;
    off += (long long)ly->out_ld[l] * ly->in_ld[l];
    ly->off_wt[l] = off;
    off += (long long)ly->in_ld[l] * ly->out_ld[l];
    ly->off_b[l] = off;
    off += ly->out_ld[l];
  }
  for (int l = 0; l < L; ++l) {
    off = (off + 255) / 256 * 256;  // 1024-byte aligned tiles
    ly->off_iw[l] = off;
    off += (long long)((in_dims[l] + 31) / 32) * ly->out_ld[l] * 32;
    ly->off_iwt[l] = off;
    off += (long long)((out_dims[l] + 31) / 32) * ly->in_ld[l] * 32;
  }
  for (int l = 0; l < L; ++l) {
    off = (off + 255) / 256 * 256;
    ly->off_ih[l] = off;
    off += (long long)((in_dims[l] + 63) / 64) * ly->out_ld[l] * 32;   // 64 halfs = 32 floats per row and K block
    off = (off + 255) / 256 * 256;
    ly->off_iht[l] = off;
    off += (long long)((out_dims[l] + 63) / 64) * ly->in_ld[l] * 32;
  }
  ly->total = off;
  return 0;
}

}  // namespace vdn

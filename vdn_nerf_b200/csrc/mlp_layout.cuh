// Packed parameter layout shared by every MLP kernel family.
//
// One contiguous fp32 buffer per network and per optimiser step holds, for every linear layer l:
//   W_l   [out_ld, in_ld]   effective weight (weight-norm applied: g * v / ||v||_row), zero padded
//   WT_l  [in_ld, out_ld]   its transpose (the dgrad / input-gradient passes read it as the "B" operand)
//   b_l   [out_ld]
// with in_ld = round_up(in_dim, 16), out_ld = round_up(out_dim, 16).  The gradient buffer of a network has
// the same layout (the WT region is unused), so weight-norm backward is one kernel over the whole network.
#pragma once
#include "common.cuh"

#define VDN_MAX_LAYERS 16

namespace vdn {

struct MlpLayout {
  int L;
  int in_dim[VDN_MAX_LAYERS], out_dim[VDN_MAX_LAYERS];
  int in_ld[VDN_MAX_LAYERS], out_ld[VDN_MAX_LAYERS];
  long long off_w[VDN_MAX_LAYERS], off_wt[VDN_MAX_LAYERS], off_b[VDN_MAX_LAYERS];
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
  ly->total = off;
  return 0;
}

}  // namespace vdn

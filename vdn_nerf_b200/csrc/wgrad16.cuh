// Grouped weight-gradient kernel of the fused chains (tensor-core mode): ONE launch accumulates the weight (and bias)
// gradients of every layer of a network,
//
//     dW_l[out, in] += scale * sum_p  X_l[p, out] * Y_l[p, in]         (W-bar += z-bar^T u, delta^T q-bar)
//
// from the fp16 operand tensors the chain kernels (chain_engine.cuh) left in HBM: X = a cotangent or the normals-pass
// delta, Y = the layer input or a phase-1 cotangent, tile-blocked.  Exactly one operand of every segment is a cotangent
// and carries the call's loss scale sigma (chain_engine.cuh), so the flush multiplies by 1 / sigma (Args::sigma[1]).
// A "job" is one (layer, output-row block) with up to two (X, Y) segments that share an accumulator
// (W-bar_l = [z-bar ; delta]^T [u ; q-bar]); the batch is split over the CTAs so that (jobs x splits) fills one wave of
// the 148 SMs, every CTA streaming its points exactly once:
//   * operands arrive by TMA TENSOR MAPS (cp.async.bulk.tensor.3d, UTMALDG) over the tile-blocked 16-bit tensors
//     ([tile][8-column group][128 rows][8], chain_engine.cuh): one box = 64 points x all needed column groups, so a
//     64-point stage is two TMA instructions (X and Y, up to 32 KB each), three stages in flight; column groups and
//     tiles outside a tensor are zero-filled by the TMA unit;
//   * the reduction runs over the points, so both operands are MN-major for tcgen05.mma kind::f16; the box lands in
//     shared memory as [column group][64 rows][16 bytes], which is exactly the canonical MN-major NO-SWIZZLE
//     ("interleaved") operand layout: 8 x 8 core matrices of 128 contiguous bytes, 128 bytes between K groups,
//     1024 bytes between column groups;
//   * a 256 x 256 fp32 accumulator (two 128-lane halves x 256 TMEM columns) lives in tensor memory for the whole split
//     and is added to the packed gradient with red.global.add.v4.f32 (no partial slabs, no reduce launch);
//   * four warps sum the columns of X from the staged tiles for the bias gradient while the tensor pipe works.
// The kernel is bandwidth bound by construction (128 FLOP per operand byte): the figure of merit is HBM GB/s.
#pragma once
#include "chain_engine.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace vdn {
namespace wg {

constexpr int MAX_MAPS = 56;
constexpr int MAX_JOBS = 40;
constexpr int MAX_UNITS = 160;
constexpr int STAGES = 3;
constexpr uint32_t STAGE_BYTES = 65536;
constexpr int THREADS = 192;      // warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2-5: bias sums + flush
constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024;

struct Job {
  int nseg;                     // 1 or 2 segments accumulated into the same tile
  int xmap[2], ymap[2];         // tensor-map indices
  int xcol[2], ycol[2];         // first column (element) inside the tensor
  int m_tiles;                  // 128-row halves of the output block (1 or 2)
  int n_mma;                    // MMA N (multiple of 16, <= 256)
  int rows, cols;               // valid extent of the output block
  long long dw_off;             // float offset of dW[row 0 of the block][col 0] in the packed gradient
  int dw_ld;
  float scale;
  long long db_off;             // float offset of the bias gradient of row 0 of the block, -1: none
  float db_scale;
};
struct Unit { int job, c0, c1; };   // 64-point chunks [c0, c1)

struct alignas(64) Args {
  CUtensorMap maps[MAX_MAPS];
  Job jobs[MAX_JOBS];
  Unit units[MAX_UNITS];
  float* dpacked;
  const float* sigma;           // device {sigma, 1 / sigma} of the backward call (null: 1)
};

// fp16 x fp16 -> fp32, both operands MN-major
__device__ __forceinline__ uint32_t idesc_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// MN-major operand without swizzle: core matrices of 8 K-rows x 16 bytes (128 contiguous bytes); the next 8 K-rows are
// lbo bytes further, the next 8 M/N elements sbo bytes further
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// The tile-blocked tensor [tile][column group][128 rows][8 x 16 bit] is described to the TMA unit as a 3-D tensor of
// 32-bit words {512 words = 128 rows x 16 bytes, column groups, tiles}: a box {256 words = 64 rows, box column groups, 1}
// at (4 * row0, cg0, tile) moves 1 KB contiguous runs (a 16-byte innermost box dimension would be the slowest shape the
// TMA unit handles).
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int row0, int cg0, int tile, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tm), "r"(row0 * 4), "r"(cg0), "r"(tile), "r"(bar)
               : "memory");
}

static __global__ void __launch_bounds__(THREADS, 1) wgrad16_kernel(const __grid_constant__ Args a, int* __restrict__ fault) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[STAGES], empty[STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const Unit un = a.units[blockIdx.x];
  const Job& jb = a.jobs[un.job];
  const int nchunks = un.c1 - un.c0;
  const int x_cg = jb.m_tiles * 16, y_cg = (jb.n_mma + 7) >> 3;     // column groups per stage of X and Y
  const int nit = nchunks * jb.nseg;           // stages to process: chunk-major, segment-minor

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1 + 4); }
    mbar_init(smem_u32(&acc_full), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  if (tid == 0) {
    // ================= TMA producer =================
    const uint32_t bytes = (uint32_t)(x_cg + y_cg) * 1024u;
    for (int it = 0; it < nit && ok; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      const int seg = it % jb.nseg, chunk = un.c0 + it / jb.nseg;
      ok = mbar_wait_backoff(smem_u32(&empty[s]), ph ^ 1, 64);
      const uint32_t base = s0 + (uint32_t)s * STAGE_BYTES, bar = smem_u32(&full[s]);
      mbar_arrive_expect_tx(bar, bytes);
      tma_load_3d(base, &a.maps[jb.xmap[seg]], (chunk & 1) * 64, jb.xcol[seg] >> 3, chunk >> 1, bar);
      tma_load_3d(base + 32768u, &a.maps[jb.ymap[seg]], (chunk & 1) * 64, jb.ycol[seg] >> 3, chunk >> 1, bar);
    }
  } else if (tid == 32) {
    // ================= MMA issuer =================
    for (int it = 0; it < nit && ok; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      ok = mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      const uint32_t base = s0 + (uint32_t)s * STAGE_BYTES;
      const uint32_t idesc = idesc_mn(128, (uint32_t)jb.n_mma);
      for (int mh = 0; mh < jb.m_tiles; ++mh)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16_ss(tmem_base + (uint32_t)(mh * 256), desc_mn(base + (uint32_t)(mh * 16384 + ks * 256), 128u, 1024u),
                      desc_mn(base + 32768u + (uint32_t)(ks * 256), 128u, 1024u), idesc, (it | ks) ? 1u : 0u);
      umma_commit(smem_u32(&empty[s]));
    }
    umma_commit(smem_u32(&acc_full));
  } else if (warp >= 2) {
    // ================= bias sums (during the main loop), then the flush =================
    const int t = tid - 64;                        // 0..127: features 2t, 2t+1 of the X tile
    float bs0 = 0.f, bs1 = 0.f;
    const bool want_b = jb.db_off >= 0;
    for (int it = 0; it < nit && ok; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      const int seg = it % jb.nseg;
      // always wait for the stage before releasing it: the `empty` barrier counts one arrival per warp and stage use,
      // a warp running ahead of the fill would arrive into the wrong barrier phase
      ok = mbar_wait(smem_u32(&full[s]), ph);
      if (ok && want_b && seg == 0 && 2 * t < jb.m_tiles * 128) {
        // X tile in shared memory: [column group][64 rows][16 bytes]; features 2t, 2t+1 = word (t & 3) of group t >> 2
        const uint32_t base = s0 + (uint32_t)s * STAGE_BYTES + (uint32_t)(t >> 2) * 1024u + (uint32_t)(t & 3) * 4u;
#pragma unroll 8
        for (uint32_t k = 0; k < 64; ++k) {
          uint32_t w;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + k * 16u));
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
          bs0 += f.x; bs1 += f.y;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    if (ok) ok = mbar_wait(smem_u32(&acc_full), 0);
    tc_fence_after();
    if (ok && nit > 0) {
      const float isig = a.sigma ? __ldg(a.sigma + 1) : 1.0f;
      const float scale = jb.scale * isig;
      if (want_b) {
        float* db = a.dpacked + jb.db_off;
        if (2 * t < jb.rows) atomicAdd(db + 2 * t, bs0 * jb.db_scale * isig);
        if (2 * t + 1 < jb.rows) atomicAdd(db + 2 * t + 1, bs1 * jb.db_scale * isig);
      }
      const int q = warp & 3;                      // TMEM lane quarter this warp may read
      for (int mh = 0; mh < jb.m_tiles; ++mh) {
        const int r = mh * 128 + q * 32 + lane;
        float* drow = a.dpacked + jb.dw_off + (long long)r * jb.dw_ld;
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * 256);
        for (int cc = 0; cc < jb.n_mma; cc += 32) {
          if (cc >= jb.cols) break;
          float v[32];
          tmem_ld32(tb + (uint32_t)cc, v);          // columns beyond n_mma inside the 32 are stale: masked by `cols` below
          tmem_ld_wait();
          if (r < jb.rows) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int c = cc + j;
              if (c + 3 < jb.cols) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c), "f"(v[j] * scale),
                             "f"(v[j + 1] * scale), "f"(v[j + 2] * scale), "f"(v[j + 3] * scale)
                             : "memory");
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (c + e < jb.cols) atomicAdd(drow + c + e, v[j + e] * scale);
              }
            }
          }
        }
      }
    }
  }
  if (!ok && fault) *fault = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Collects the operand tensors and jobs of one network, then launches.
struct Builder {
  Args a;
  int nmaps = 0, njobs = 0;
  long long N = 0;
  double cost[MAX_JOBS];
  bool bad = false;

  Builder(long long n, float* dpacked, const float* sigma) : N(n) {
    memset(&a, 0, sizeof(a));
    a.dpacked = dpacked;
    a.sigma = sigma;
  }
  // tile-blocked 16-bit tensor of width W (chain_engine.cuh), read with boxes of `box_cg` column groups x 64 rows
  int add_map(const void* ptr, int W, int box_cg) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || nmaps >= MAX_MAPS || ((uintptr_t)ptr & 127) || (W & 7) || box_cg < 1 || box_cg > 32) {
      ce::bad_value("wgrad tensor map arguments", nmaps);
      bad = true;
      return 0;
    }
    const cuuint64_t tiles = (cuuint64_t)((N + 127) / 128);
    const cuuint64_t dims[3] = {512, (cuuint64_t)(W / 8), tiles};
    const cuuint64_t strides[2] = {2048, (cuuint64_t)W * 256};
    const cuuint32_t box[3] = {256, (cuuint32_t)box_cg, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(&a.maps[nmaps], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ce::bad_value("cuTensorMapEncodeTiled", (int)r); bad = true; return 0; }
    return nmaps++;
  }
  // X operand of a job with `rows` output rows / Y operand with `cols` output columns
  int add_x(const void* ptr, int W, int rows) { return add_map(ptr, W, ((rows + 127) / 128) * 16); }
  int add_y(const void* ptr, int W, int cols) { return add_map(ptr, W, (((cols + 15) & ~15) + 7) / 8); }
  Job* add_job(int rows, int cols, long long dw_off, int dw_ld, float scale, long long db_off, float db_scale) {
    if (njobs >= MAX_JOBS || rows < 1 || rows > 256 || cols < 1 || cols > 256) {
      ce::bad_value("wgrad job shape", njobs);
      bad = true;
      return &a.jobs[0];
    }
    Job* j = &a.jobs[njobs++];
    memset(j, 0, sizeof(*j));
    j->rows = rows; j->cols = cols; j->m_tiles = (rows + 127) / 128; j->n_mma = (cols + 15) & ~15;
    j->dw_off = dw_off; j->dw_ld = dw_ld; j->scale = scale; j->db_off = db_off; j->db_scale = db_scale;
    return j;
  }
  static void add_seg(Job* j, int xmap, int xcol, int ymap, int ycol) {
    const int s = j->nseg++;
    if (s >= 2) return;
    j->xmap[s] = xmap; j->xcol[s] = xcol; j->ymap[s] = ymap; j->ycol[s] = ycol;
  }
  int launch(cudaStream_t st, int family) {
    if (bad) return ce::bad_value("wgrad builder", nmaps);
    if (njobs == 0 || N <= 0) return 0;
    const int chunks = (int)((N + 63) / 64);
    const int sms = ce::num_sms();
    double total = 0.0;
    for (int j = 0; j < njobs; ++j) {
      const Job& jb = a.jobs[j];
      if (jb.nseg < 1 || jb.nseg > 2) return ce::bad_value("wgrad job segments", j);
      cost[j] = (double)jb.nseg * (jb.m_tiles * 16 + (jb.n_mma + 7) / 8);
      total += cost[j];
    }
    // splits per job proportional to its bytes, at least one, at most one per 64-point chunk, sum <= number of SMs
    int splits[MAX_JOBS], sum = 0;
    for (int j = 0; j < njobs; ++j) {
      int s = (int)(cost[j] / total * (sms - njobs)) + 1;
      if (s > chunks) s = chunks;
      splits[j] = s;
      sum += s;
    }
    if (sum > MAX_UNITS || sum > sms) return ce::bad_value("wgrad units", sum);
    int nu = 0;
    double flops = 0.0, bytes = 0.0;
    for (int j = 0; j < njobs; ++j) {
      for (int s = 0; s < splits[j]; ++s) {
        Unit& u = a.units[nu++];
        u.job = j;
        u.c0 = (int)((long long)chunks * s / splits[j]);
        u.c1 = (int)((long long)chunks * (s + 1) / splits[j]);
      }
      const Job& jb = a.jobs[j];
      flops += 2.0 * (double)N * jb.nseg * jb.m_tiles * 128 * jb.n_mma;
      bytes += 2.0 * (double)N * jb.nseg * (jb.rows + jb.cols);
    }
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    prof_begin(family, st, flops, bytes);
    VDN_LAUNCH(wgrad16_kernel, nu, THREADS, SMEM, st, a, g_tc_fault);
    prof_end(family, st);
    return ce::debug_sync(st, "wgrad16_kernel", nu);
  }
};

}  // namespace wg
}  // namespace vdn

// Chain engine (tensor-core mode): ONE persistent tcgen05 kernel that runs a whole layer chain of an MLP pass for
// tiles of 128 points, the activations / cotangents never leaving the SM between layers.  The host describes a pass
// as a table of PHASES; a phase is one GEMM  D[128 x n] = A[128 x K] * B[n x K]^T  plus an element-wise epilogue
// that turns D into the next phase's A operand and / or into the few tensors that have to reach HBM.
//
//   passes built on it (sdf_net.cu, relu_nets.cu):
//     SDF training forward        fields.py:72-92      (softplus chain, stores the fp16 layer inputs)
//     SDF analytic normals        fields.py:97-108     (reverse-mode input gradient)
//     SDF backward phase 1 / 2    replaces autograd's double backward (Appendix A (1), (2))
//     RenderingNetwork / NeRF forward and input-gradient passes   fields.py:148-176, 324-355
//
// Execution model (the one of sdf_chain_tc.cuh, generalised): per CTA two tiles ("slots" X and Y) are in flight and
// alternate phases.  The A operand of a slot lives in TENSOR MEMORY as packed 16-bit pairs (columns 256 + 128 s ..,
// lane = point) and is the A operand of tcgen05.mma kind::f16 (fp16 A, fp16 weights, fp32 accumulate); one
// 256-column fp32 accumulator D is shared by both slots: sixteen epilogue warps stream it through registers group by
// group, release it at half time for the other slot's MMAs, and finish the element-wise work while the tensor pipe is
// busy with the other slot.  Weight K blocks ([n x 64] fp16 SWIZZLE_128B images, mlp_layout.cuh) stream through a
// 4-stage cp.async.bulk / mbarrier ring that holds exactly one phase; both slots use the same blocks.
//
// What reaches HBM is fp16 and tile-blocked ([tile][8-column group][128 rows][8], see blk_index): every phase may store
// the A operand it produces - that tensor is at the same time the saved activation of the backward passes and an
// operand of the grouped weight-gradient kernel (wgrad16.cuh), which reads it through TMA tensor maps.  Cotangents
// carry the launch's power-of-two loss scale (Args::sigma).
//
// Two element-wise code paths: the FAST path (fast_phase / fast_group: fully unrolled, resident in the instruction
// cache, no spills) for regular 64-column blocks, and the GENERAL path (run_group, looped) for everything else - fp32
// side outputs, rank-1 terms, scratch, the ragged block of a skip layer.  tools/diag_chain_tl.py (build with
// -DVDN_CHAIN_TL) prints the per-CTA timeline that guided their design.
//
// Thread mapping of the epilogue warps: warp w -> TMEM lane quarter q = w & 3 (rows 32 q .. 32 q + 31) and column
// block hq = w >> 2 (columns 64 hq .. 64 hq + 63), eight groups of eight columns.
#pragma once
#include "gemm_tc.cuh"
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

// Debug build (-DVDN_CHAIN_TL): CTA 0 records clock64() stamps of every (tile pair, phase, slot) into the buffer given to
// vdn_debug_timeline(): [0..3] MMA issuer (start, accumulator drained, A ready, MMAs issued), [4 + 2 w], [5 + 2 w] epilogue
// warp w (accumulator full, phase done), [40 + g] warp 12's general-path group g done; 48 values per (pair, phase, slot).
// tools/diag_chain_tl.py prints it.
#ifdef VDN_CHAIN_TL
#define VDN_TL(ptr, i) do { if (ptr) (ptr)[i] = clock64(); } while (0)
#else
#define VDN_TL(ptr, i) do { } while (0)
#endif

namespace vdn {
extern long long* g_tc_dbg;   // api.cu: optional device buffer for debug time stamps (vdn_debug_timeline)
namespace ce {

constexpr int EPI_WARPS = 16;
constexpr int THREADS = (EPI_WARPS + 2) * 32;   // + warp 16: TMEM alloc + MMA issue; warp 17: weight stream
// Four 32 KB weight stages hold exactly the K blocks of one phase (K <= 256): slot X streams them, slot Y reuses all of
// them, and the next phase's blocks arrive while the epilogues run.  Six stages prefetched further ahead but left only
// ~28 KB of L1 for the epilogue's operand lines; with four, L1 is ~92 KB and the step is 4% faster (measured; three and
// two stages are no better).
constexpr int WSTAGES = 4;
constexpr int KEEP = 4;                          // weight blocks of a phase kept in the ring for the second slot
constexpr uint32_t W_STAGE = 32768;
constexpr int MAX_PHASES = 16;
constexpr int MAX_STASH = 2;                     // stash slots per tile (fp16, 64 values per thread each)
constexpr uint32_t TAIL_BYTES = 128 * 128;        // skip-connection tail of a tile's 128 rows: up to 64 values of 16 bits each
constexpr size_t SMEM = WSTAGES * W_STAGE + 1024 + MAX_PHASES * 256 * sizeof(float) + 2 * 256 * sizeof(float) + TAIL_BYTES;

constexpr float kB2 = 144.26950408889634f;       // beta / ln 2 (Softplus(beta=100), fields.py:50)
constexpr float kInvB2 = 1.0f / 144.26950408889634f;

enum Op : int {
  OP_OUT32 = 0,   // fp32 output only: o32[m, c] = act(x) * o32_mul (act 0 none, 1 sigmoid, 2 relu)
  OP_SOFTPLUS,    // A = fp16 log2(1 + 2^x)                                     (SDF forward, base-2 softplus units)
  OP_NSTEP,       // A = fp16 S * x,  S = 1 - 2^-aux0                            (SDF normals pass: delta_{l-1})
  OP_P1STEP,      // A = S * x * a_mul ; o16b = 100 (1 - S) aux1 x               (backward of the normals pass)
  OP_P2STEP,      // A = S * x + aux1                                            (ordinary backward with injection)
  OP_RELU,        // A = fp16 max(x, 0)
  OP_LINEAR,      // A = fp16 x
  OP_STASH,       // scratch <- raw accumulator
  OP_MASK,        // A = (aux0 > 0 ? x : 0)   (aux0 null: x)
};
// x = acc * dsc + bias (+ stash) (+ r1[m] * row[c])

struct Phase {
  // ---- tensor-core side ----
  long long img_off;     // float offset (into `packed`) of K block 0 of the 16-bit image of B
  int img_rows;          // rows per K block of that image
  int row0;              // first image row (multiple of 8)
  int kb0;               // first K block
  int n_mma;             // N of the MMA (multiple of 16, <= 256)
  int nks;               // K steps of 16 (1..16)
  int a_col;             // packed-column offset of the A operand inside the slot
  // ---- epilogue ----
  int op;
  int width;             // valid output columns of this phase (<= n_mma)
  float dsc;             // accumulator scale
  long long bias_off;    // float offset of the bias in `packed`, -1: none
  float bias_mul;
  int a_out;             // 1: the epilogue writes the slot (A of the next phase)
  int a_wr;              // number of A columns written (multiple of 8, zero padded beyond width + tail)
  float a_mul;           // OP_P1STEP: scale of the A written
  const void* aux0; int ld0;
  const void* aux1; int ld1;
  void* o16a; int ldo16a; float o16a_mul;   // fp16 copy of the A values (times o16a_mul)
  void* o16b; int ldo16b;                   // second fp16 output (OP_P1STEP)
  float* o32; int ldo32; int o32_c0, o32_w; float o32_mul; int act;   // fp32 store of act(x) for columns [o32_c0, o32_c0 + o32_w)
  int o32_unscale;       // the fp32 output is a cotangent: multiply by 1 / sigma
  int o32_vec;           // set by launch(): full groups of that store may use 16-byte accesses
  int fast;              // set by launch(): number of leading 64-column blocks whose warps take the fast path (0: none)
  int edge;              // set by launch(): the block after them is the ragged edge of a skip layer (slow path)
  float* o32b; int ldo32b; int o32b_c0, o32b_w;                       // second fp32 store of x (no activation), other columns
  const void* tail; int ldt; int tail_w; float tail_mul;   // A columns [width, width + tail_w) from here
  const float* r1; int r1_stride; float r1_mul; int r1_row;   // rank-1 term, r1_row in {0, 1}
  int r1_scaled;         // the rank-1 coefficient is an unscaled fp32 cotangent: multiply by sigma
  int stash_w, stash_r;  // stash index written (OP_STASH) / read (-1: none)
  const void* aload; int al_ld, al_w;       // after this phase: load [128 x al_w] 16-bit values into the slot (A of the next)
};

struct Args {
  int P;
  long long N;
  const float* packed;
  const void* a0; int a0_ld, a0_w;          // A operand of phase 0
  long long row_off[2]; int row_len[2];     // rank-1 row vectors (float offsets into packed, -1: none)
  uint16_t* stash;                          // fp16 scratch [grid][2][MAX_STASH][8][512][8] (stays in L2)
  int pf1_next;                             // L1 prefetch of the next tile's first loads during the last phase
  int pf_mode;                              // software L2 prefetch one phase ahead: 0 none (default), 1 per line, 2 bulk (TMA)
  const float* sigma;                       // backward passes: device {sigma, 1 / sigma}, the power-of-two loss scale the
                                            // fp16 cotangents of this launch carry (null: 1)
  Phase ph[MAX_PHASES];
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// saturating variant for cotangents: a value beyond the fp16 range becomes +-65504 instead of infinity
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void unpack_h8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// a' = log2(1 + 2^t), same evaluation as sdf_chain_tc.cuh
__device__ __forceinline__ float softplus2(float t) {
  const float w = ex2f(-fabsf(t));
  float q = fmaf(0.04008112847805023f, w, -0.1803952157497406f);
  q = fmaf(q, w, 0.4036492109298706f);
  q = fmaf(q, w, -0.7047332525253296f);
  q = fmaf(q, w, 1.4414016008377075f);
  return fmaf(w, q, fmaxf(t, 0.0f));
}
// ---- 16-bit tensors of the chains: TILE-BLOCKED layout ---------------------------------------------------------
// A [Npad, W] tensor (W a multiple of 8) is stored as [tile = m / 128][column group = c / 8][row = m % 128][8 elements]:
// the 32 lanes of an epilogue warp (32 consecutive rows, the same eight columns) read or write 512 contiguous bytes with
// one 16-byte access each - a row-major layout would make every such access touch 32 different lines.  The grouped
// weight-gradient kernel reads the same tensors through 4-D TMA tensor maps (wgrad16.cuh).
__host__ __device__ __forceinline__ long long blk_index(long long m, int W, int c) {
  return (m >> 7) * ((long long)W * 128) + (long long)(c >> 3) * 1024 + (m & 127) * 8 + (c & 7);
}
// inverse: linear position in the blocked tensor -> (row m, column c); lets pointwise kernels write it coalesced
__host__ __device__ __forceinline__ void blk_decode(long long idx, int W, long long* m, int* c) {
  const long long per_tile = (long long)W * 128;
  const long long tile = idx / per_tile;
  const int rem = (int)(idx - tile * per_tile);
  *m = tile * 128 + ((rem & 1023) >> 3);
  *c = (rem >> 10) * 8 + (rem & 7);
}
__device__ __forceinline__ uint4 ld16(const void* base, long long m, int W, int c) {      // c multiple of 8
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + blk_index(m, W, c)));
}
__device__ __forceinline__ void st16(void* base, long long m, int W, int c, const uint32_t (&p)[4]) {
  *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + blk_index(m, W, c)) = make_uint4(p[0], p[1], p[2], p[3]);
}
// the 128-byte lines of columns [c0, c0 + 64) of rows m .. m + 31 (a warp's rows) -> L2; issued by lanes 0, 8, 16, 24
__device__ __forceinline__ void prefetch16(const void* base, long long m, int W, int c0, int wmax, int lane) {
  if (lane & 7) return;
#pragma unroll
  for (int g = 0; g < 8; ++g)
    if (c0 + 8 * g < wmax)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint16_t*>(base) + blk_index(m, W, c0 + 8 * g)));
}

// ---- cold paths, kept out of line so that the per-phase hot loop stays small (the kernel interprets nine epilogue
// kinds; with everything inlined its hot paths did not fit the instruction cache) ---------------------------------
__device__ __forceinline__ float apply_act(float t, int act) {
  if (act == 1) return __frcp_rn(1.0f + ex2f(-1.4426950408889634f * t));      // sigmoid
  if (act == 2) return fmaxf(t, 0.0f);
  return t;
}
// fp32 side output of eight values, element-wise (ragged ranges, unaligned or strided destinations)
static __device__ __noinline__ void o32_scalar(float* dst, int ld, int c_first, int w, int act, float mul, long long m, int cg,
                                               float x0, float x1, float x2, float x3, float x4, float x5, float x6, float x7) {
  const float x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
  float* op = dst + m * ld;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int c = cg + j - c_first;
    if (c >= 0 && c < w) op[c] = apply_act(x[j], act) * mul;
  }
}
// columns >= width of an A tile: skip-connection tail or zero padding; r is patched in place.  The tail values of the
// thread's row were copied into shared memory with cp.async at the start of the phase (phase_body), so the element-wise
// loop never waits for them.  Inlined on purpose, with everything in registers.  Earlier versions made the ONE ragged
// layer of the SDF net cost as much as five regular ones (per-CTA timeline, tools/diag_chain_tl.py): an out-of-line
// function taking `const Phase&` turned every field access into a generic load from kernel-parameter space (~1 us each),
// passing `r` by pointer put the caller's arrays into local memory, and loading the tail values where they are used cost
// one DRAM round trip (1.6 us) per group of eight columns.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ragged_tail(bool has_tail, uint32_t stail, int tail_w, float tail_mul, int width, int cg,
                                            float (&r)[8], float (&r2)[8], bool has_r2) {
  const int t0 = cg - width;                        // tail column of element 0 (negative: the group straddles the edge)
  const int tw = has_tail ? tail_w : 0;
  const int g0 = (t0 > 0 ? t0 : 0) >> 3;
  uint4 u0 = make_uint4(0u, 0u, 0u, 0u), u1 = make_uint4(0u, 0u, 0u, 0u);
  if (tw > 0) cp_async_wait_all();
  if (g0 * 8 < tw) u0 = lds128(stail + (uint32_t)g0 * 16u);
  if (g0 * 8 + 8 < tw) u1 = lds128(stail + (uint32_t)g0 * 16u + 16u);
  const int sh = t0 - g0 * 8;                       // element j lives at position sh + j of the 16 loaded ones (-7 .. 7)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = sh + j, t = t0 + j;
    // 16-bit element e of {u0, u1}: word e >> 1, half e & 1 - selected with compile-time word indices
    uint32_t word = 0;
    word = (e >> 1) == 0 ? u0.x : word; word = (e >> 1) == 1 ? u0.y : word; word = (e >> 1) == 2 ? u0.z : word;
    word = (e >> 1) == 3 ? u0.w : word; word = (e >> 1) == 4 ? u1.x : word; word = (e >> 1) == 5 ? u1.y : word;
    word = (e >> 1) == 6 ? u1.z : word; word = (e >> 1) == 7 ? u1.w : word;
    const uint16_t h = (e & 1) ? (uint16_t)(word >> 16) : (uint16_t)(word & 0xffffu);
    const float v = (t >= 0 && t < tw) ? __half2float(__ushort_as_half(h)) * tail_mul : 0.0f;
    if (t >= 0) {
      r[j] = v;
      if (has_r2) r2[j] = 0.0f;
    }
  }
}

// Every 16-bit tensor of the chains is fp16: the A operands in tensor memory, the weight images, and everything that
// goes to or comes from HBM (aux0, aux1, o16a, o16b) - the weight-gradient kernel needs one format for both operands.
// Forward quantities fit fp16's range as they are (a' = kB2 softplus(z), ReLU activations).  Cotangents (OP_P1STEP,
// OP_P2STEP, OP_MASK) do not - d loss / d z is ~1e-6 - so a backward launch carries ONE power-of-two loss scale sigma
// (Args::sigma, computed on the device from the largest incoming cotangent): every cotangent tensor of the launch is
// sigma times the true one, fp32 outputs are multiplied by 1 / sigma on the way out, and conversions saturate instead
// of overflowing.  Eleven significant bits instead of bf16's eight, and one MMA pass instead of a hi/lo weight pair.
template <int OP> struct OpTraits {
  static constexpr bool bwd = (OP == OP_P1STEP || OP == OP_P2STEP || OP == OP_MASK);       // A is a scaled cotangent
  static constexpr bool aux0 = (OP == OP_NSTEP || OP == OP_P1STEP || OP == OP_P2STEP || OP == OP_MASK);
  static constexpr bool aux1 = (OP == OP_P1STEP || OP == OP_P2STEP);
};
template <int OP> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return OpTraits<OP>::bwd ? pack_h2_sat(lo, hi) : pack_h2(lo, hi);
}

// element offset of (row m, column c0 + 32 half) in a tile-blocked tensor of width W; group gi adds gi * 1024
__device__ __forceinline__ long long blk_base(long long m, int W, int c0, int half) {
  return (m >> 7) * ((long long)W * 128) + (long long)((c0 >> 3) + 4 * half) * 1024 + (m & 127) * 8;
}

// One group of eight columns (cg = c0 + 8 g .. + 7) of one phase for one thread (row m): the general element-wise path.
// The caller loops over the groups WITHOUT unrolling, so that this body exists once per epilogue kind: the general path
// runs rarely (first / last layers, the ragged skip layer), its instructions are never resident in the instruction cache,
// and the fully unrolled version spent most of its time fetching them (per-CTA timeline, tools/diag_chain_tl.py: 2 us per
// group against 0.3 us in the resident fast path).
template <int OP>
__device__ __forceinline__ void run_group(const Phase& ph, const float (&v)[8], int g, const uint4& q0, const uint4& q1,
                                          const float* sb, const float* srow, long long m, long long N, int c0, uint32_t tA,
                                          uint16_t* stash_base, float sig, float isig, uint32_t stail) {
  using T = OpTraits<OP>;
  const int width = ph.width;
  const float dsc = ph.dsc;
  const bool rowok = m < N;
  const bool has_r1 = ph.r1 != nullptr;
  const float r1v = (has_r1 && rowok) ? ph.r1[m * ph.r1_stride] * ph.r1_mul * (ph.r1_scaled ? sig : 1.0f) : 0.0f;   // padded rows stay exactly zero
  const float* rr = srow + ph.r1_row * 256;
  const int cg = c0 + 8 * g;
  // destinations of this group (tile-blocked 16-bit copies; fp32 row-major side output)
  uint4* pa = ph.o16a ? reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ph.o16a) + blk_base(m, ph.ldo16a, c0, 0)) + g * 128 : nullptr;
  uint4* pb = (OP == OP_P1STEP && ph.o16b)
                  ? reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ph.o16b) + blk_base(m, ph.ldo16b, c0, 0)) + g * 128 : nullptr;
  const bool o32_on = ph.o32 && rowok;
  const bool o32b_on = ph.o32b && rowok;
  float4* pv = (o32_on && ph.o32_vec) ? reinterpret_cast<float4*>(ph.o32 + m * ph.ldo32 + (cg - ph.o32_c0)) : nullptr;
  const float om = ph.o16a_mul;
  const float o32_mul = ph.o32_unscale ? ph.o32_mul * isig : ph.o32_mul;
  float x[8];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(sb + cg), b1 = *reinterpret_cast<const float4*>(sb + cg + 4);
    x[0] = fmaf(v[0], dsc, b0.x); x[1] = fmaf(v[1], dsc, b0.y); x[2] = fmaf(v[2], dsc, b0.z);
    x[3] = fmaf(v[3], dsc, b0.w); x[4] = fmaf(v[4], dsc, b1.x); x[5] = fmaf(v[5], dsc, b1.y);
    x[6] = fmaf(v[6], dsc, b1.z); x[7] = fmaf(v[7], dsc, b1.w);
  }
  if (has_r1) {
    const float4 w0 = *reinterpret_cast<const float4*>(rr + cg), w1 = *reinterpret_cast<const float4*>(rr + cg + 4);
    x[0] = fmaf(r1v, w0.x, x[0]); x[1] = fmaf(r1v, w0.y, x[1]); x[2] = fmaf(r1v, w0.z, x[2]); x[3] = fmaf(r1v, w0.w, x[3]);
    x[4] = fmaf(r1v, w1.x, x[4]); x[5] = fmaf(r1v, w1.y, x[5]); x[6] = fmaf(r1v, w1.z, x[6]); x[7] = fmaf(r1v, w1.w, x[7]);
  }
  if (OP != OP_STASH && ph.stash_r >= 0) {      // written earlier by this very thread: plain (coherent) load
    const uint4 u = *reinterpret_cast<const uint4*>(stash_base + ((size_t)ph.stash_r * 8 + g) * (512 * 8));
    float s8[8];
    unpack_h8(u, s8);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += s8[j];
  }
  if (OP == OP_STASH) {
    *reinterpret_cast<uint4*>(stash_base + ((size_t)ph.stash_w * 8 + g) * (512 * 8)) =
        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]),
                   pack_h2(v[6], v[7]));
    return;
  }
  // fp32 side outputs of x (final outputs, skip-connection tails)
  if (o32_on && cg + 8 > ph.o32_c0 && cg < ph.o32_c0 + ph.o32_w) {
    if (pv && cg >= ph.o32_c0 && cg + 8 <= ph.o32_c0 + ph.o32_w) {      // o32_vec: no activation, unit scale handled below
      const float mul = o32_mul;
      pv[0] = make_float4(x[0] * mul, x[1] * mul, x[2] * mul, x[3] * mul);
      pv[1] = make_float4(x[4] * mul, x[5] * mul, x[6] * mul, x[7] * mul);
    } else {
      o32_scalar(ph.o32, ph.ldo32, ph.o32_c0, ph.o32_w, ph.act, o32_mul, m, cg, x[0], x[1], x[2], x[3], x[4], x[5], x[6],
                 x[7]);
    }
  }
  if (o32b_on && cg + 8 > ph.o32b_c0 && cg < ph.o32b_c0 + ph.o32b_w)
    o32_scalar(ph.o32b, ph.ldo32b, ph.o32b_c0, ph.o32b_w, 0, 1.0f, m, cg, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
  if (OP == OP_OUT32) return;
  // ---- the A value (and the optional second output) ----
  float r[8], r2[8];
  const bool in = cg < width;
  if (OP == OP_SOFTPLUS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = softplus2(x[j]);
  } else if (OP == OP_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = fmaxf(x[j], 0.0f);
  } else if (OP == OP_LINEAR) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
  } else if (OP == OP_MASK) {
    if (ph.aux0 && in) {
      float h8[8];
      unpack_h8(q0, h8);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = h8[j] > 0.0f ? x[j] : 0.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = x[j];
    }
  } else {   // OP_NSTEP / OP_P1STEP / OP_P2STEP: softplus'(z) = 1 - 2^-a' from the saved activation
    float s_[8];
    if (in) {
      float a8[8];
      unpack_h8(q0, a8);
#pragma unroll
      for (int j = 0; j < 8; ++j) s_[j] = 1.0f - ex2f(-a8[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_[j] = 0.0f;
    }
    if (OP == OP_NSTEP) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = s_[j] * x[j];
    } else if (OP == OP_P1STEP) {
      float d8[8];
      if (in) {
        unpack_h8(q1, d8);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d8[j] = 0.0f;
      }
      const float am = ph.a_mul;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j] = s_[j] * x[j] * am;
        r2[j] = 100.0f * (1.0f - s_[j]) * d8[j] * x[j];
      }
    } else {
      float z8[8];
      if (ph.aux1 && in) {
        unpack_h8(q1, z8);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z8[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(s_[j], x[j], z8[j]);
    }
  }
  // ragged edge: columns >= width carry the skip tail or zero padding
  if (cg + 8 > width) ragged_tail(ph.tail != nullptr, stail, ph.tail_w, ph.tail_mul, width, cg, r, r2, OP == OP_P1STEP);
  uint32_t p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = pack2<OP>(r[2 * i], r[2 * i + 1]);
  if (ph.a_out) tc::tmem_st4(tA + (uint32_t)(4 * g), p);
  if (pa) {
    if (OP == OP_P2STEP) {               // stored pre-scaled (z-bar * dsc / kB2: the weight gradient's operand)
      pa[0] = make_uint4(pack2<OP>(r[0] * om, r[1] * om), pack2<OP>(r[2] * om, r[3] * om),
                                pack2<OP>(r[4] * om, r[5] * om), pack2<OP>(r[6] * om, r[7] * om));
    } else {
      pa[0] = make_uint4(p[0], p[1], p[2], p[3]);
    }
  }
  if (OP == OP_P1STEP && pb)
    pb[0] = make_uint4(pack2<OP>(r2[0], r2[1]), pack2<OP>(r2[2], r2[3]), pack2<OP>(r2[4], r2[5]), pack2<OP>(r2[6], r2[7]));
}

// Fast path for REGULAR phases (Phase::fast, set by launch()): every column of an active warp is a valid output (width a
// multiple of 64, nothing written beyond it), no fp32 side output, no rank-1 term, no scratch, no skip tail - most
// layers.  The 64 columns of a thread are streamed group by group: eight accumulator values and the 16-byte auxiliary
// operands of group g + 1 are fetched while group g is computed, so ~80 registers are live and nothing spills (the
// first version kept 32 + 32 prefetched registers per half and lost a quarter of its issue slots to local-memory
// traffic).  D is released when the last group has left tensor memory.
template <int OP>
__device__ __forceinline__ void fast_group(const Phase& ph, const float (&v)[8], const uint4& q0, const uint4& q1, const float* sb,
                                           int g, uint32_t tA, uint4* pa, uint4* pb, float dsc, float om, float am, bool a_out,
                                           bool has0, bool has1) {
  using T = OpTraits<OP>;
  float x[8];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(sb + 8 * g), b1 = *reinterpret_cast<const float4*>(sb + 8 * g + 4);
    x[0] = fmaf(v[0], dsc, b0.x); x[1] = fmaf(v[1], dsc, b0.y); x[2] = fmaf(v[2], dsc, b0.z); x[3] = fmaf(v[3], dsc, b0.w);
    x[4] = fmaf(v[4], dsc, b1.x); x[5] = fmaf(v[5], dsc, b1.y); x[6] = fmaf(v[6], dsc, b1.z); x[7] = fmaf(v[7], dsc, b1.w);
  }
  float r[8], r2[8];
  if (OP == OP_SOFTPLUS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = softplus2(x[j]);
  } else if (OP == OP_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = fmaxf(x[j], 0.0f);
  } else if (OP == OP_LINEAR) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = x[j];
  } else if (OP == OP_MASK) {
    if (has0) {
      float h8[8];
      unpack_h8(q0, h8);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = h8[j] > 0.0f ? x[j] : 0.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = x[j];
    }
  } else {
    float a8[8], s_[8];
    unpack_h8(q0, a8);
#pragma unroll
    for (int j = 0; j < 8; ++j) s_[j] = 1.0f - ex2f(-a8[j]);          // softplus'(z) = 1 - 2^-a'
    if (OP == OP_NSTEP) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = s_[j] * x[j];
    } else if (OP == OP_P1STEP) {
      float d8[8];
      unpack_h8(q1, d8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sx = s_[j] * x[j];
        r[j] = sx * am;
        r2[j] = 100.0f * d8[j] * (x[j] - sx);          // 100 (1 - S) delta x
      }
    } else {
      if (has1) {
        float z8[8];
        unpack_h8(q1, z8);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(s_[j], x[j], z8[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = s_[j] * x[j];
      }
    }
  }
  uint32_t p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = pack2<OP>(r[2 * i], r[2 * i + 1]);
  if (a_out) tc::tmem_st4(tA + (uint32_t)(4 * g), p);
  if (pa) {
    if (OP == OP_P2STEP)          // phase 2 stores z-bar pre-scaled for the weight gradient
      pa[g * 128] = make_uint4(pack2<OP>(r[0] * om, r[1] * om), pack2<OP>(r[2] * om, r[3] * om), pack2<OP>(r[4] * om, r[5] * om),
                               pack2<OP>(r[6] * om, r[7] * om));
    else
      pa[g * 128] = make_uint4(p[0], p[1], p[2], p[3]);
  }
  if (OP == OP_P1STEP && pb)
    pb[g * 128] = make_uint4(pack2<OP>(r2[0], r2[1]), pack2<OP>(r2[2], r2[3]), pack2<OP>(r2[4], r2[5]), pack2<OP>(r2[6], r2[7]));
}

template <int OP>
__device__ __forceinline__ void fast_phase(const Phase& ph, bool last, uint32_t tD, uint32_t tA, int c0, long long m,
                                           const float* sb, uint32_t bar_d_full, uint32_t par, uint32_t bar_d_drained,
                                           uint32_t bar_a_ready, bool* ok, long long* dw) {
  using namespace tc;
  using T = OpTraits<OP>;
  const bool has0 = T::aux0 && ph.aux0 != nullptr, has1 = T::aux1 && ph.aux1 != nullptr;
  const uint4* p0 = has0 ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux0) + blk_base(m, ph.ld0, c0, 0)) : nullptr;
  const uint4* p1 = has1 ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux1) + blk_base(m, ph.ld1, c0, 0)) : nullptr;
  uint4* pa = ph.o16a ? reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ph.o16a) + blk_base(m, ph.ldo16a, c0, 0)) : nullptr;
  uint4* pb = (OP == OP_P1STEP && ph.o16b)
                  ? reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ph.o16b) + blk_base(m, ph.ldo16b, c0, 0)) : nullptr;
  const float dsc = ph.dsc, om = ph.o16a_mul, am = ph.a_mul;
  const bool a_out = ph.a_out != 0;
  uint4 qa[2], qb[2];
  qa[0] = qa[1] = qb[0] = qb[1] = make_uint4(0u, 0u, 0u, 0u);
  if (has0) qa[0] = __ldg(p0);
  if (has1) qb[0] = __ldg(p1);
  // the lines of the next groups -> L1 (they sit in L2 since the previous phase's prefetch): an L2 hit costs ~0.4 us,
  // more than a group's worth of work, and registers for a deeper look-ahead do not exist
#pragma unroll
  for (int g = 1; g < 4; ++g) {
    if (has0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + g * 128));
    if (has1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + g * 128));
  }
  *ok = mbar_wait_relaxed(bar_d_full, par);
  VDN_TL(dw, 0);
  tc_fence_after();
  // groups 0..3 stream through two 8-value buffers; when group 4 arrives the remaining three are fetched at once and the
  // accumulator is released at half time, so the other tile's MMAs overlap the second half of this epilogue
  float v[2][8], w[3][8];
  tmem_ld8(tD, v[0]);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g <= 4) tmem_ld_wait();                       // group g has arrived (groups 5..7 arrive together with 4)
    if (g < 4) {
      tmem_ld8(tD + (uint32_t)(8 * (g + 1)), v[(g + 1) & 1]);
    } else if (g == 4) {
      tmem_ld8(tD + 40u, w[0]);
      tmem_ld8(tD + 48u, w[1]);
      tmem_ld8(tD + 56u, w[2]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_d_drained);                     // the accumulator has left tensor memory
      if (!last && !ph.a_out && !ph.aload) mbar_arrive(bar_a_ready);
    }
    if (g < 7) {
      if (has0) qa[(g + 1) & 1] = __ldg(p0 + (g + 1) * 128);
      if (has1) qb[(g + 1) & 1] = __ldg(p1 + (g + 1) * 128);
      if (g + 4 < 8) {
        if (has0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + (g + 4) * 128));
        if (has1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + (g + 4) * 128));
      }
    }
    if (g <= 4)
      fast_group<OP>(ph, v[g & 1], qa[g & 1], qb[g & 1], sb + c0, g, tA, pa, pb, dsc, om, am, a_out, has0, has1);
    else
      fast_group<OP>(ph, w[g - 5], qa[g & 1], qb[g & 1], sb + c0, g, tA, pa, pb, dsc, om, am, a_out, has0, has1);
  }
}

// One (phase, slot) of the epilogue role for one epilogue kind.
template <int OP>
__device__ __forceinline__ bool phase_body(const Phase& ph, bool last, int s, uint32_t tD, uint32_t tA, int c0, long long m,
                                           long long N, const float* sb, const float* srow, uint16_t* stb, uint32_t bar_d_full,
                                           uint32_t par, uint32_t bar_d_drained, uint32_t bar_a_ready, float sig, float isig, long long* dw, uint32_t stail) {
  using namespace tc;
  if (OP != OP_OUT32 && OP != OP_STASH && ph.fast) {
    bool ok = true;
    if (ph.edge) {
      // Skip layer: three regular 64-column blocks and the ragged one.  The ragged block's eight groups go through the
      // general element-wise code, which is a long dependent chain per group - left to the block's own four warps (one
      // per scheduler, nothing to interleave with) it took 4x a regular block.  So it is SHARED: every warp of a row
      // quarter may touch all of the quarter's tensor-memory lanes, and each of the four takes two groups, before its
      // own regular block.
      const int hq = c0 >> 6, ce = ph.fast * 64;
      const uint32_t tDe = tD - (uint32_t)c0 + (uint32_t)ce, tAe = tA - (uint32_t)(hq * 32) + (uint32_t)(ph.fast * 32);
      const int g0 = 2 * ((hq + 1) & 3);
      const bool has0 = OpTraits<OP>::aux0 && ph.aux0 != nullptr, has1 = OpTraits<OP>::aux1 && ph.aux1 != nullptr;
      uint4 q0[2], q1[2];
      q0[0] = q0[1] = q1[0] = q1[1] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (ce + 8 * (g0 + k) < ph.width) {
          if (has0) q0[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux0) + blk_base(m, ph.ld0, ce, 0)) + (g0 + k) * 128);
          if (has1) q1[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux1) + blk_base(m, ph.ld1, ce, 0)) + (g0 + k) * 128);
        }
      }
      if (ph.tail) {
        for (int k = 0; k * 8 < ph.tail_w && k < 8; ++k)
          cp_async16(stail + (uint32_t)k * 16u, reinterpret_cast<const uint16_t*>(ph.tail) + blk_index(m, ph.ldt, 8 * k));
        cp_async_commit();
      }
      ok = mbar_wait_relaxed(bar_d_full, par);
      tc_fence_after();
      float v[2][8];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (ce + 8 * (g0 + k) < ph.n_mma) {
          tmem_ld8(tDe + (uint32_t)(8 * (g0 + k)), v[k]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = 0.0f;
        }
      }
      tmem_ld_wait();
#pragma unroll 1
      for (int k = 0; k < 2; ++k)
        run_group<OP>(ph, k ? v[1] : v[0], g0 + k, k ? q0[1] : q0[0], k ? q1[1] : q1[0], sb, srow, m, N, ce, tAe, stb, sig, isig, stail);
    }
    if ((c0 >> 6) < ph.fast) {
      bool ok2 = true;
      fast_phase<OP>(ph, last, tD, tA, c0, m, sb, bar_d_full, par, bar_d_drained, bar_a_ready, &ok2, dw);
      return ok && ok2;
    }
    // this warp's own columns are not regular outputs of the phase: keep the barrier protocol in step
    if (!ph.edge) ok = mbar_wait_relaxed(bar_d_full, par);
    VDN_TL(dw, 0);
    tc_fence_before();
    mbar_arrive(bar_d_drained);
    if (!last && !ph.a_out && !ph.aload) mbar_arrive(bar_a_ready);
    return ok;
  }
  if (ph.tail && c0 + 64 > ph.width) {      // this warp's columns include tail columns: fetch the row's tail values now
    for (int k = 0; k * 8 < ph.tail_w && k < 8; ++k)
      cp_async16(stail + (uint32_t)k * 16u, reinterpret_cast<const uint16_t*>(ph.tail) + blk_index(m, ph.ldt, 8 * k));
    cp_async_commit();
  }
  // general path: one group of eight columns at a time (run_group), the next group's auxiliary operands in flight
  const int ncols = ph.a_out ? ph.a_wr : ph.width;      // columns this phase touches
  const bool has0 = OpTraits<OP>::aux0 && ph.aux0 != nullptr, has1 = OpTraits<OP>::aux1 && ph.aux1 != nullptr;
  const uint4* p0 = has0 ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux0) + blk_base(m, ph.ld0, c0, 0)) : nullptr;
  const uint4* p1 = has1 ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ph.aux1) + blk_base(m, ph.ld1, c0, 0)) : nullptr;
  uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0, q0n = q0, q1n = q0;
  if (has0 && c0 < ph.width) q0n = __ldg(p0);
  if (has1 && c0 < ph.width) q1n = __ldg(p1);
  const bool ok = mbar_wait_relaxed(bar_d_full, par);
  VDN_TL(dw, 0);
  tc_fence_after();
  int ng = (ncols - c0 + 7) >> 3;                        // groups of this warp that the phase touches (0 .. 8)
  ng = ng < 0 ? 0 : (ng > 8 ? 8 : ng);
  bool released = false;
  // the accumulator values of group g + 1 are requested before group g is processed (a tensor-memory load issued right
  // after the group's tensor-memory store would wait for it)
  float v[8], vn[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) vn[j] = 0.0f;
  if (ng > 0 && c0 < ph.n_mma) tmem_ld8(tD, vn);
#pragma unroll 1
  for (int g = 0; g < ng; ++g) {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = vn[j];
    if (g + 1 < ng) {
      if (c0 + 8 * (g + 1) < ph.n_mma) {
        tmem_ld8(tD + (uint32_t)(8 * (g + 1)), vn);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vn[j] = 0.0f;
      }
    } else {                    // the accumulator has left tensor memory
      tc_fence_before();
      mbar_arrive(bar_d_drained);
      // the slot's next A operand is what it holds already: release the MMA issuer right away
      if (!last && !ph.a_out && !ph.aload) mbar_arrive(bar_a_ready);
      released = true;
    }
    q0 = q0n; q1 = q1n;
    if (g + 1 < ng && c0 + 8 * (g + 1) < ph.width) {
      if (has0) q0n = __ldg(p0 + (g + 1) * 128);
      if (has1) q1n = __ldg(p1 + (g + 1) * 128);
    }
    run_group<OP>(ph, v, g, q0, q1, sb, srow, m, N, c0, tA, stb, sig, isig, stail);
#ifdef VDN_CHAIN_TL
    if (dw && (threadIdx.x >> 5) == 12) (dw - 28)[40 + g] = clock64();      // warp 12: group g done
#endif
  }
  if (!released) {              // none of this warp's columns belongs to the phase
    mbar_arrive(bar_d_drained);
    if (!last && !ph.a_out && !ph.aload) mbar_arrive(bar_a_ready);
  }
  return ok;
}

static __global__ void __launch_bounds__(THREADS, 1)
chain_kernel(const __grid_constant__ Args a, int* __restrict__ fault, long long* __restrict__ dbg) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_full[WSTAGES], w_empty[WSTAGES], a_ready[2], d_full, d_drained;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sW = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* sB = reinterpret_cast<float*>(smem_raw + (sW - smem_u32(smem_raw)) + WSTAGES * W_STAGE);
  float* sRow = sB + MAX_PHASES * 256;
  const uint32_t sTail = smem_u32(sRow + 512);      // [128 rows][128 bytes]
  const long long ntiles = (a.N + 127) / 128;
  const long long G = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(smem_u32(&w_full[s]), 1); mbar_init(smem_u32(&w_empty[s]), 1); }
    for (int j = 0; j < 2; ++j) mbar_init(smem_u32(&a_ready[j]), EPI_WARPS * 32);
    mbar_init(smem_u32(&d_full), 1);
    mbar_init(smem_u32(&d_drained), EPI_WARPS * 32);
    mbar_fence_init();
  }
  for (int i = tid; i < a.P * 256; i += THREADS) {      // biases (scaled), zero beyond the phase's width
    const int p = i >> 8, n = i & 255;
    const Phase& ph = a.ph[p];
    sB[i] = (ph.bias_off >= 0 && n < ph.width) ? a.packed[ph.bias_off + n] * ph.bias_mul : 0.0f;
  }
  for (int i = tid; i < 512; i += THREADS) {             // rank-1 row vectors
    const int r = i >> 8, n = i & 255;
    sRow[i] = (a.row_off[r] >= 0 && n < a.row_len[r]) ? a.packed[a.row_off[r] + n] : 0.0f;
  }
  if (warp == EPI_WARPS) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // accumulator D: columns [0,256); 16-bit A tiles: slot X columns [256,384), slot Y columns [384,512)
  const uint32_t tmem_base = tmem_base_s;
  bool ok = true;

  if (warp < EPI_WARPS) {
    // ================= epilogue warps =================
    const int q = warp & 3, hq = warp >> 2;
    const int row = q * 32 + lane;
    const int c0 = hq * 64;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tD = lane_base + (uint32_t)c0;
    uint32_t dcnt = 0;
    const uint32_t stail = sTail + (uint32_t)row * 128u;
    const float sig = a.sigma ? __ldg(a.sigma) : 1.0f, isig = a.sigma ? __ldg(a.sigma + 1) : 1.0f;
    // [128 x w] 16-bit values, row-major in HBM -> slot s (packed columns 0 ..), then signal the slot
    auto aload = [&](int s, long long tile, const void* src, int ld, int w) {
      const long long m = tile * 128 + row;               // < Npad: the buffers are padded to whole tiles
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (c0 + 8 * g < w) {
          const uint4 u = ld16(src, m, ld, c0 + 8 * g);
          const uint32_t r[4] = {u.x, u.y, u.z, u.w};
          tmem_st4(lane_base + 256u + (uint32_t)(s * 128 + hq * 32 + 4 * g), r);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&a_ready[s]));
    };
    // the lines of the tensors a phase reads (this warp's rows and columns) -> L2, one phase ahead
    auto prefetch = [&](const Phase& ph, long long m) {
      if (a.pf_mode == 2) {       // a tile of a blocked tensor is contiguous: one bulk prefetch per tensor and tile
        if (warp == 0 && lane == 0) {
          const long long tile = m >> 7;
          if (ph.aux0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint16_t*>(ph.aux0) + tile * ph.ld0 * 128),
                         "r"((ph.width < ph.ld0 ? ph.width : ph.ld0) * 256));
          if (ph.aux1)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint16_t*>(ph.aux1) + tile * ph.ld1 * 128),
                         "r"((ph.width < ph.ld1 ? ph.width : ph.ld1) * 256));
        }
        return;
      }
      if (ph.aux0) prefetch16(ph.aux0, m, ph.ld0, c0, ph.width, lane);
      if (ph.aux1) prefetch16(ph.aux1, m, ph.ld1, c0, ph.width, lane);
    };
    if ((long long)blockIdx.x < ntiles) aload(0, blockIdx.x, a.a0, a.a0_ld, a.a0_w);
    if (blockIdx.x + G < ntiles) aload(1, blockIdx.x + G, a.a0, a.a0_ld, a.a0_w);
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int p = 0; p < a.P && ok; ++p) {
        const Phase& ph = a.ph[p];
        const bool last = (p == a.P - 1);
        const float* sb = sB + p * 256;
        for (int s = 0; s < 2 && ok; ++s) {
          if (s && !hasY) break;
          const long long tile = tX + s * G;
          const long long m = tile * 128 + row;
          // No software L2 prefetch by default: prefetch.global.L2 and cp.async.bulk.prefetch.L2 of the next phase's
          // operands both DOUBLED the DRAM reads of these kernels on B200 (ncu dram__bytes_read 939 MB vs 567 MB for the
          // same launch) and slowed them down; the lines are fetched into L1 a few steps ahead instead (fast_phase).
          // VDN_PF2=1 / 2 re-enable the two variants for measurement.
          if (a.pf_mode) {
            if (!last) {
              prefetch(a.ph[p + 1], m);
              if (ph.aload) prefetch16(ph.aload, m, ph.al_ld, c0, ph.al_w, lane);
            } else {
              const long long tn = tX + (2 + s) * G;
              if (tn < ntiles) {
                prefetch(a.ph[0], tn * 128 + row);
                prefetch16(a.a0, tn * 128 + row, a.a0_ld, c0, a.a0_w, lane);
              }
            }
          }
          if (last && a.pf1_next) {
            // tile boundary: the slot's next tile starts with loads nothing has announced yet (its A operand, the
            // auxiliary operands of phase 0) - bring this thread's lines into L1 while the last phase runs
            const long long tn = tX + (2 + s) * G;
            if (tn < ntiles) {
              const long long mn = tn * 128 + row;
              if (c0 < a.a0_w) {
                const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.a0) + blk_base(mn, a.a0_ld, c0, 0));
                for (int g = 0; g < 8 && c0 + 8 * g < a.a0_w; ++g) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + g * 128));
              }
              const Phase& p0 = a.ph[0];
              if (c0 < p0.width) {
                if (p0.aux0) {
                  const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p0.aux0) + blk_base(mn, p0.ld0, c0, 0));
                  for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + g * 128));
                }
                if (p0.aux1) {
                  const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p0.aux1) + blk_base(mn, p0.ld1, c0, 0));
                  for (int g = 0; g < 4; ++g) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + g * 128));
                }
              }
            }
          }
          const uint32_t tA = lane_base + 256u + (uint32_t)(s * 128 + hq * 32);
          uint16_t* stb = a.stash ? a.stash + (((size_t)blockIdx.x * 2 + s) * MAX_STASH * 8) * (512 * 8) + (size_t)tid * 8 : nullptr;
          const uint32_t bf = smem_u32(&d_full), bd = smem_u32(&d_drained), ba = smem_u32(&a_ready[s]), par = dcnt & 1;
          ++dcnt;
          long long* dw = nullptr;
          (void)dw;
#ifdef VDN_CHAIN_TL
          if (dbg && blockIdx.x == 0 && lane == 0)
            dw = dbg + ((((tX - blockIdx.x) / (2 * G)) * MAX_PHASES + p) * 2 + s) * 48 + 4 + 2 * warp;
#endif
          switch (ph.op) {
            case OP_OUT32: ok = phase_body<OP_OUT32>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_SOFTPLUS: ok = phase_body<OP_SOFTPLUS>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_NSTEP: ok = phase_body<OP_NSTEP>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_P1STEP: ok = phase_body<OP_P1STEP>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_P2STEP: ok = phase_body<OP_P2STEP>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_RELU: ok = phase_body<OP_RELU>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_LINEAR: ok = phase_body<OP_LINEAR>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            case OP_STASH: ok = phase_body<OP_STASH>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
            default: ok = phase_body<OP_MASK>(ph, last, s, tD, tA, c0, m, a.N, sb, sRow, stb, bf, par, bd, ba, sig, isig, dw, stail); break;
          }
          VDN_TL(dw, 1);
          if (!last) {
            if (ph.a_out) {
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(smem_u32(&a_ready[s]));
            } else if (ph.aload) {
              aload(s, tile, ph.aload, ph.al_ld, ph.al_w);
            }
          } else {
            const long long tn = tX + (2 + s) * G;     // the slot's next tile
            if (tn < ntiles) aload(s, tn, a.a0, a.a0_ld, a.a0_w);
          }
        }
      }
    }
  } else if (tid == EPI_WARPS * 32) {
    // ================= MMA issuer: A from tensor memory, weights from shared memory =================
    uint32_t acnt[2] = {0, 0};
    uint32_t wt = 0, drained = 0;
    const uint32_t tAcol = tmem_base + 256u;
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int p = 0; p < a.P && ok; ++p) {
        const Phase& ph = a.ph[p];
        const uint32_t idesc = umma_idesc_f16(128, (uint32_t)ph.n_mma);
        const int nkb = (ph.nks + 3) >> 2;                       // K blocks streamed per phase
        const int keep = nkb < KEEP ? nkb : KEEP;
        for (int s = 0; s < 2 && ok; ++s) {
          if (s && !hasY) break;
#ifdef VDN_CHAIN_TL
          long long* di = nullptr;
          if (dbg && blockIdx.x == 0) di = dbg + ((((tX - blockIdx.x) / (2 * G)) * MAX_PHASES + p) * 2 + s) * 48;
#endif
          VDN_TL(di, 0);
          // the accumulator of the previous phase must have been drained (first phase ever: passes immediately)
          ok = mbar_wait(smem_u32(&d_drained), (drained & 1) ^ 1);
          ++drained;
          VDN_TL(di, 1);
          ok = ok && mbar_wait(smem_u32(&a_ready[s]), acnt[s] & 1);
          ++acnt[s];
          VDN_TL(di, 2);
          tc_fence_after();
          // Slot X streams all blocks of the phase in order and leaves the LAST `keep` in the ring; slot Y uses those
          // first (no wait), releases them, then takes the others, which were fetched again behind them.  (Holding the
          // last blocks rather than the first keeps the ring deadlock-free when a phase has more blocks than stages.)
          for (int j = 0; j < nkb && ok; ++j) {
            const bool held = (s == 1 && j < keep);
            const int kb = s == 0 ? j : (held ? nkb - keep + j : j - keep);
            const uint32_t w = wt + (uint32_t)(s == 0 || held ? kb : nkb + kb);
            const uint32_t ws = w % WSTAGES, wph = (w / WSTAGES) & 1;
            if (!held) {
              ok = mbar_wait(smem_u32(&w_full[ws]), wph);
              tc_fence_after();
            }
            const uint32_t b0 = sW + ws * W_STAGE;
            const int nk = ph.nks - 4 * kb < 4 ? ph.nks - 4 * kb : 4;
            // unrolled: the four descriptors are independent, so the issue rate is set by the MMA instructions themselves
            // and not by a serial address computation per instruction (the issuer shares its scheduler with four
            // epilogue warps)
            const uint32_t a0c = tAcol + (uint32_t)(s * 128 + ph.a_col + kb * 32);
            const uint64_t d0 = umma_desc_sw128(b0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < nk) umma_f16_ts(tmem_base, a0c + (uint32_t)(ks * 8), d0 + (uint64_t)(ks * 2), idesc, (j | ks) ? 1u : 0u);
            if (s == 1 || !hasY || kb < nkb - keep) umma_commit(smem_u32(&w_empty[ws]));
          }
          umma_commit(smem_u32(&d_full));
          VDN_TL(di, 3);
        }
        wt += (uint32_t)(hasY ? 2 * nkb - keep : nkb);
      }
    }
  } else if (tid == (EPI_WARPS + 1) * 32) {
    // ================= weight stream (TMA engine), in the order the MMA issuer consumes the blocks =================
    uint32_t wt = 0;
    for (long long tX = blockIdx.x; tX < ntiles && ok; tX += 2 * G) {
      const bool hasY = tX + G < ntiles;
      for (int p = 0; p < a.P && ok; ++p) {
        const Phase& ph = a.ph[p];
        const uint32_t bytes = (uint32_t)ph.n_mma * 128u;
        const int nkb = (ph.nks + 3) >> 2;
        const int keep = nkb < KEEP ? nkb : KEEP;
        const int nfetch = hasY ? 2 * nkb - keep : nkb;
        for (int f = 0; f < nfetch && ok; ++f, ++wt) {
          const int kb = f < nkb ? f : f - nkb;                  // slot X: all blocks; slot Y: the first nkb - keep again
          const uint32_t ws = wt % WSTAGES, wph = (wt / WSTAGES) & 1;
          ok = mbar_wait_backoff(smem_u32(&w_empty[ws]), wph ^ 1, 256);
          mbar_arrive_expect_tx(smem_u32(&w_full[ws]), bytes);
          bulk_g2s(sW + ws * W_STAGE, a.packed + ph.img_off + ((size_t)(ph.kb0 + kb) * ph.img_rows + ph.row0) * 32, bytes,
                   smem_u32(&w_full[ws]));
        }
      }
    }
  }
  if (!ok && fault) *fault = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------------------
inline Phase make_phase() {
  Phase p;
  memset(&p, 0, sizeof(p));
  p.dsc = 1.0f; p.bias_off = -1; p.bias_mul = 1.0f; p.a_mul = 1.0f; p.o16a_mul = 1.0f; p.o32_mul = 1.0f; p.tail_mul = 1.0f;
  p.r1_mul = 1.0f; p.stash_w = -1; p.stash_r = -1;
  return p;
}
inline void init_args(Args* a) {
  memset(a, 0, sizeof(*a));
  a->row_off[0] = a->row_off[1] = -1;
}
// B = rows [row0, row0 + n) of the fp16 image at img_off ([kblocks][img_rows][64]), K blocks kb0 ..
inline void set_mma(Phase* p, long long img_off, int img_rows, int row0, int kb0, int n, int k, int a_col = 0) {
  p->img_off = img_off; p->img_rows = img_rows; p->row0 = row0; p->kb0 = kb0;
  p->n_mma = (n + 15) & ~15; p->nks = (k + 15) / 16; p->a_col = a_col;
}
inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}
inline int grid_for(long long N) {
  const long long ntiles = (N + 127) / 128;
  const int sms = num_sms();
  return (int)(ntiles < sms ? ntiles : sms);
}

inline size_t stash_floats(long long N) { return (size_t)grid_for(N) * 2 * MAX_STASH * 8 * 512 * 8 / 2; }

// Debug aid (environment variable VDN_SYNC=1): synchronise after every launch of the new kernels and name the one that
// failed on stderr.  Off by default: the library never synchronises.
static inline int debug_sync(cudaStream_t st, const char* what, int tag) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VDN_SYNC"); on = (e && e[0] == '1') ? 1 : 0; }
  cudaError_t e = (cudaError_t)::vdn::take_launch_error();
  if (on && e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (on && e != cudaSuccess) fprintf(stderr, "[vdn] %s (tag %d) failed: %s\n", what, tag, cudaGetErrorString(e));
  return (int)e;
}

// invalid table / descriptor (a bug in the caller, fatal for the call): say where
static inline int bad_value(const char* where, int detail) {
  fprintf(stderr, "[vdn] invalid value: %s (%d)\n", where, detail);
  return (int)cudaErrorInvalidValue;
}
static inline int trace_err(int e, const char* where) {
  if (e) fprintf(stderr, "[vdn] %s: CUDA error %d (%s)\n", where, e, cudaGetErrorString((cudaError_t)e));
  return e;
}

// Validates the table (the kernel trusts it) and launches.  Returns a cudaError_t value.
static inline int launch(const Args& a_in, cudaStream_t st, int family) {
  Args a = a_in;
  {
    static int pf2 = -1;
    if (pf2 < 0) { const char* e = getenv("VDN_PF2"); pf2 = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0; }
    a.pf_mode = pf2;
    static int pf1n = -1;
    if (pf1n < 0) { const char* e = getenv("VDN_PF1NEXT"); pf1n = (e && e[0] == '0') ? 0 : 1; }      // on by default (-1 % step time, measured)
    a.pf1_next = pf1n;
  }
  for (int p = 0; p < a.P; ++p) {
    Phase& ph = a.ph[p];
    const bool common = !ph.o32b && !ph.r1 && ph.stash_r < 0 && ph.op != OP_OUT32 && ph.op != OP_STASH &&
                        (!(ph.op == OP_NSTEP || ph.op == OP_P1STEP || ph.op == OP_P2STEP) || ph.aux0) &&
                        (ph.op != OP_P1STEP || ph.aux1);
    // regular phase: every column of an active warp is an output, nothing else happens
    const bool whole = common && (ph.width & 63) == 0 && (!ph.a_out || ph.a_wr == ph.width) && !ph.o32 && !ph.tail;
    // skip layer: the outputs end inside the last 64-column block, which also carries the tail / zero padding up to a_wr and
    // possibly an fp32 side output of exactly those columns - that block takes the general path, the others the fast one
    const bool edge = common && !whole && ph.a_out && ph.a_wr == 256 && ph.width > 192 && ph.width < 256 &&
                      (!ph.o32 || ph.o32_c0 >= 192);
    ph.fast = (whole || edge) ? ph.width / 64 : 0;
    ph.edge = edge ? 1 : 0;
    ph.o32_vec = (ph.o32 && ph.act == 0 && ((uintptr_t)ph.o32 & 15) == 0 && (ph.ldo32 & 3) == 0 && (ph.o32_c0 & 7) == 0) ? 1 : 0;
  }
  if (a.N <= 0) return 0;
  if (a.P < 1 || a.P > MAX_PHASES || !a.a0 || (a.a0_w & 7) || a.a0_w > 256) return bad_value("chain args", a.P);
  double flops = 0.0, bytes = 0.0;
  for (int p = 0; p < a.P; ++p) {
    const Phase& ph = a.ph[p];
    if (ph.n_mma < 16 || ph.n_mma > 256 || (ph.n_mma & 15) || ph.nks < 1 || ph.nks > 16 || (ph.row0 & 7) ||
        ph.width > ph.n_mma || (ph.a_out && (p == a.P - 1 || (ph.a_wr & 7) || ph.a_wr > 256)) ||
        (ph.aload && (ph.a_out || p == a.P - 1 || (ph.al_w & 7))) || (ph.img_off & 255))
      return bad_value("chain phase shape", p);
    if ((ph.op == OP_STASH || ph.stash_r >= 0) && !a.stash) return bad_value("chain phase stash", p);
    if (ph.tail && (ph.tail_w < 1 || ph.tail_w > 64 || (ph.ldt & 7))) return bad_value("chain phase tail", p);
    // every operand is fp16 (OpTraits); cotangent scaling needs the launch's sigma
    if ((ph.o16b && ph.op != OP_P1STEP) || (ph.o16a_mul != 1.0f && ph.op != OP_P2STEP) ||
        ((ph.r1_scaled || ph.o32_unscale) && !a.sigma))
      return bad_value("chain phase format", p);
    flops += 2.0 * (double)a.N * ph.n_mma * ph.nks * 16;
    const double row16 = 2.0 * (double)a.N;
    if (ph.aux0) bytes += row16 * ph.width;
    if (ph.aux1) bytes += row16 * ph.width;
    if (ph.o16a) bytes += row16 * (ph.a_out ? ph.a_wr : ph.width);
    if (ph.o16b) bytes += row16 * ph.width;
    if (ph.o32) bytes += 4.0 * (double)a.N * ph.o32_w;
    if (ph.o32b) bytes += 4.0 * (double)a.N * ph.o32b_w;
    if (ph.aload) bytes += row16 * ph.al_w;
  }
  bytes += 2.0 * (double)a.N * a.a0_w;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  prof_begin(family, st, flops, bytes);
  long long* dbg = nullptr;
#ifdef VDN_CHAIN_TL
  {   // record the VDN_TL_LAUNCH-th chain launch of this translation unit
    static int tl_n = 0;
    const char* e = getenv("VDN_TL_LAUNCH");
    if (e && atoi(e) == tl_n) dbg = g_tc_dbg;
    ++tl_n;
  }
#endif
  VDN_LAUNCH(chain_kernel, grid_for(a.N), THREADS, SMEM, st, a, g_tc_fault, dbg);
  prof_end(family, st);
  return debug_sync(st, "chain_kernel", a.ph[0].op);
}

}  // namespace ce
}  // namespace vdn

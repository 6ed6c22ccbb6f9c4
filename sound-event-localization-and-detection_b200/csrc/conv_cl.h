// Channels-last tensor-core path (conv_cl.cu, wgrad_cl.cu): operand layouts, parameter blocks and
// host entry points.
//
// "CL operand" = bf16 copy of an activation (or gradient) tensor laid out [n][h][w][Cp]: channels
// innermost, each of the nc quaternion components padded to a multiple of 16 channels
// (Cp = nc * cpad, pad channels are zero).  A TMA box [128 w x 64 ch] of it is a K-major UMMA A
// operand in the canonical 128B-swizzled layout; a convolution tap is an offset on the w / h
// coordinates (any value, out-of-range rows are zero-filled = padding), so no shifted copies are
// needed; every 16-channel K slab lies inside one component, so one tcgen05.mma serves one
// (out component, in component) block and carries the block's Hamilton sign in its negate bit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
namespace cl {

constexpr int kTileM = 128;     // positions per accumulator tile (TMEM lanes)
constexpr int kThreads = 320;   // wgrad: warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kMmaWarps = 4;          // fprop: MMA-issuing warps (warp 1 and warps 6 ...)
constexpr int kEpiSets = 3;           // fprop: epilogue warp sets of 4 (one warp per TMEM lane quarter); the sets split
                                      // the 16-channel column chunks of a tile between them
constexpr int kFirstExtraEpiWarp = 6 + (kMmaWarps - 1);
constexpr int kFpropThreads = 192 + 32 * (kMmaWarps - 1) + 128 * (kEpiSets - 1);
constexpr int kMaxTaps = 9;
constexpr int kMaxStages = 8;
constexpr int kOpTableEntries = 512;  // (16-channel slab, out component) pairs: 1024 padded channels x 8

SELDQ_HD int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct OperandLayout {
  int nc;      // components that are padded separately (1 in dense mode)
  int cc;      // real channels per component
  int cpad;    // padded channels per component (multiple of 16)
  int Cp;      // nc * cpad
  int BK;      // channels per TMA box: 64 (128B swizzle), 32 (64B) or 16 (32B)
};

// dense = the layer's K side has fewer than 8 channels per component (first CNN layer) or the algebra
// is real: the tensor is laid out as ONE component and the signed expanded weight tile is built in
// shared memory only.
inline OperandLayout operand_layout(int nc, int channels, bool dense) {
  OperandLayout l;
  if (dense) {
    l.nc = 1; l.cc = channels;
    const int p16 = round_up(channels, 16);
    l.cpad = (p16 == 16 || p16 == 32 || p16 == 64) ? p16 : round_up(channels, 64);
  } else {
    l.nc = nc; l.cc = channels / nc;
    l.cpad = round_up(l.cc, 16);                               // nc = 4 | 8  =>  Cp is a multiple of 64
  }
  l.Cp = l.nc * l.cpad;
  l.BK = l.Cp < 64 ? l.Cp : 64;
  return l;
}

// is the K side of this pass a "dense" operand?  (kc = per-component channels on the K side)
inline bool is_dense(const ConvGeom& g) {
  const int kc = g.transposed ? g.Oc : g.Ic;
  return g.tab.nc == 1 || kc < 8;
}

struct FpropParams {
  ConvGeom g;                   // pass orientation (transposed = 1 for dgrad); dense prologue reads it
  const float* w[8];            // compact fp32 weights (dense prologue only)
  const uint8_t* packed;        // pre-packed bf16 weight images (non-dense)
  const float* bias;
  float* out;                   // fp32 NCHW ...
  __half* out16;                // ... or, when not null, IEEE fp16 NCHW instead (storage format of the fused CNN path)
  long long out_sN, out_sC, out_sH;
  int N, OH, OW;
  int tiles_w, total_tiles, ngroups, total_units;
  int group_order[8];           // heaviest groups first
  int ntaps;
  int off_h[kMaxTaps], off_w[kMaxTaps];
  // K side
  int BK, chunks, slabs_per_chunk, cpad_in, J;
  uint32_t stage_bytes, a_sbo, a_swz;
  uint32_t box_bytes;           // one TMA box [128 w x BK ch]; a stage holds `tps` of them (taps per stage)
  int tps;
  // Row-shared taps (rs = 1): the `tps` = KW taps of one kernel row differ only by a small shift along w, so one
  // box of 128 + span rows serves them all -- each tap's A descriptor starts rs_row[tap % KW] rows into the box (the
  // tensor core derives the 128B-swizzle phase from the address, tools/umma_probe.py u3) -- and the activation
  // tile crosses L2 -> shared memory once per kernel row instead of once per tap.
  int rs, rs_min_off, box_rows;
  uint32_t stage_tx;            // bytes one stage's TMA traffic signals on its mbarrier
  int rs_row[kMaxTaps];
  uint32_t chunk_mask[8];       // per group: which channel chunks carry at least one non-zero block
  // N side
  int dense, ncomp_out, Pc, NBp, gc;
  int n_img;
  uint32_t slab_bytes;          // NBp * 32: one [NBp x 16] bf16 B tile
  uint32_t img_bytes;           // ntaps * J * slab_bytes
  uint32_t w_bytes;             // n_img * img_bytes
  int8_t op_img[8][8];          // [in comp b][out comp a] -> image index or -1
  int8_t op_neg[8][8];
  int nstages, acc_stages, tmem_cols;
  // Pair fusion (fuse = 1).  A small-N tcgen05.mma costs ~45 cycles whatever its N (it is bound by loading its A
  // slab: tools/umma_ts.py), so the per-block MMAs of N = 24..48 leave half of the tensor pipe idle.  Two out
  // components a0, a1 = a0 ^ pair_xor always read the two images of one image pair {i, i ^ pair_xor}; with the
  // pair's tiles interleaved by 8-row groups in shared memory ([8 rows of image i][8 rows of i ^ pair_xor] ...) ONE
  // MMA of N = 2 * NB8 serves both.  Which slot holds whose product, and the sign of a1's product relative to a0's
  // (a0's own sign is the negate-B bit), depend on the in component only through TWO classes, so each pair
  // accumulates into two column sets and the epilogue adds / subtracts them (epi_col / epi_sgn).
  int fuse, pair_xor;           // fuse = F: 0 (off), 2 (pairs, pair_xor = 1 | 2) or 4 (quads, pair_xor = 3: four column
                                // sets, one per in component mod 4 -- fits where 16 * Pc <= 512, the CNN layers)
  int NB8;                      // out channels of one component rounded up to 8 (fuse) -- NBp rounds to 16
  int NBmma;                    // N of one MMA: NBp, or 2 * NB8 when fused
  int mma_per_slab;             // MMAs per 16-channel slab and unit: gc, or gc / 2 when fused
  int acc_cols;                 // accumulator columns of one unit
  int op_entries;               // valid length of op_tbl
  int8_t comp_of[8][8];         // [group][local component] -> out component
  uint16_t epi_col[8][8][4];    // fused: [group][local component][set] -> first accumulator column of its 8-channel
  int8_t epi_sgn[8][8][4];      //        groups (+ og * 8 F), and the sign to apply (0: the set does not exist)
  // MMA op table [group][chunk][lane]: the MMAs of one stage dealt to the lanes of the MMA warp (copied to
  // shared memory):
  //   x = valid << 31 | first << 30 | last-of-stage << 29 | accumulator column offset << 20 (9 bits) | slab << 16 |
  //       (16-byte offset of the weight tile inside one tap's tile set)
  //   y = the complete instruction descriptor (carries the block's sign in its negate-B bit)
  // `first` marks the first MMA of a unit into that out component's accumulator columns (tap 0 only)
  uint32_t tap_stride16;        // J * slab_bytes >> 4
  // Sibling launch (nprob = 2): two convolutions of EQUAL geometry -- conv1_filter / conv1_gate or conv2_skip /
  // conv2_residual of a residual block (model.py:118-119, :130-131), forward or dgrad -- share one launch: even
  // CTAs run problem 0, odd CTAs problem 1, each with its own weight tiles, input map and output.  (At batch 1 a
  // launch is bound by its fixed costs, not by its ~2 us of MMAs: one launch for two layers halves them.)
  int nprob;
  const uint8_t* packed1;       // problem 1's weight tiles
  float* out1;                  // problem 1's fp32 output
  // Fused glue of the fp32 output path, per problem (model.py:114-132):
  //   epi_mode 0: out = v    1: out += v (running sum of the skip outputs)    2: out = addend + v (x + residual)
  //   stats != null: per-channel (sum, sum of squares) of the STORED values are added there (double[2 C], zeroed by
  //   the caller): the BatchNorm batch statistics of the layer that follows, so no separate pass reads the output
  int epi_mode[2];
  const float* addend[2];
  double* stats[2];
  // optional timeline of CTA 0 (tools/fprop_trace.py): when not null, the roles write globaltimer stamps here
  unsigned long long* trace;
  uint2 op_tbl[kOpTableEntries];
};

// the fused-glue options of one problem as the host passes them (seldq_conv_epilogue_t of include/seldq.h)
struct FpropEpilogue {
  int mode = 0;
  const float* addend = nullptr;
  double* stats = nullptr;
};

// wgrad: D[(a,o), (b,i)] per tap = sum_t GY[(a,o), t] * X[t + off(tap), (b,i)]
//   A = gy from its pitched NCHW bf16 copy (K-major, time contiguous), rows gathered per component
//   B = x from its CL operand (MN-major, channels contiguous), the tap is a row offset
constexpr int kWgradStages = 4;
struct WgradParams {
  ConvGeom g;                   // forward orientation
  float* gw[2][8];              // compact fp32 gradients of up to two problems, accumulated with atomicAdd
  int nprob;                    // 1, or 2: two convolutions of equal geometry that read the same x (e.g. the filter
                                // and gate convolutions of a residual block) share one launch
  int ncomp, OS;                // M rows = ncomp * OS (= 128)
  int o_tiles;
  int cpad_in, Cp, nchunks;     // x operand: Cp = ncomp*cpad_in channels = nchunks boxes of 64
  int taps_per_group, tap_groups, ntaps;
  int off_h[kMaxTaps], off_w[kMaxTaps];
  int OH, OW, N;
  int chunks_w;
  long long ksteps;
  int splits, nstages, tmem_cols;
  int x_one_box;                // x operand: one rank-5 TMA box per tap (all channel chunks) instead of one per chunk
  uint32_t stage_bytes, b_tap_bytes;
  int8_t pair_n[8];             // which (a,b) blocks feed compact tensor e
  int8_t pair_a[8][8], pair_b[8][8], pair_neg[8][8];
  unsigned long long* trace;    // debug timeline of CTA (0, 0) (seldq_debug_fprop_trace, tools/wgrad_trace.py)
};

int num_sms();
void set_fprop_trace(void* dev_buf);      // debug: timeline of CTA 0 of every later fprop launch (null: off)
unsigned long long* fprop_trace();

}  // namespace cl

// x / gy operand layouts of a convolution (forward-orientation geometry)
cl::OperandLayout x_operand_layout(const ConvGeom& fwd);
cl::OperandLayout gy_operand_layout(const ConvGeom& fwd);

size_t packed_weight_bytes(const ConvGeom& pass_geom);
int launch_pack_weights(const ConvGeom& pass_geom, const float* const* host_w, void* packed, cudaStream_t st);
size_t pack_table_entry_bytes();
int fill_pack_table_entry(const ConvGeom& pass_geom, const float* const* host_w, void* packed, void* entry, int* items);
int launch_pack_weights_multi(const void* dev_table, int count, int max_items, cudaStream_t st);
int launch_cl_fprop(const ConvGeom& g, const void* in_cl, const float* const* host_w, const void* packed,
                    const float* bias, float* out, void* out_f16, cudaStream_t st,
                    const cl::FpropEpilogue* epi = nullptr);
// two sibling convolutions of equal geometry in one launch (FpropParams::nprob)
int plan_cl_fprop_pair(const ConvGeom& g, cl::FpropParams* p, size_t* smem_bytes);
int launch_cl_fprop_pair(const ConvGeom& g, const void* const in_cl[2], const void* const packed[2], float* const out[2],
                         const cl::FpropEpilogue epi[2], cudaStream_t st);
int launch_cl_wgrad(const ConvGeom& g, const void* x_cl, const void* gy_nchw16, float* const* host_gw,
                    cudaStream_t st, const void* gy2_nchw16 = nullptr, float* const* host_gw2 = nullptr);
// fp32 NCHW -> CL operand (and, optionally, the pitched NCHW bf16 copy wgrad reads gy from)
int launch_stage_operand(const float* src, void* dst_cl, void* dst_nchw16, const cl::OperandLayout& l, int n, int c,
                         int h, int w, cudaStream_t st);
inline int nchw16_pitch(int w) { return (w + 7) & ~7; }

}  // namespace seldq

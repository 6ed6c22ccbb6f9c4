// Parameter blocks and host entry points of the tcgen05 kernels (conv_umma.cu, wgrad_umma.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
namespace umma {

constexpr int kTileM = 128;     // positions per accumulator tile (TMEM lanes)
constexpr int kThreads = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kMaxTaps = 9;
constexpr int kMaxStages = 8;
constexpr int kMaxAtoms = 96;   // 8-channel K atoms in one component's K sequence

struct FpropParams {
  ConvGeom g;                   // pass orientation (transposed = 1 for dgrad)
  const float* w[8];            // compact fp32 weights (device)
  const float* bias;
  float* out;
  __nv_bfloat16* out_bf16;      // optional bf16 mirror of out (row pitch o16_sH)
  long long out_sN, out_sC, out_sH;
  long long o16_sN, o16_sC, o16_sH;
  int N, OH, OW, P;
  int tiles_w, total_tiles;
  int dense;                    // 1: signed expanded tile in smem, one MMA spans all out channels
  int ncomp_in, ncomp_out;
  int Rc;                       // in channels per component  = TMA box rows
  int Pc;                       // out channels per component
  int NBp;                      // Pc rounded up to 16        = MMA N and TMEM column stride
  int n_img;                    // resident weight images (compact tensors)
  int ntaps, G, stages_per_comp, atoms_per_tap, stage_atoms, kpairs;
  int nstages, acc_stages, tmem_cols;
  int debug;                    // bring-up switches (env SELDQ_DEBUG): 1 no MMA, 2 no TMA, 4 no TMEM loads
  int off_h[kMaxTaps], off_w[kMaxTaps];
  int tap_sidx[kMaxTaps];       // which shifted mirror serves the tap; off_w already includes its shift
  int nops[8];
  int8_t op_img[8][8], op_neg[8][8], op_out[8][8];
  int8_t atom_tap[kMaxAtoms], atom_chan[kMaxAtoms];
};

// wgrad: D[(a,o), (tap,b,i)] = sum_t G[(a,o), t] * X[(b,i), t + off(tap)]   (both operands K-major)
constexpr int kWgradStages = 4;
struct WgradParams {
  ConvGeom g;                   // forward orientation
  float* gw[8];                 // compact fp32 gradients (device), accumulated with atomicAdd
  int dense;                    // rows / cols are plain expanded channels (first layer)
  int ncomp;                    // component groups on each side (dense: 1)
  int OS, IS;                   // rows / cols per component group  (M = ncomp*OS <= 128, NW = ncomp*IS)
  int NW;                       // columns per tap
  int o_tiles, i_tiles, tap_groups, taps_per_group;
  int ntaps;
  int off_h[kMaxTaps], off_w[kMaxTaps];
  int tap_sidx[kMaxTaps];
  int OH, OW, N;
  int chunks_w;                 // ceil(OW / 64)
  long long ksteps;             // N * OH * chunks_w
  int splits;
  int nstages;
  int tmem_cols;
  int m_rows;                   // 64 or 128
};

int num_sms();

}  // namespace umma

// A "mirror set" is the bf16 copy of an NCHW / NCW fp32 tensor the TMA path reads: nshifts copies
// laid out [shift][n][c][h][pitch], copy k holding the rows shifted RIGHT by shifts[k] elements
// (zeros shifted in), pitch = mirror_pitch(w).  TMA needs the innermost box coordinate 16-byte
// aligned, so a tap whose offset is off reads the copy with shift = (-off) mod 8 at the aligned
// coordinate w0 + off + shift.  shifts[0] is always 0.
struct MirrorSet {
  const void* data;
  int nshifts;
  int shifts[8];
};
inline int mirror_pitch(int w) { return (w + 7 + 7) & ~7; }
// shift list a pass needs: which = 0 for the tensor read with the forward taps (x: fwd, wgrad),
// which = 1 for the tensor read with the transposed taps (gy: dgrad)
void mirror_shifts(const ConvGeom& fwd_geom, int which, int* shifts, int* nshifts);
// (W, H, C, N, shift) tensor map over a mirror set, box {64, 1, box_rows, 1, 1}, 128-byte swizzle
int encode_mirror_map(CUtensorMap* tm, const MirrorSet& m, int w, int h, int c, int n, int box_rows);

int launch_umma_fprop(const ConvGeom& g, const MirrorSet& in, const float* const* host_w, const float* bias,
                      float* out, cudaStream_t st);
int launch_umma_wgrad(const ConvGeom& g, const MirrorSet& x, const MirrorSet& gy, float* const* host_gw,
                      cudaStream_t st);

}  // namespace seldq

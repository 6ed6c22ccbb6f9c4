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
  int pitch;      // row pitch in elements; 0 = mirror_pitch(w)
};
inline int mirror_pitch(int w) { return (w + 7 + 7) & ~7; }
// shift list a pass needs: which = 0 for the tensor read with the forward taps (x: fwd, wgrad),
// which = 1 for the tensor read with the transposed taps (gy: dgrad)
void mirror_shifts(const ConvGeom& fwd_geom, int which, int* shifts, int* nshifts);
// (W, H, C, N, shift) tensor map over a mirror set, box {64, 1, box_rows, 1, 1}, 128-byte swizzle
int encode_mirror_map(CUtensorMap* tm, const MirrorSet& m, int w, int h, int c, int n, int box_rows);

int launch_umma_wgrad(const ConvGeom& g, const MirrorSet& x, const MirrorSet& gy, float* const* host_gw,
                      cudaStream_t st);

}  // namespace seldq

// Shifted bf16 "mirror sets" (conv_umma.h) and their TMA views.  They serve the one layer whose K side is
// too narrow for the channels-last operand layout to be used by the weight-gradient kernel: the first CNN
// layer (Cin = 8 | 16), see wgrad_umma.cu.  Every other pass reads channels-last operands (conv_cl.h).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "launch.h"
#include "tensor_map.h"

namespace seldq {

// shifts[0] = 0; then one entry per distinct (-off_w) mod 8 over the taps of the pass
void mirror_shifts(const ConvGeom& fwd_geom, int which, int* shifts, int* nshifts) {
  int n = 0;
  shifts[n++] = 0;
  for (int kw = 0; kw < fwd_geom.KW; ++kw) {
    const int off = which == 0 ? kw * fwd_geom.dw - fwd_geom.pw : fwd_geom.pw - kw * fwd_geom.dw;
    const int s = ((-off) % 8 + 8) % 8;
    bool seen = false;
    for (int i = 0; i < n; ++i) seen |= shifts[i] == s;
    if (!seen && n < 8) shifts[n++] = s;
  }
  *nshifts = n;
}

// (W, H, C, N, shift) view of a mirror set with box {64, 1, rows, 1, 1}, 128-byte swizzle
int encode_mirror_map(CUtensorMap* tm, const MirrorSet& m, int w, int h, int c, int n, int box_rows) {
  const uint64_t pitch = (uint64_t)(m.pitch > 0 ? m.pitch : mirror_pitch(w));
  const uint64_t dims[5] = {pitch, (uint64_t)h, (uint64_t)c, (uint64_t)n, (uint64_t)m.nshifts};
  const uint64_t strides[4] = {pitch * 2, pitch * h * 2, pitch * h * c * 2, pitch * h * c * n * 2};
  const uint32_t box[5] = {64, 1, (uint32_t)box_rows, 1, 1};
  return encode_tensor_map(tm, m.data, 2, 5, dims, strides, box, /*128B*/ 3);
}

}  // namespace seldq

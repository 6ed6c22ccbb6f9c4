// Parameter block and host entry points of the fused first-layer backward kernel (wgrad_first.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "conv_umma.h"
#include "epilogue.h"

namespace seldq {
namespace first {

struct Params {
  ConvGeom g;                   // forward orientation
  float* gw[8];                 // compact fp32 gradients, accumulated with atomicAdd
  // CNN-block tail (epilogue.h): conv output, BN coefficients, arg-max flags, pooled gradient, BN-backward means
  const __half* y;              // conv output, fp16 (epilogue.cu)
  const float* coef;
  const uint8_t* idx;
  const float* gz;
  const float2* dmean;
  int pool;
  float scale;                  // inverted-dropout scale
  int N, C, H, W, HP;           // conv output (N, C, H, W), pooled height
  int R;                        // input channels = rows per tap of the A tile
  int ntaps;
  int off_h[umma::kMaxTaps], off_w[umma::kMaxTaps], tap_sidx[umma::kMaxTaps];
  int chunks_w;
  long long ksteps;             // N * H * chunks_w
  int rows;                     // ntaps * R <= 128
  int NB;                       // output channels rounded up to 16: N of the MMA
  uint32_t a_bytes, b_bytes, stage_bytes;
  int8_t pair_n[8];             // which (a, b) blocks feed compact tensor e
  int8_t pair_a[8][8], pair_b[8][8], pair_neg[8][8];
};

}  // namespace first

bool first_layer_bwd_supported(const ConvGeom& fwd);
int launch_first_layer_bwd(const ConvGeom& fwd, const MirrorSet& x, const epi::TailParams& tail, const float2* dmean,
                           float* const* host_gw, cudaStream_t st);
// the two BatchNorm-backward reductions of a CNN-block tail (epilogue.cu): dsums[2c], dsums[2c+1] and, behind the
// 2*C doubles, the fp32 means the apply step reads
int launch_cnn_tail_bwd_reduce(epi::TailParams& p, double* dsums, cudaStream_t st);

}  // namespace seldq

// Parameter block and host entry points of the bandwidth-bound glue kernels (epilogue.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
namespace epi {

struct TailParams {
  const __half* y;              // conv output, IEEE fp16 [N][C][H][W] (see epilogue.cu: why fp16 and not bf16)
  const float* coef;            // [C][4] = {a, b, mean, rstd} from bn_finalize_kernel
  int N, C, H, W, pool;
  // channels-last operand layout of the tensor being written (z forward, d(conv out) backward)
  int nc, cc, cpad, Cp;
  // forward outputs
  __nv_bfloat16* z_cl;          // [N][H/pool][W][Cp] or null
  float* z32;                   // [N][C][H/pool][W] or null
  uint8_t* idx;                 // [N][C][H/pool][W]: arg-max row | 0x80 keep flag (written fwd, read bwd)
  __half* ymax;                 // [N][C][H/pool][W]: conv output at the arg-max (written fwd, read bwd) or null
  // backward
  const float* gz;              // gradient w.r.t. the pooled output, fp32 [N][C][H/pool][W]
  __nv_bfloat16* d_t16;         // d(conv out), bf16 [N][C][H][pitch] or null
  __nv_bfloat16* d_cl;          // d(conv out), channels-last operand [N][H][W][Cp] or null
  int pitch;
  // dropout
  float drop_p;
  const long long* seed_ptr;    // device counter, advanced by the caller every step
  uint32_t salt;
  // tiling (filled by the launchers)
  int tiles_w, tiles_c;
  long long total_blocks;
};

}  // namespace epi

int launch_bn_stats(const void* src, int is_bf16, int n, int c, long long plane, double* sums, cudaStream_t st);
int launch_bn_finalize(const double* sums, const float* gamma, const float* beta, int c, double count, float eps,
                       float momentum, float* running_mean, float* running_var, float* coef, cudaStream_t st);
int launch_cnn_tail_fwd(epi::TailParams& p, cudaStream_t st);
int launch_cnn_tail_bwd(epi::TailParams& p, double* dsums, cudaStream_t st);

}  // namespace seldq

// Parameter block and host entry point of the TCN residual-block glue kernels (tcn_glue.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
namespace tcn {

enum Op {
  OP_PREACT_FWD = SELDQ_TCN_PREACT_FWD,
  OP_ROW_STATS = SELDQ_TCN_ROW_STATS,
  OP_GATE_FWD = SELDQ_TCN_GATE_FWD,
  OP_RESIDUAL_FWD = SELDQ_TCN_RESIDUAL_FWD,
  OP_GATE_BWD_REDUCE = SELDQ_TCN_GATE_BWD_REDUCE,
  OP_GATE_BWD_APPLY = SELDQ_TCN_GATE_BWD_APPLY,
  OP_PREACT_BWD_REDUCE = SELDQ_TCN_PREACT_BWD_REDUCE,
  OP_PREACT_BWD_APPLY = SELDQ_TCN_PREACT_BWD_APPLY,
  OP_GATE_BWD = SELDQ_TCN_GATE_BWD,
  OP_PREACT_BWD = SELDQ_TCN_PREACT_BWD,
  OP_GATE_FWD_STATS = SELDQ_TCN_GATE_FWD_STATS,
  OP_RESIDUAL_PREACT_FWD = SELDQ_TCN_RESIDUAL_PREACT_FWD
};

// one train-mode BatchNorm1d: batch statistics as per-channel (sum, sum of squares) in double
struct BnRef {
  const double* sums;
  const float* gamma;
  const float* beta;
  float* running_mean;   // momentum update by the forward kernels when not null
  float* running_var;
};

struct GlueParams {
  int N, C, T;           // tensors are fp32 [N][C][T]
  int C2;                // channels of the skip tensors (residual_fwd)
  int cc, cpad, Cp;      // channels-last operand layout of C channels (conv_cl.h)
  int pitch;             // pitched NCW operand: T rounded up to 8
  double count;          // elements per channel behind the statistics (N * T)
  double inv_count;
  float eps, momentum;
  BnRef bn[2];
  const float* in[5];
  float* out32;
  float* out32b;         // residual_preact_fwd: the next block's x = tanh(BN1(r'))
  __nv_bfloat16* out_cl[2];
  __nv_bfloat16* out_t16[2];
  double* dsums;         // reduce kernels: output (zeroed by the caller); apply kernels: input
  double* stats_out[2];  // row_stats
  float* accum;          // running sum of the skip connections
  float drop_p;
  const long long* seed_ptr;
  uint32_t salt;
  unsigned int* sync;    // fused (reduce + apply in one launch) kernels: grid-barrier counter, zeroed by the caller
  int tiles_t, tiles_c;  // filled by the launcher
  long long total_blocks;
};

}  // namespace tcn

int launch_tcn_glue(int op, tcn::GlueParams& p, int flag, cudaStream_t st);
// 1 when the single-launch (grid-barrier) steps can hold every tile of an (N, C, T) tensor resident at once
int tcn_glue_fused_supported(int n, int c, int t);

}  // namespace seldq

// Fused multi-head attention (attention.cu): host entry points.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {

int attention_supported(int batch, int heads, int seq, int head_dim);
// bytes of one row-major [BH][S][64] / one transposed [BH][d][S] bf16 operand copy
size_t attention_rm_bytes(int batch, int heads, int seq);
size_t attention_tr_bytes(int batch, int heads, int seq, int head_dim);
// layout 0: src is (N, E, S); layout 1: src is (N, S, E) and delta[bh][s] = sum_d src * o is written too (o, delta may
// be null).  rm / tr may not be null.
int launch_attention_stage(const float* src, const float* o, void* rm, void* tr, float* delta, int batch, int heads, int S,
                           int d, int layout, float scale, cudaStream_t st);
int launch_attention_fwd(int batch, int heads, int S, int d, const void* q_rm, const void* k_rm, const void* v_tr,
                         float* out, float* lse, cudaStream_t st);
int launch_attention_bwd(int batch, int heads, int S, int d, const void* q_rm, const void* k_rm, const void* v_rm,
                         const void* q_tr, const void* k_tr, const void* do_rm, const void* do_tr, float* lse,
                         const float* delta, float* dq, float* dk, float* dv, cudaStream_t st);

}  // namespace seldq

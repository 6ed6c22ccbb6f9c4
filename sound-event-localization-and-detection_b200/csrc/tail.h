// TC_Block tail glue (tail.cu): activation + MaxPool1d in one kernel per direction.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
int launch_act_pool_fwd(const float* x, float* y, long long rows, int T, int pool, int act, cudaStream_t st);
int launch_act_pool_bwd(const float* x, const float* y, const float* gy, float* gx, long long rows, int T, int pool, int act,
                        cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2, double eps, float* step,
                int advance, cudaStream_t st);
}  // namespace seldq

// Rotation weights and quaternion point-wise operators (rotation.cu): launchers.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
int launch_rotation_weight(const float* const* w, long long d0, long long d1, long long taps, int quaternion_format,
                           int transpose_out, float* out, cudaStream_t st);
int launch_rotation_weight_bwd(const float* const* w, const float* g_out, long long d0, long long d1, long long taps,
                               int quaternion_format, int transpose_out, float* const* gw, cudaStream_t st);
int launch_quaternion_pointwise(int op, const float* a, const float* b, float* out, long long outer, long long m,
                                cudaStream_t st);
}  // namespace seldq

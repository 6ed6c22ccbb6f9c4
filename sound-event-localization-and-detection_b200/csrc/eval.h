// Device side of the evaluation path (eval.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {
int launch_seld_events(const float* sed, const float* doa, int clips, int frames, int classes, int overlaps, float max_loc,
                       float* rows, int* counts, cudaStream_t st);
}

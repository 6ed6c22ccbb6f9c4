// Evaluation path, device side: the frame / class / overlap scan of gen_submission_list_task2
// (utility_functions.py:184-210) -- a Python triple loop over 600 x 42 cells per clip in the reference -- as one kernel
// over a batch of clips.  A cell is active when its SED output rounds to non-zero (np.round: half to even); an active
// cell yields the row [frame, class, x, y, z, overlap] with the DOA outputs scaled back by max_loc_value; rows come out
// in the reference's order (frame, then cell) through a per-clip prefix sum over the frames.
#include <cuda_runtime.h>

#include "eval.h"
#include "launch.h"

namespace seldq {
namespace eval {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) seld_events_kernel(const float* __restrict__ sed, const float* __restrict__ doa,
                                                              int frames, int cells, int overlaps, float max_loc,
                                                              float* __restrict__ rows, int* __restrict__ counts) {
  __shared__ int s_scan[kThreads];
  __shared__ int s_base;
  const int clip = blockIdx.x;
  const float* csed = sed + (size_t)clip * frames * cells;
  const float* cdoa = doa + (size_t)clip * frames * cells * 3;
  float* crows = rows + (size_t)clip * frames * cells * 6;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int f0 = 0; f0 < frames; f0 += kThreads) {
    const int f = f0 + threadIdx.x;
    int n = 0;
    if (f < frames)
      for (int j = 0; j < cells; ++j) n += rintf(csed[(size_t)f * cells + j]) != 0.f;
    // block-wide inclusive scan of the per-frame counts (Hillis-Steele; 600 frames: three rounds of this loop)
    s_scan[threadIdx.x] = n;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1) {
      const int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    int at = s_base + s_scan[threadIdx.x] - n;
    if (f < frames)
      for (int j = 0; j < cells; ++j)
        if (rintf(csed[(size_t)f * cells + j]) != 0.f) {
          float* r = crows + (size_t)at * 6;
          const float* l = cdoa + ((size_t)f * cells + j) * 3;
          r[0] = (float)f; r[1] = (float)(j / overlaps);
          r[2] = l[0] * max_loc; r[3] = l[1] * max_loc; r[4] = l[2] * max_loc;
          r[5] = (float)(j % overlaps);
          ++at;
        }
    __syncthreads();
    if (threadIdx.x == kThreads - 1) s_base += s_scan[kThreads - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[clip] = s_base;
}

}  // namespace eval

int launch_seld_events(const float* sed, const float* doa, int clips, int frames, int classes, int overlaps, float max_loc,
                       float* rows, int* counts, cudaStream_t st) {
  if (clips < 1 || frames < 1 || classes < 1 || overlaps < 1) return fail(SELDQ_ERR_INVALID, "seld_events: bad shape");
  eval::seld_events_kernel<<<clips, eval::kThreads, 0, st>>>(sed, doa, frames, classes * overlaps, overlaps, max_loc, rows,
                                                             counts);
  return check_launch("seld_events_kernel");
}

}  // namespace seldq

// __global__ wrapper of the STFT magnitude / phase kernel (phase functions in stft.cuh).
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "launch.h"
#include "stft.cuh"

namespace seldq {
namespace stft {

__global__ void __launch_bounds__(NT, 1) stft_magphase_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& s = *reinterpret_cast<Shared*>(smem_raw);
  init_tables(s, threadIdx.x);
  __syncthreads();
  Thread th;
  Stats acc = {0.f, 0.f, 0.f, 0.f};
  int buf = 0;
  long long batch = blockIdx.x;
  int signal = 0, t0 = 0;
  if (batch < p.total) {
    batch_decode(p, batch, &signal, &t0);
    phase_a_load(p, th, threadIdx.x, signal, t0);
  }
  for (; batch < p.total; batch += gridDim.x, buf ^= 1) {
    phase_a_compute(s, th, threadIdx.x);
    __syncwarp();                                  // a frame lives in one half-warp
    phase_b(s, th, threadIdx.x);
    __syncwarp();
    phase_b2(s, th, threadIdx.x);
    __syncwarp();
    phase_c(p, s, th, threadIdx.x, buf);
    // The only block-wide barrier of a batch.  It also orders this batch's writes of tile[buf] after the reads
    // phase_d made of it two batches ago: every thread passed the previous batch's barrier after those reads.
    __syncthreads();
    const int cur_signal = signal, cur_t0 = t0;
    const long long next = batch + gridDim.x;
    if (next < p.total) {                          // the next batch's samples are in flight during the stores below
      batch_decode(p, next, &signal, &t0);
      phase_a_load(p, th, threadIdx.x, signal, t0);
    }
    phase_d(p, s, acc, threadIdx.x, cur_signal, cur_t0, buf);
  }
  if (p.stats != nullptr) {
    // feature statistics (sum, sum of squares per plane): warp shuffle, then one double atomicAdd per warp and quantity
    const int planes = p.output_phase ? 2 : 1;
    for (int pl = 0; pl < planes; ++pl) {
      float a = pl == 0 ? acc.mag1 : acc.ph1, b = pl == 0 ? acc.mag2 : acc.ph2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(p.stats + 2 * pl, (double)a);
        atomicAdd(p.stats + 2 * pl + 1, (double)b);
      }
    }
  }
}

}  // namespace stft

int launch_stft(stft::Params& p, int n_signals, cudaStream_t st) {
  const size_t smem = sizeof(stft::Shared);
  static thread_local bool configured = false;
  if (!configured) {
    const cudaError_t e = cudaFuncSetAttribute(stft::stft_magphase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "stft smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  p.groups = (p.n_frames + stft::FRB - 1) / stft::FRB;
  p.total = (long long)n_signals * p.groups;
  const long long sms = cl::num_sms();
  const unsigned grid = (unsigned)(p.total < sms ? p.total : sms);     // one resident block per SM
  stft::stft_magphase_kernel<<<grid, stft::NT, smem, st>>>(p);
  return check_launch("stft_magphase_kernel");
}

}  // namespace seldq

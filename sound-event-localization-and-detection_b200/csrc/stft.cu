// __global__ wrapper of the STFT magnitude / phase kernel (phase functions in stft.cuh).
#include <cuda_runtime.h>

#include "launch.h"
#include "stft.cuh"

namespace seldq {
namespace stft {

__global__ void __launch_bounds__(NT) stft_magphase_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& s = *reinterpret_cast<Shared*>(smem_raw);
  if (threadIdx.x == 0) s.samples = reinterpret_cast<float*>(smem_raw + sizeof(Shared));
  __syncthreads();
  load(p, s, threadIdx.x, blockIdx.x, blockIdx.y);
  __syncthreads();
  for (int round = 0; round < FR / FPR; ++round) {
    if (blockIdx.x * FR + round * FPR >= p.n_frames) break;   // block-uniform: no frames left
    pack(p, s, threadIdx.x, round);
    __syncthreads();
    int src = 0;
#pragma unroll
    for (int Ns = 1; Ns < NC; Ns *= 4) {
      fft_pass(s, threadIdx.x, Ns, src);
      __syncthreads();
      src ^= 1;
    }
    emit(p, s, threadIdx.x, blockIdx.x, blockIdx.y, round);
    __syncthreads();
  }
}

}  // namespace stft

int launch_stft(const stft::Params& p, int n_signals, cudaStream_t st) {
  const size_t smem = sizeof(stft::Shared) + sizeof(float) * stft::span(p.hop);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    const cudaError_t e = cudaFuncSetAttribute(stft::stft_magphase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "stft smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  dim3 grid((p.n_frames + stft::FR - 1) / stft::FR, n_signals);
  if (grid.y > 65535) return fail(SELDQ_ERR_UNSUPPORTED, "more than 65535 signals in one STFT call");
  stft::stft_magphase_kernel<<<grid, stft::NT, smem, st>>>(p);
  return check_launch("stft_magphase_kernel");
}

}  // namespace seldq

// Host-side launcher declarations shared between the kernel translation units and seldq_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {

namespace simt {
struct ConvParams;
struct WgradParams;
}  // namespace simt
namespace stft {
struct Params;
}

// launch-configuration errors surface here; execution errors surface at the caller's next sync
inline int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SELDQ_OK;
}

int launch_conv_simt(const simt::ConvParams& p, cudaStream_t st);
int launch_wgrad_simt(simt::WgradParams& p, cudaStream_t st);
int launch_bias_grad(const float* gy, float* gb, int C, int N, int H, int W, long long sN, long long sC, long long sH,
                     long long sW, int accumulate, cudaStream_t st);
int launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st);
int launch_cast_bf16_mirror(const float* src, void* dst, long long rows, int w, int pitch, const int* shifts,
                            int nshifts, cudaStream_t st);
int launch_stft(stft::Params& p, int n_signals, cudaStream_t st);

}  // namespace seldq

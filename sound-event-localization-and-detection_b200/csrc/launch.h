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

// Fork / join of a library-internal side stream off a caller's stream (capturable into a CUDA graph): work launched on
// side() between fork() and join() runs concurrently with what the caller's stream does in between.  One side stream
// and event pair per host thread and call site (`slot`).
struct ForkJoin {
  cudaStream_t st, side_ = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool ok = false;
  ForkJoin(cudaStream_t caller, int slot) : st(caller) {
    struct Res { cudaStream_t s = nullptr; cudaEvent_t f = nullptr, j = nullptr; int dev = -1; };
    static thread_local Res res[4];
    int dev = 0;
    cudaGetDevice(&dev);
    Res& r = res[slot & 3];
    if (r.s == nullptr || r.dev != dev) {
      if (cudaStreamCreateWithFlags(&r.s, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&r.f, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&r.j, cudaEventDisableTiming) != cudaSuccess)
        return;
      r.dev = dev;
    }
    side_ = r.s; ev_fork = r.f; ev_join = r.j;
    ok = cudaEventRecord(ev_fork, st) == cudaSuccess && cudaStreamWaitEvent(side_, ev_fork, 0) == cudaSuccess;
  }
  cudaStream_t side() const { return ok ? side_ : st; }
  // always call, also after a failed launch: a capture must not be left forked
  bool join() { return !ok || (cudaEventRecord(ev_join, side_) == cudaSuccess && cudaStreamWaitEvent(st, ev_join, 0) == cudaSuccess); }
};

int launch_conv_simt(const simt::ConvParams& p, cudaStream_t st);
int launch_wgrad_simt(simt::WgradParams& p, cudaStream_t st);
int launch_bias_grad(const float* gy, float* gb, int C, int N, int H, int W, long long sN, long long sC, long long sH,
                     long long sW, int accumulate, cudaStream_t st);
int launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st);
int launch_cast_bf16_mirror(const float* src, void* dst, long long rows, int w, int pitch, const int* shifts,
                            int nshifts, cudaStream_t st);
int launch_stft(stft::Params& p, int n_signals, cudaStream_t st);

}  // namespace seldq

// Programmatic dependent launch (PDL): a kernel launched through launch_pdl may start while the kernel before it
// in the stream is still running its tail; it may touch that kernel's results (and overwrite buffers it reads)
// only after pdl_wait().  Every kernel here calls pdl_trigger() at entry, which lets its successor start as soon as
// all of this kernel's CTAs are resident, and pdl_wait() after a prologue that touches no activation memory
// (barrier init, TMEM allocation, tensor-map prefetch, op tables, weights: the kernels that write weights -- the
// optimiser and the packers -- never trigger early, so they have completed before any dependent starts).
// In a captured step this turns the kernel-to-kernel dependencies into programmatic graph edges.
// SELDQ_PDL=0 launches everything with plain stream serialisation.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace seldq {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SELDQ_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace seldq

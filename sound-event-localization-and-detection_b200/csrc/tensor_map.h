// Host-side construction of TMA tensor maps.  cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so that libseldq.so does not link against libcuda (it must load on a
// CPU-only host for the symbol-export test).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace seldq {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;  // immutable once resolved
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// swizzle: 0 none, 1 32B, 2 64B, 3 128B (CUtensorMapSwizzle numbering)
inline int encode_tensor_map(CUtensorMap* map, const void* gaddr, int elem_bytes, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes /* rank-1 entries, dims 1.. */, const uint32_t* box,
                             int swizzle) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(SELDQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  CUtensorMapDataType dt;
  switch (elem_bytes) {
    case 2: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    case 1: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    default: return fail(SELDQ_ERR_INVALID, "tensor map element size %d", elem_bytes);
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(gaddr), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SELDQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SELDQ_OK;
}

}  // namespace seldq

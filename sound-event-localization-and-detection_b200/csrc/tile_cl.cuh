// Shared pieces of the bandwidth-bound glue kernels (epilogue.cu, tcn_glue.cu): warp / block reductions and the
// [128 w][64 ch] bf16 shared-memory tile through which NCHW-ordered results leave as channels-last rows.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace seldq {
namespace epi {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two values; result valid in thread 0
__device__ __forceinline__ void block_sum2(float& a, float& b) {
  __shared__ float sa[32], sb[32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sa[warp] = a; sb[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? sa[lane] : 0.f;
    b = lane < nw ? sb[lane] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

__device__ __forceinline__ uint32_t hash_u32(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
// counter-based dropout mask: the same (seed, salt, element) always gives the same decision, so the backward
// pass needs no stored mask beyond the keep bit
// the per-element variant of the CNN tail: the (seed, salt) part is mixed once per thread (dropout_key), an element then
// costs two 32-bit multiplies (lowbias32 finaliser) instead of two 64-bit ones
__device__ __forceinline__ uint32_t dropout_key(unsigned long long seed, uint32_t salt) {
  return hash_u32((seed * 0x9E3779B97F4A7C15ULL) ^ ((unsigned long long)salt << 44));
}
__device__ __forceinline__ bool dropout_keep_fast(uint32_t key, long long elem, uint32_t thresh) {
  uint32_t x = (uint32_t)elem ^ key ^ ((uint32_t)((unsigned long long)elem >> 32) * 0x27d4eb2fu);
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x >= thresh;
}
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, uint32_t salt, long long elem, uint32_t thresh) {
  return hash_u32((seed * 0x9E3779B97F4A7C15ULL) ^ ((unsigned long long)salt << 44) ^ (unsigned long long)elem) >= thresh;
}

constexpr int kVecTileW = 128;
constexpr int kVecPitch = 128;          // bytes per w row of the shared tile: 64 ch x 2 B
// element (w, c) of the tile lives at w * 128 + ((2 c) ^ (((w >> 3) & 15) << 3)): the 16 lanes of a phase-1
// half-warp (same channel, w = 8 l + j) hit 16 different banks, and an aligned 16-byte group of 8 channels
// stays an aligned 16-byte group (its 8-byte halves swap when bit 3 of the key is set)
__device__ __forceinline__ uint32_t vec_tile_off(int wl, int byte_in_row) {
  return (uint32_t)(wl * kVecPitch + (byte_in_row ^ (((wl >> 3) & 15) << 3)));
}

template <int ROWS = kVecTileW>
__device__ __forceinline__ void vec_tile_store_cl(const uint8_t* tile, __nv_bfloat16* dst_cl, long long row0,
                                                  int w_base, int W, int Cp, int ct) {
  // thread -> (w = tid / 8 + 32 k, channels 8 (tid % 8) ... + 7)
  const int c0 = (threadIdx.x & 7) * 8;
  if (ct * 64 + c0 >= Cp) return;
#pragma unroll
  for (int k = 0; k < ROWS / 32; ++k) {
    const int wl = (threadIdx.x >> 3) + 32 * k;
    const int w = w_base + wl;
    if (w < W) {
      const uint32_t key = (uint32_t)((wl >> 3) & 15) << 3;
      uint4 v = *reinterpret_cast<const uint4*>(tile + wl * kVecPitch + ((c0 * 2) ^ (key & ~8u)));
      if (key & 8u) v = make_uint4(v.z, v.w, v.x, v.y);
      *reinterpret_cast<uint4*>(dst_cl + (row0 + w) * Cp + ct * 64 + c0) = v;
    }
  }
}


}  // namespace epi
}  // namespace seldq

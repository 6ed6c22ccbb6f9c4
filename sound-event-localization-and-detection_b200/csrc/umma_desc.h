// tcgen05 shared-memory / instruction descriptor bit layouts (PTX ISA "matrix descriptor" / "instruction descriptor"
// tables) as constexpr helpers without any device code, so that the host-side planner (conv_cl_plan.h) and the CPU
// emulation harness (tests/host_emul) can include them.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SELDQ_DESC_HD __host__ __device__
#else
#define SELDQ_DESC_HD
#endif

namespace seldq {
namespace ptx {

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
//   [32,46) stride-dimension byte offset >> 4     [46,48) version = 1 on sm_100
//   [49,52) base offset (0: operand tiles are aligned to their swizzle repeat)
//   [61,64) swizzle: 0 none, 1 128B(base 32B), 2 128B, 4 64B, 6 32B
constexpr uint32_t kSwizzleNone = 0, kSwizzle128B = 2, kSwizzle64B = 4, kSwizzle32B = 6;

SELDQ_DESC_HD constexpr uint64_t make_smem_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(swizzle & 7) << 61);
}

// Instruction descriptor (32 bit) for kind::f16 / kind::tf32:
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 f16, 1 bf16, 2 tf32)
//   [13] negate A  [14] negate B  [15] A major (0 K, 1 MN)  [16] B major  [17,23) N>>3  [24,29) M>>4
SELDQ_DESC_HD constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                      uint32_t b_mn_major, uint32_t neg_a, uint32_t neg_b) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (neg_a << 13) | (neg_b << 14) | (a_mn_major << 15) |
         (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}


}  // namespace ptx
}  // namespace seldq

// Bring-up probes for the tensor-core path (driven by tools/umma_probe.py on a B200):
//   * seldq_probe_tma_load : one TMA box -> shared memory -> raw dump (checks the swizzled layout)
//   * seldq_probe_umma     : host-provided shared-memory images + raw descriptors -> tcgen05.mma ->
//                            TMEM -> dump (checks descriptor semantics: major-ness, LBO/SBO, swizzle,
//                            negate bits) without recompiling
// They expose hardware behaviour only; nothing in the product path calls them, and they are built into their own
// tools library (tools/probe/libseldq_probe.so), not into libseldq.so.
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "seldq_probe.h"
#include "../../sound-event-localization-and-detection_b200/csrc/launch.h"
#include "../../sound-event-localization-and-detection_b200/csrc/tensor_map.h"
#include "../../sound-event-localization-and-detection_b200/csrc/umma_ptx.cuh"

namespace seldq {
namespace probe {

constexpr uint32_t kSmemBytes = 96 * 1024;

__global__ void __launch_bounds__(32) tma_load_kernel(const __grid_constant__ CUtensorMap map_param,
                                                     const CUtensorMap* map_global, int rank, int c0, int c1,
                                                     int c2, int c3, uint32_t box_bytes, uint32_t smem_offset,
                                                     uint32_t* out, uint32_t dump_words) {
  const CUtensorMap* mapp = map_global ? map_global : &map_param;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (uint32_t i = threadIdx.x; i < kSmemBytes / 4; i += 32) w[i] = 0x7f7f7f7fu;  // sentinel
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  __syncwarp();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, box_bytes);
    if (rank == 2) ptx::tma_load_2d(smem + smem_offset, mapp, &bar, c0, c1);
    else ptx::tma_load_4d(smem + smem_offset, mapp, &bar, c0, c1, c2, c3);
  }
  ptx::mbar_wait(&bar, 0);
  __syncwarp();
  for (uint32_t i = threadIdx.x; i < dump_words; i += 32) out[i] = w[i];
}

__global__ void __launch_bounds__(128) umma_kernel(const uint4* a_image, uint32_t a_bytes, const uint4* b_image,
                                                  uint32_t b_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  int n_mma, uint32_t a_step, uint32_t b_step, int n_cols,
                                                  uint32_t tmem_cols, float* out, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + ((a_bytes + 1023) & ~1023u);
  for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(a_smem)[i] = a_image[i];
  for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(b_smem)[i] = b_image[i];
  ptx::fence_proxy_async();
  if (threadIdx.x == 32) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc(&tmem_base, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (threadIdx.x == 0) {
    const uint64_t ad = a_desc + (uint64_t)(ptx::smem_u32(a_smem) >> 4);
    const uint64_t bd = b_desc + (uint64_t)(ptx::smem_u32(b_smem) >> 4);
    if (!(flags & 1))
      for (int i = 0; i < n_mma; ++i)
        ptx::umma_f16(tbase, ad + (uint64_t)i * a_step, bd + (uint64_t)i * b_step, idesc, i > 0 ? 1u : 0u);
    if (flags & 2) ptx::mbar_arrive(&bar);
    else ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < n_cols; c += 8) {
    uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!(flags & 4)) {
      ptx::tmem_ld8(tbase + ((warp * 32u) << 16) + (uint32_t)c, r);
      ptx::tmem_ld_wait();
    }
    for (int j = 0; j < 8; ++j)
      if (c + j < n_cols) out[(warp * 32 + lane) * n_cols + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, tmem_cols);
}


// tcgen05.mma issue-rate probe: one thread issues n_mma back-to-back MMAs (M = 128, K = 16, bf16) with the
// product kernels' operand layouts (A: K-major 128B-swizzled [128 x 64] stages, one 16-channel slab per
// MMA; B: K-major unswizzled [N x 16] tiles) cycling over `d_cycle` accumulator column ranges, and reports
// SM cycles from the first issue to the completion of the last.  mode bit 0: A is MN-major ([64 t x 128])
__global__ void __launch_bounds__(128) umma_rate_kernel(uint32_t n, int n_mma, int d_cycle, int mode,
                                                       long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (uint32_t i = threadIdx.x; i < (128u * 1024u) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  if (threadIdx.x == 32) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc(&tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = tmem_base;
  if ((mode & 8) && threadIdx.x < 32) {
    // lane-parallel preparation: lane l owns MMA l of every stage (nl = d_cycle >> 8 lanes per stage); the
    // descriptors are computed by all lanes at once and the issue is a per-lane tcgen05.mma
    const int nl = d_cycle >> 8, dcyc = d_cycle & 255;
    const int lane = threadIdx.x;
    const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;
    const uint64_t a_hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
    const uint64_t b_hi = ptx::make_smem_desc_hi(n * 16u, 128, ptx::kSwizzleNone);
    const uint32_t idesc = ptx::make_idesc_bf16(128, n, 0, 0, 0, (uint32_t)(lane & 1));
    const uint32_t a_off = (uint32_t)(lane & 3) * 32u, b_off = (uint32_t)((lane >> 2) & 7) * (n * 32u);
    const uint32_t d = tbase + (uint32_t)((lane >> 2) % dcyc) * n;
    const long long t0 = clock64();
    uint32_t stage = 0;
    for (int i = 0; i < n_mma; i += nl) {
      const uint64_t ad = ptx::smem_desc(a_hi, a0 + stage * 16384u + a_off);
      const uint64_t bd = ptx::smem_desc(b_hi, b0 + b_off + stage * 16u);
      if (lane < nl) ptx::umma_f16(d, ad, bd, idesc, 1u);
      __syncwarp();
      stage = (stage + 1) & 3;
    }
    const long long t1 = clock64();
    if (lane == 0) {
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
    }
  } else if ((mode & 16) && threadIdx.x == 0) {
    // one thread, ready-made entries {a offset, b offset, idesc, d offset} fetched from shared memory
    __shared__ uint4 ent[32];
    for (int l = 0; l < 32; ++l)
      ent[l] = make_uint4((uint32_t)(l & 3) * 2u, ((uint32_t)((l >> 2) & 7) * (n * 32u)) >> 4,
                          ptx::make_idesc_bf16(128, n, 0, 0, 0, (uint32_t)(l & 1)), (uint32_t)((l >> 2) % (d_cycle & 255)) * n);
    const int nl = d_cycle >> 8;
    const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;
    const uint64_t a_hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
    const uint64_t b_hi = ptx::make_smem_desc_hi(n * 16u, 128, ptx::kSwizzleNone);
    const long long t0 = clock64();
    uint32_t stage = 0;
    for (int i = 0; i < n_mma; i += nl) {
      const uint32_t a16 = (a0 + stage * 16384u) >> 4, b16 = (b0 + stage * 16u) >> 4;
#pragma unroll 4
      for (int l = 0; l < nl; ++l) {
        const uint4 e = ent[l];
        ptx::umma_f16(tbase + e.w, a_hi | (uint64_t)(a16 + e.x), b_hi | (uint64_t)(b16 + e.y), e.z, 1u);
      }
      stage = (stage + 1) & 3;
    }
    const long long t1 = clock64();
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  } else if (!(mode & 24) && threadIdx.x == 0) {
    const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;
    const uint64_t a_hi = (mode & 1) ? ptx::make_smem_desc_hi(8192, 1024, ptx::kSwizzle128B)
                                     : ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
    const uint64_t b_hi = ptx::make_smem_desc_hi(n * 16u, 128, ptx::kSwizzleNone);
    const uint32_t idesc = ptx::make_idesc_bf16(128, n, (mode & 1) ? 1 : 0, 0, 0, 0);
    const long long t0 = clock64();
    int dc = 0;
    if (mode & 2) {
      // descriptors precomputed, 8 MMAs per iteration, no address arithmetic in the loop
      uint64_t ad[4], bd[2];
      for (int k = 0; k < 4; ++k) ad[k] = ptx::smem_desc(a_hi, a0 + (uint32_t)k * ((mode & 1) ? 2048u : 32u));
      for (int k = 0; k < 2; ++k) bd[k] = ptx::smem_desc(b_hi, b0 + (uint32_t)k * (n * 32u));
      const uint32_t d1 = tbase + ((d_cycle > 1) ? n : 0u);
      for (int i = 0; i < n_mma; i += 8) {
        ptx::umma_f16(tbase, ad[0], bd[0], idesc, 1u);
        ptx::umma_f16(d1, ad[0], bd[1], idesc, 1u);
        ptx::umma_f16(tbase, ad[1], bd[0], idesc, 1u);
        ptx::umma_f16(d1, ad[1], bd[1], idesc, 1u);
        ptx::umma_f16(tbase, ad[2], bd[0], idesc, 1u);
        ptx::umma_f16(d1, ad[2], bd[1], idesc, 1u);
        ptx::umma_f16(tbase, ad[3], bd[0], idesc, 1u);
        ptx::umma_f16(d1, ad[3], bd[1], idesc, 1u);
      }
    } else
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t a_addr = a0 + (uint32_t)((i >> 2) & 3) * 16384u + (uint32_t)(i & 3) * ((mode & 1) ? 2048u : 32u);
      const uint32_t b_addr = b0 + (uint32_t)(i & 7) * (n * 32u);
      ptx::umma_f16(tbase + (uint32_t)dc * n, ptx::smem_desc(a_hi, a_addr), ptx::smem_desc(b_hi, b_addr), idesc, 1u);
      if (++dc == d_cycle) dc = 0;
    }
    const long long t1 = clock64();
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

// A-from-tensor-memory probe.  The product MMAs are small (N = 32 | 48 per Hamilton block) and in the SS form every
// one re-reads its 4 KB A slab from shared memory (~45 cycles, r1b_umma_rate.txt).  Here the slab is copied to
// tensor memory once (tcgen05.cp.128x256b) and G MMAs read it from there.
//   mode 0  correctness: K = 64 (4 slabs) through the SS form and through cp + TS form, bitwise comparison of the
//           accumulators; out[0] = mismatching elements, out[1] = non-zero elements of the SS result
//   mode 1  TS MMAs only (A resident), mode 2  one cp + G TS MMAs per slab, mode 3  cp only,
//   mode 4  G SS MMAs per slab (same loop, the baseline);  out[2b], out[2b+1] = issue, issue + drain cycles
template <int G, int MODE>
__device__ __forceinline__ void ts_slab_loop(uint32_t n, int n_slabs, int nbuf, uint32_t tD, uint32_t tA, uint32_t a0,
                                             uint32_t b0, int d_cycle) {
  const uint64_t a_hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
  const uint64_t b_hi = ptx::make_smem_desc_hi(n * 16u, 128, ptx::kSwizzleNone);
  uint64_t bd[G];
  uint32_t id[G], dd[G];
#pragma unroll
  for (int j = 0; j < G; ++j) {
    bd[j] = ptx::smem_desc(b_hi, b0 + (uint32_t)(j & 7) * (n * 32u));
    id[j] = ptx::make_idesc_bf16(128, n, 0, 0, 0, (uint32_t)(j & 1));
    dd[j] = tD + (uint32_t)(j % d_cycle) * n;
  }
  const uint64_t ad0 = ptx::smem_desc(a_hi, a0);
  int buf = 0;
  for (int s = 0; s < n_slabs; ++s) {
    const uint64_t ad = ad0 + (uint64_t)(((s >> 2) & 3) * 1024 + (s & 3) * 2);   // stage * 16 KB + slab * 32 B, >> 4
    const uint32_t ab = tA + (uint32_t)buf * 8u;
    if (MODE == 2 || MODE == 3) ptx::tmem_cp_128x256b(ab, ad);
#pragma unroll
    for (int j = 0; j < G; ++j) {
      if (MODE == 1 || MODE == 2) ptx::umma_f16_ts(dd[j], ab, bd[j], id[j], 1u);
      if (MODE == 4) ptx::umma_f16(dd[j], ad, bd[j], id[j], 1u);
    }
    if (++buf == nbuf) buf = 0;
  }
}

__global__ void __launch_bounds__(128) umma_ts_kernel(uint32_t n, int n_slabs, int g, int mode, int nbuf,
                                                     long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  // small integers as bf16 (exact in fp32 accumulation, any placement is a valid operand)
  for (uint32_t i = threadIdx.x; i < (128u * 1024u) / 2; i += 128) {
    const uint32_t h = (i * 2654435761u) >> 13;
    const float v = mode == 0 ? (float)((int)(h % 7u) - 3) : 0.f;
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16_rn(v);
  }
  ptx::fence_proxy_async();
  if (threadIdx.x == 32) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc(&tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = tmem_base;
  const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;
  const uint32_t tA = tbase + 448;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      const uint64_t a_hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
      const uint64_t b_hi = ptx::make_smem_desc_hi(n * 16u, 128, ptx::kSwizzleNone);
      for (int k = 0; k < 4 * n_slabs; ++k) {
        const uint64_t ad = ptx::smem_desc(a_hi, a0 + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
        const uint64_t bd = ptx::smem_desc(b_hi, b0 + (uint32_t)(k & 7) * (n * 32u));
        const uint32_t idesc = ptx::make_idesc_bf16(128, n, 0, 0, 0, (uint32_t)(k & 1));
        ptx::umma_f16(tbase, ad, bd, idesc, k > 0 ? 1u : 0u);
      }
      for (int k = 0; k < 4 * n_slabs; ++k) {
        const uint64_t ad = ptx::smem_desc(a_hi, a0 + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
        const uint64_t bd = ptx::smem_desc(b_hi, b0 + (uint32_t)(k & 7) * (n * 32u));
        const uint32_t idesc = ptx::make_idesc_bf16(128, n, 0, 0, 0, (uint32_t)(k & 1));
        const uint32_t ab = tA + (uint32_t)(k % nbuf) * 8u;
        ptx::tmem_cp_128x256b(ab, ad);
        // g MMAs read the slab: the first computes, the others re-add and subtract it (net zero)
        ptx::umma_f16_ts(tbase + 256, ab, bd, idesc, k > 0 ? 1u : 0u);
        for (int j = 1; j < g; ++j) ptx::umma_f16_ts(tbase + 256, ab, bd, idesc ^ ((uint32_t)(j & 1) << 14) ^ (1u << 14), 1u);
      }
      ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    const uint32_t warp = threadIdx.x >> 5;
    unsigned long long bad = 0, nz = 0;
    for (uint32_t c = 0; c < n; c += 8) {
      uint32_t r0[8], r1[8];
      ptx::tmem_ld8(tbase + ((warp * 32u) << 16) + c, r0);
      ptx::tmem_ld8(tbase + ((warp * 32u) << 16) + 256 + c, r1);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 8; ++j) { bad += r0[j] != r1[j]; nz += r0[j] != 0; }
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(out), bad);
    atomicAdd(reinterpret_cast<unsigned long long*>(out) + 1, nz);
  } else if (threadIdx.x == 0) {
    if (mode == 1 || mode == 2) ptx::tmem_cp_128x256b(tA, ptx::smem_desc(ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B), a0));
    const int dcyc = (int)(256u / n) < 4 ? (int)(256u / n) : 4;
    const long long t0 = clock64();
#define SELDQ_TS_CASE(GG, MM) if (g == GG && mode == MM) ts_slab_loop<GG, MM>(n, n_slabs, nbuf, tbase, tA, a0, b0, dcyc);
#define SELDQ_TS_G(MM) SELDQ_TS_CASE(1, MM) SELDQ_TS_CASE(2, MM) SELDQ_TS_CASE(3, MM) SELDQ_TS_CASE(4, MM) SELDQ_TS_CASE(6, MM) SELDQ_TS_CASE(8, MM)
    SELDQ_TS_G(1) SELDQ_TS_G(2) SELDQ_TS_G(3) SELDQ_TS_G(4)
#undef SELDQ_TS_G
#undef SELDQ_TS_CASE
    const long long t1 = clock64();
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

}  // namespace probe
}  // namespace seldq

using namespace seldq;

extern "C" int seldq_probe_tensor_map(void* host_map_128B, const void* gaddr, int32_t elem_bytes, int32_t rank,
                                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                      int32_t swizzle) {
  if (!host_map_128B || rank < 1 || rank > 5) return fail(SELDQ_ERR_INVALID, "bad tensor-map probe arguments");
  alignas(64) CUtensorMap m;
  const int rc = encode_tensor_map(&m, gaddr, elem_bytes, rank, dims, strides_bytes, box, swizzle);
  if (rc) return rc;
  memcpy(host_map_128B, &m, sizeof(m));
  return SELDQ_OK;
}

extern "C" int seldq_probe_tma_load(const void* host_map_128B, const void* dev_map_128B, int32_t rank,
                                    const int32_t* coords, uint32_t box_bytes, uint32_t smem_offset,
                                    void* out_smem_dump, uint32_t dump_bytes, void* stream) {
  if (rank != 2 && rank != 4) return fail(SELDQ_ERR_INVALID, "tma probe supports rank 2 and 4");
  if (smem_offset + box_bytes > probe::kSmemBytes || dump_bytes > probe::kSmemBytes)
    return fail(SELDQ_ERR_INVALID, "tma probe exceeds its %u byte window", probe::kSmemBytes);
  alignas(64) CUtensorMap m;
  memcpy(&m, host_map_128B, sizeof(m));
  cudaError_t e = cudaFuncSetAttribute(probe::tma_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)probe::kSmemBytes);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "probe smem opt-in: %s", cudaGetErrorString(e));
  probe::tma_load_kernel<<<1, 32, probe::kSmemBytes, (cudaStream_t)stream>>>(
      m, (const CUtensorMap*)dev_map_128B, rank, coords[0], coords[1], rank > 2 ? coords[2] : 0,
      rank > 3 ? coords[3] : 0, box_bytes, smem_offset,
      (uint32_t*)out_smem_dump, dump_bytes / 4);
  return check_launch("probe::tma_load_kernel");
}

extern "C" int seldq_probe_umma(const void* a_image, uint32_t a_bytes, const void* b_image, uint32_t b_bytes,
                                uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int32_t n_mma, uint32_t a_desc_step,
                                uint32_t b_desc_step, int32_t n_cols, float* out_128xN, void* stream) {
  if ((a_bytes & 15) || (b_bytes & 15) || n_cols < 8 || n_cols > 512)
    return fail(SELDQ_ERR_INVALID, "bad umma probe arguments");
  // generous window: a wrong descriptor hypothesis should read garbage, not fault
  const uint32_t smem = 160 * 1024;
  if (((a_bytes + 1023) & ~1023u) + b_bytes > smem) return fail(SELDQ_ERR_INVALID, "umma probe images too large");
  uint32_t cols = 32;
  while (cols < (uint32_t)n_cols) cols <<= 1;
  cudaError_t e = cudaFuncSetAttribute(probe::umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "probe smem opt-in: %s", cudaGetErrorString(e));
  probe::umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint4*)a_image, a_bytes, (const uint4*)b_image,
                                                            b_bytes, a_desc, b_desc, idesc, n_mma, a_desc_step,
                                                            b_desc_step, n_cols, cols, out_128xN,
                                                            getenv("SELDQ_DEBUG") ? atoi(getenv("SELDQ_DEBUG")) : 0);
  return check_launch("probe::umma_kernel");
}

// out: 2 x blocks int64 (issue cycles, issue + drain cycles)
extern "C" int seldq_probe_umma_rate(uint32_t n, int32_t n_mma, int32_t d_cycle, int32_t mode, int32_t blocks,
                                     void* out, void* stream) {
  if (n < 8 || n > 256 || (n & 7) || (d_cycle & 255) < 1 || (uint32_t)(d_cycle & 255) * n > 512 || blocks < 1 ||
      ((mode & 24) && ((d_cycle >> 8) < 1 || (d_cycle >> 8) > 32 || n_mma % (d_cycle >> 8))))
    return fail(SELDQ_ERR_INVALID, "bad umma rate probe arguments");
  const uint32_t smem = 128 * 1024 + 4096;
  cudaError_t e = cudaFuncSetAttribute(probe::umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "probe smem opt-in: %s", cudaGetErrorString(e));
  probe::umma_rate_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(n, n_mma, d_cycle, mode, (long long*)out);
  return check_launch("probe::umma_rate_kernel");
}

// A-from-tensor-memory probe (see probe::umma_ts_kernel); out: 2 x blocks int64
extern "C" int seldq_probe_umma_ts(uint32_t n, int32_t n_slabs, int32_t g, int32_t mode, int32_t nbuf, int32_t blocks,
                                   void* out, void* stream) {
  if (n < 8 || n > 256 || (n & 7) || n_slabs < 1 || mode < 0 || mode > 4 || nbuf < 1 || nbuf > 8 || blocks < 1 ||
      !(g == 1 || g == 2 || g == 3 || g == 4 || g == 6 || g == 8) || (mode == 0 && n_slabs > 4))
    return fail(SELDQ_ERR_INVALID, "bad umma ts probe arguments");
  const uint32_t smem = 128 * 1024 + 4096;
  cudaError_t e = cudaFuncSetAttribute(probe::umma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "probe smem opt-in: %s", cudaGetErrorString(e));
  probe::umma_ts_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(n, n_slabs, g, mode, nbuf, (long long*)out);
  return check_launch("probe::umma_ts_kernel");
}

extern "C" const char* seldq_probe_last_error(void) { return seldq::error_buffer(); }

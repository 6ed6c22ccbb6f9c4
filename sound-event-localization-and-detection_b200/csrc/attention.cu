// Fused multi-head self-attention on tcgen05 tensor cores (MultiHeadAttention, model.py:12-51).
//
//     energy = q k^T ;  attention = softmax(energy / sqrt(d)) over the keys ;  out = attention v        (model.py:40-48)
//
// The reference materialises the (N, heads, S, S) energy and attention tensors (185 MB per sample at S = 2400) and so
// did the PyTorch ops this replaces (two batched GEMMs around a softmax: ~0.65 ms of a 4.3 ms step).  Here the S x S
// tensors never leave the SM: one CTA owns 128 rows (queries, or keys in the dK / dV pass) of one (sample, head) and
// sweeps the other axis in chunks of 128:
//
//   forward   T = Q_tile K_chunk^T (tcgen05.mma, accumulator in TMEM) -> 8 softmax warps read T (tcgen05.ld), write
//             P = exp2(T - rowmax) as the bf16 A operand of the next MMA into shared memory (128B-swizzled, K-major)
//             -> O += P V_chunk.  Two sweeps: the first only finds the row maxima (three K = 16 MMAs per chunk: the
//             extra QK^T costs less than an online-softmax rescale of O would), the second needs no rescaling.
//   dK, dV    rows = keys:    T1 = K_tile Q_chunk^T (= S^T),  T2 = V_tile dO_chunk^T (= dP^T)
//             P^T = exp2(T1 - lse_i),  dS^T = P^T (T2 - delta_i)  ->  dV += P^T dO_chunk,  dK += dS^T Q_chunk
//   dQ        rows = queries: T1 = Q_tile K_chunk^T,  T2 = dO_tile V_chunk^T,  dS = P (T2 - delta_i)  ->  dQ += dS K_chunk
//   (P is recomputed from the saved log-sum-exp instead of being stored: no S x S tensor in HBM in either direction.)
//
// Every MMA operand is K-major with the 128-byte swizzle, the one layout this library already feeds the tensor cores
// (conv_cl.cu): row-major [S][64] bf16 copies of q (pre-scaled by log2(e) / sqrt(d)), k, v, dO serve as the 128-row
// operands of the score products, and the tensors' own [d][S] order serves as the [d x 128] B operand of the output
// products; attn_stage_kernel writes both forms once per call.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread; owns the TMEM allocation), warps 2..9 =
// softmax / epilogue (warp w reads TMEM lane quarter w % 4; the two warps of a quarter split the chunk's columns).
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "attention.h"
#include "launch.h"
#include "tensor_map.h"
#include "umma_ptx.cuh"

namespace seldq {
namespace attn {

constexpr int kThreads = 320;
constexpr int kRows = 128;           // rows per CTA = TMEM lanes
constexpr int kChunk = 128;          // columns per chunk
constexpr int kMaxStages = 3;
constexpr uint32_t kTileBytes = 128 * 128;       // [128 rows x 64 bf16]: one K-major operand tile
constexpr uint32_t kPBytes = 2 * kTileBytes;     // [128 x 128] bf16 P / dS tile = two K tiles

enum { MODE_FWD = 0, MODE_DKV = 1, MODE_DQ = 2 };

struct Params {
  int S, d, H, BH, E;
  int n_chunks;
  int nstages;
  uint32_t stage_bytes, b3_tile_bytes;     // b3_tile_bytes = d * 128: one [d x 64] K tile of a transposed operand
  float* out;                              // forward: (N, S, E) fp32
  float* lse;                              // [BH][S], log2 domain (forward writes, backward reads)
  const float* delta;                      // [BH][S] = sum_d dO * O (backward)
  float* o1;                               // dV (DKV) | dQ (DQ), (N, E, S) fp32
  float* o2;                               // dK (DKV)
  float scale1, scale2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
// 8 consecutive values of one row of a [128 x 128] bf16 tile -> its place in the K-major 128B-swizzled layout:
// K tile (col / 64) of 16 KB, row r at r * 128 B, 16-byte chunk ch stored at position ch ^ (r & 7)
__device__ __forceinline__ void store_p8(uint8_t* tile, int row, int col, const float (&v)[8]) {
  const int kt = col >> 6, ch = (col & 63) >> 3;
  const uint4 q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                             pack_bf16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(tile + (size_t)kt * kTileBytes + (size_t)row * 128 + ((ch ^ (row & 7)) << 4)) = q;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
            const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmB4,
            const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t a_full, full_bar[kMaxStages], empty_bar[kMaxStages], t_full[2], t_empty[2], p_full[2],
      p_empty[2], o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][kRows];
  __shared__ float col_lse[2][kChunk], col_del[2][kChunk];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int bh = blockIdx.y, r0 = blockIdx.x * kRows;
  constexpr bool kBwd = MODE != MODE_FWD;

  // shared memory: [A1][A2 (bwd)][stages: B1, B2 (bwd), B3, B4 (DKV)][P tiles]
  uint8_t* sA1 = smem;
  uint8_t* sA2 = smem + kTileBytes;
  uint8_t* stages = smem + (kBwd ? 2 : 1) * kTileBytes;
  uint8_t* sP = stages + (size_t)p.nstages * p.stage_bytes;
  const uint32_t offB2 = kTileBytes, offB3 = (kBwd ? 2u : 1u) * kTileBytes, offB4 = offB3 + 2u * p.b3_tile_bytes;

  if (threadIdx.x == 0) {
    ptx::mbar_init(&a_full, 1);
    for (int i = 0; i < p.nstages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 8);
      ptx::mbar_init(&p_full[i], 8); ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::mbar_init(&o_full, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA1);
    ptx::prefetch_tensormap(&tmB1);
    ptx::prefetch_tensormap(&tmB3);
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // TMEM columns: forward T[0] = 0, T[1] = 128, O = 256; backward T1 = 0, T2 = 128, O1 = 256, O2 = 320
  const uint32_t colO1 = 256, colO2 = 320;

  if (warp == 0) {
    // ===== TMA producer ===================================================================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&a_full, kBwd ? 2 * kTileBytes : kTileBytes);
      tma_load_3d(sA1, &tmA1, &a_full, 0, r0, bh);
      if (kBwd) tma_load_3d(sA2, &tmA2, &a_full, 0, r0, bh);
      uint32_t slot = 0, parity = 0;
      if (MODE == MODE_FWD) {
        for (int c = 0; c < p.n_chunks; ++c) {             // first sweep: keys only
          ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[slot], kTileBytes);
          tma_load_3d(stages + (size_t)slot * p.stage_bytes, &tmB1, &full_bar[slot], 0, c * kChunk, bh);
          if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
        }
      }
      const uint32_t tx = (kBwd ? 2u : 1u) * kTileBytes + (MODE == MODE_DKV ? 4u : 2u) * p.b3_tile_bytes;
      for (int c = 0; c < p.n_chunks; ++c) {
        ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[slot], tx);
        uint8_t* st = stages + (size_t)slot * p.stage_bytes;
        tma_load_3d(st, &tmB1, &full_bar[slot], 0, c * kChunk, bh);
        if (kBwd) tma_load_3d(st + offB2, &tmB2, &full_bar[slot], 0, c * kChunk, bh);
        for (int kt = 0; kt < 2; ++kt) {
          tma_load_3d(st + offB3 + kt * p.b3_tile_bytes, &tmB3, &full_bar[slot], c * kChunk + kt * 64, 0, bh);
          if (MODE == MODE_DKV)
            tma_load_3d(st + offB4 + kt * p.b3_tile_bytes, &tmB4, &full_bar[slot], c * kChunk + kt * 64, 0, bh);
        }
        if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====================================================================================
    if (ptx::elect_one()) {
      const uint64_t kdesc = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);     // K-major, 8-row groups 1 KB apart
      const uint32_t idesc_s = ptx::make_idesc_bf16(128, kChunk, 0, 0, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_bf16(128, (uint32_t)p.d, 0, 0, 0, 0);
      const int ksl = p.d >> 4;                           // K = 16 slabs of a score product
      // T[128 x 128] = A[128 x d] B[128 x d]^T
      auto score = [&](uint32_t d_col, const uint8_t* a, const uint8_t* b) {
        const uint32_t aa = ptx::smem_u32(a), ba = ptx::smem_u32(b);
        for (int k = 0; k < ksl; ++k)
          ptx::umma_f16(tmem_base + d_col, ptx::smem_desc(kdesc, aa + 32u * k), ptx::smem_desc(kdesc, ba + 32u * k), idesc_s,
                        k > 0 ? 1u : 0u);
      };
      // O[128 x d] (+)= P[128 x 128] B[d x 128]^T, both as two K tiles of 64
      auto outp = [&](uint32_t d_col, const uint8_t* pt, const uint8_t* b, bool first) {
        const uint32_t pa = ptx::smem_u32(pt), ba = ptx::smem_u32(b);
        for (int kt = 0; kt < 2; ++kt)
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_f16(tmem_base + d_col, ptx::smem_desc(kdesc, pa + kt * kTileBytes + 32u * ks),
                          ptx::smem_desc(kdesc, ba + kt * p.b3_tile_bytes + 32u * ks), idesc_o,
                          (first && kt == 0 && ks == 0) ? 0u : 1u);
      };
      ptx::mbar_wait(&a_full, 0);
      ptx::tc_fence_after();
      uint32_t slot = 0, parity = 0, it = 0;
      if (MODE == MODE_FWD) {
        for (int c = 0; c < p.n_chunks; ++c, ++it) {       // first sweep: scores only
          const uint32_t tb = it & 1, u = it >> 1;
          ptx::mbar_wait(&full_bar[slot], parity);
          ptx::mbar_wait(&t_empty[tb], (u & 1) ^ 1);
          ptx::tc_fence_after();
          score(tb * kChunk, sA1, stages + (size_t)slot * p.stage_bytes);
          ptx::umma_commit(&t_full[tb]);
          ptx::umma_commit(&empty_bar[slot]);
          if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
        }
      }
      uint32_t prev_slot = 0;
      for (int c = 0; c <= p.n_chunks; ++c) {
        if (c < p.n_chunks) {
          const uint32_t tb = kBwd ? 0u : (it & 1), u = kBwd ? it : (it >> 1);
          ptx::mbar_wait(&full_bar[slot], parity);
          ptx::mbar_wait(&t_empty[tb], (u & 1) ^ 1);
          ptx::tc_fence_after();
          const uint8_t* st = stages + (size_t)slot * p.stage_bytes;
          score(kBwd ? 0u : tb * kChunk, sA1, st);
          if (kBwd) score(kChunk, sA2, st + offB2);
          ptx::umma_commit(&t_full[tb]);
          ++it;
        }
        if (c > 0) {
          const int cc = c - 1;
          const uint32_t pb = kBwd ? 0u : (cc & 1), u = kBwd ? cc : (cc >> 1);
          ptx::mbar_wait(&p_full[pb], u & 1);
          ptx::tc_fence_after();
          const uint8_t* st = stages + (size_t)prev_slot * p.stage_bytes;
          if (MODE == MODE_FWD) outp(colO1, sP + pb * kPBytes, st + offB3, cc == 0);
          if (MODE == MODE_DKV) {
            outp(colO1, sP, st + offB3, cc == 0);                  // dV += P^T dO
            outp(colO2, sP + kPBytes, st + offB4, cc == 0);        // dK += dS^T Q
          }
          if (MODE == MODE_DQ) outp(colO1, sP, st + offB3, cc == 0);            // dQ += dS K
          ptx::umma_commit(&p_empty[pb]);
          ptx::umma_commit(&empty_bar[prev_slot]);
        }
        if (c < p.n_chunks) {
          prev_slot = slot;
          if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
        }
      }
      ptx::umma_commit(&o_full);
    }
  } else {
    // ===== softmax / epilogue warps =======================================================================
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const int row = q * 32 + lane, col0 = hf * 64;
    const int ct = (int)threadIdx.x - 64;                  // 0..255 among these warps
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int gr = r0 + row;                               // global row (query, or key in the dK / dV pass)
    uint32_t it = 0;
    float m = -INFINITY, l = 0.f;
    float row_lse = 0.f, row_del = 0.f;
    if (MODE == MODE_DQ && gr < p.S) {
      row_lse = p.lse[(size_t)bh * p.S + gr];
      row_del = p.delta[(size_t)bh * p.S + gr];
    }
    if (MODE == MODE_FWD) {
      for (int c = 0; c < p.n_chunks; ++c, ++it) {         // first sweep: row maxima
        const uint32_t tb = it & 1, u = it >> 1;
        ptx::mbar_wait(&t_full[tb], u & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          ptx::tmem_ld32(t_lane + tb * kChunk + col0 + 32 * g, v);
          ptx::tmem_ld_wait();
          const int key0 = c * kChunk + col0 + 32 * g;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j < p.S) m = fmaxf(m, __uint_as_float(v[j]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[tb]);
      }
      red[hf][row] = m;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      m = fmaxf(red[0][row], red[1][row]);
    }
    for (int c = 0; c < p.n_chunks; ++c, ++it) {
      const uint32_t tb = kBwd ? 0u : (it & 1), tu = kBwd ? (uint32_t)c : (it >> 1);
      const uint32_t pb = kBwd ? 0u : (c & 1), pu = kBwd ? (uint32_t)c : (uint32_t)(c >> 1);
      const int cb = c & 1;
      if (MODE == MODE_DKV) {
        // per-column statistics of the chunk (queries): log-sum-exp and delta; columns beyond S give P = 0
        const int i = c * kChunk + (ct & 127);
        if (ct < 128) col_lse[cb][ct] = i < p.S ? p.lse[(size_t)bh * p.S + i] : INFINITY;
        else col_del[cb][ct - 128] = i < p.S ? p.delta[(size_t)bh * p.S + i] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      ptx::mbar_wait(&t_full[tb], tu & 1);
      ptx::mbar_wait(&p_empty[pb], (pu & 1) ^ 1);
      ptx::tc_fence_after();
      uint8_t* tileP = sP + (MODE == MODE_FWD ? pb * kPBytes : 0u);
      uint8_t* tileS = sP + (MODE == MODE_DKV ? kPBytes : 0u);           // dS tile (backward)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int cbase = col0 + 32 * g;
        uint32_t v[32];
        ptx::tmem_ld32(t_lane + (kBwd ? 0u : tb * kChunk) + cbase, v);
        if (!kBwd) {
          ptx::tmem_ld_wait();
          const int key0 = c * kChunk + cbase;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float pv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float e = key0 + j8 * 8 + j < p.S ? ex2(__uint_as_float(v[j8 * 8 + j]) - m) : 0.f;
              pv[j] = e;
              l += e;
            }
            store_p8(tileP, row, cbase + j8 * 8, pv);
          }
        } else {
          uint32_t w[32];
          ptx::tmem_ld32(t_lane + kChunk + cbase, w);
          ptx::tmem_ld_wait();
          const int key0 = c * kChunk + cbase;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float pv[8], dv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int cj = cbase + j8 * 8 + j;
              float e;
              if (MODE == MODE_DKV) {
                e = ex2(__uint_as_float(v[j8 * 8 + j]) - col_lse[cb][cj]);
                dv[j] = e * (__uint_as_float(w[j8 * 8 + j]) - col_del[cb][cj]);
              } else {
                e = key0 + j8 * 8 + j < p.S ? ex2(__uint_as_float(v[j8 * 8 + j]) - row_lse) : 0.f;
                dv[j] = e * (__uint_as_float(w[j8 * 8 + j]) - row_del);
              }
              pv[j] = e;
            }
            if (MODE == MODE_DKV) store_p8(tileP, row, cbase + j8 * 8, pv);
            store_p8(tileS, row, cbase + j8 * 8, dv);
          }
        }
      }
      ptx::fence_proxy_async();           // the tiles just written are read by the tensor core (async proxy)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&t_empty[tb]);
        ptx::mbar_arrive(&p_full[pb]);
      }
    }
    if (MODE == MODE_FWD) {
      asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone has read red[] of the first sweep
      red[hf][row] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      l = red[0][row] + red[1][row];
    }
    // ---- epilogue: the [128 x d] accumulators -------------------------------------------------------------
    ptx::mbar_wait(&o_full, 0);
    ptx::tc_fence_after();
    const int half = p.d >> 1;                              // this thread's columns [hf * d/2, + d/2)
    const int b = bh / p.H, h = bh - b * p.H;
    for (int o = 0; o < (MODE == MODE_DKV ? 2 : 1); ++o) {
      const uint32_t ocol = (o == 0 ? colO1 : colO2) + (uint32_t)(hf * half);
      for (int j0 = 0; j0 < half; j0 += 8) {
        uint32_t v[8];
        ptx::tmem_ld8(t_lane + ocol + j0, v);
        ptx::tmem_ld_wait();
        if (gr >= p.S) continue;
        if (MODE == MODE_FWD) {
          const float inv = 1.f / l;
          float* dst = p.out + ((size_t)b * p.S + gr) * p.E + h * p.d + hf * half + j0;
          *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv,
                                                        __uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv,
                                                            __uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
        } else {
          float* dst = (o == 0 ? p.o1 : p.o2) + ((size_t)b * p.E + h * p.d + hf * half + j0) * p.S + gr;
          const float sc = o == 0 ? p.scale1 : p.scale2;
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[(size_t)j * p.S] = __uint_as_float(v[j]) * sc;
        }
      }
    }
    if (MODE == MODE_FWD && hf == 0 && gr < p.S) p.lse[(size_t)bh * p.S + gr] = m + log2f(l);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- operand staging --------------------------------------------------------------------------------------------
// src (layout 0: (N, E, S) = the 1x1 projections' output; layout 1: (N, S, E) = the gradient of the attention output)
//   -> rm [BH][S][64] bf16, zero padded beyond d   (row-major: 128-row operand of the score products)
//   -> tr [BH][d][S]  bf16                         (the tensor's own order: [d x 128] operand of the output products)
// layout 1 also writes delta[bh][s] = sum_d src * o (o in the same layout).
__global__ void __launch_bounds__(256) attn_stage_kernel(const float* __restrict__ src, const float* __restrict__ o,
                                                        __nv_bfloat16* __restrict__ rm, __nv_bfloat16* __restrict__ tr,
                                                        float* __restrict__ delta, int S, int d, int H, int E, int layout,
                                                        float scale) {
  __shared__ float tile[64][65];
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H, s0 = blockIdx.x * 64;
  if (layout == 0) {
    for (int idx = threadIdx.x; idx < d * 64; idx += 256) {
      const int dd = idx >> 6, ll = idx & 63;
      float v = 0.f;
      if (s0 + ll < S) {
        v = src[((size_t)b * E + h * d + dd) * S + s0 + ll] * scale;
        tr[((size_t)bh * d + dd) * S + s0 + ll] = __float2bfloat16_rn(v);
      }
      tile[ll][dd] = v;
    }
  } else {
    for (int idx = threadIdx.x; idx < d * 64; idx += 256) {
      const int ll = idx / d, dd = idx - ll * d;
      tile[ll][dd] = s0 + ll < S ? src[((size_t)b * S + s0 + ll) * E + h * d + dd] * scale : 0.f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < d * 64; idx += 256) {
      const int dd = idx >> 6, ll = idx & 63;
      if (s0 + ll < S) tr[((size_t)bh * d + dd) * S + s0 + ll] = __float2bfloat16_rn(tile[ll][dd]);
    }
    if (threadIdx.x < 64 && s0 + threadIdx.x < S && delta != nullptr) {
      const float* orow = o + ((size_t)b * S + s0 + threadIdx.x) * E + h * d;
      float acc = 0.f;
      for (int dd = 0; dd < d; ++dd) acc += tile[threadIdx.x][dd] * orow[dd];
      delta[(size_t)bh * S + s0 + threadIdx.x] = acc;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 256) {
    const int ll = idx >> 3, ch = idx & 7;
    if (s0 + ll >= S) continue;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d0 = ch * 8 + 2 * j;
      w[j] = pack_bf16x2(d0 < d ? tile[ll][d0] : 0.f, d0 + 1 < d ? tile[ll][d0 + 1] : 0.f);
    }
    *reinterpret_cast<uint4*>(rm + ((size_t)bh * S + s0 + ll) * 64 + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace attn

// ---- host side ----------------------------------------------------------------------------------------------------
using attn::Params;

// head_dim 64 would need 224 KB of operand stages in the dK / dV pass
static bool attn_shape_ok(int S, int d) { return d >= 16 && d <= 48 && d % 16 == 0 && S >= 1 && S % 8 == 0; }

int attention_supported(int batch, int heads, int seq, int head_dim) {
  return batch >= 1 && heads >= 1 && (long long)batch * heads <= 65535 && attn_shape_ok(seq, head_dim) ? 1 : 0;
}

size_t attention_rm_bytes(int batch, int heads, int seq) { return (size_t)batch * heads * seq * 64 * 2; }
size_t attention_tr_bytes(int batch, int heads, int seq, int d) { return (size_t)batch * heads * d * seq * 2; }

static int encode_rm(CUtensorMap* tm, const void* p, int S, int BH) {
  const uint64_t dims[3] = {64, (uint64_t)S, (uint64_t)BH};
  const uint64_t strides[2] = {128, (uint64_t)S * 128};
  const uint32_t box[3] = {64, 128, 1};
  return encode_tensor_map(tm, p, 2, 3, dims, strides, box, 3);
}
static int encode_tr(CUtensorMap* tm, const void* p, int S, int d, int BH) {
  const uint64_t dims[3] = {(uint64_t)S, (uint64_t)d, (uint64_t)BH};
  const uint64_t strides[2] = {(uint64_t)S * 2, (uint64_t)S * 2 * d};
  const uint32_t box[3] = {64, (uint32_t)d, 1};
  return encode_tensor_map(tm, p, 2, 3, dims, strides, box, 3);
}

int launch_attention_stage(const float* src, const float* o, void* rm, void* tr, float* delta, int batch, int heads, int S,
                           int d, int layout, float scale, cudaStream_t st) {
  dim3 grid((unsigned)((S + 63) / 64), (unsigned)(batch * heads));
  attn::attn_stage_kernel<<<grid, 256, 0, st>>>(src, o, reinterpret_cast<__nv_bfloat16*>(rm),
                                                reinterpret_cast<__nv_bfloat16*>(tr), delta, S, d, heads, heads * d, layout,
                                                scale);
  return check_launch("attn_stage_kernel");
}

template <int MODE>
static int launch_attn(const Params& p0, const void* a1, const void* a2, const void* b1, const void* b2, const void* b3,
                       const void* b4, cudaStream_t st) {
  Params p = p0;
  const bool bwd = MODE != attn::MODE_FWD;
  p.n_chunks = (p.S + attn::kChunk - 1) / attn::kChunk;
  p.b3_tile_bytes = (uint32_t)p.d * 128u;
  p.stage_bytes = (bwd ? 2u : 1u) * attn::kTileBytes + (MODE == attn::MODE_DKV ? 4u : 2u) * p.b3_tile_bytes;
  const size_t fixed = (bwd ? 2 : 1) * (size_t)attn::kTileBytes + (MODE == attn::MODE_DQ ? 1 : 2) * (size_t)attn::kPBytes + 1024;
  const size_t budget = 224 * 1024;
  int ns = (int)((budget - fixed) / p.stage_bytes);
  if (ns > attn::kMaxStages) ns = attn::kMaxStages;
  if (ns < 2) return fail(SELDQ_ERR_UNSUPPORTED, "attention: operand stages of %u B do not fit twice", p.stage_bytes);
  p.nstages = ns;
  const size_t smem = fixed + (size_t)ns * p.stage_bytes;
  alignas(64) CUtensorMap tA1, tA2, tB1, tB2, tB3, tB4;
  int rc;
  if ((rc = encode_rm(&tA1, a1, p.S, p.BH))) return rc;
  if ((rc = encode_rm(&tA2, a2 ? a2 : a1, p.S, p.BH))) return rc;
  if ((rc = encode_rm(&tB1, b1, p.S, p.BH))) return rc;
  if ((rc = encode_rm(&tB2, b2 ? b2 : b1, p.S, p.BH))) return rc;
  if ((rc = encode_tr(&tB3, b3, p.S, p.d, p.BH))) return rc;
  if ((rc = encode_tr(&tB4, b4 ? b4 : b3, p.S, p.d, p.BH))) return rc;
  auto kern = attn::attn_kernel<MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "attention smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)((p.S + attn::kRows - 1) / attn::kRows), (unsigned)p.BH);
  kern<<<grid, attn::kThreads, smem, st>>>(tA1, tA2, tB1, tB2, tB3, tB4, p);
  return check_launch("attn_kernel");
}

int launch_attention_fwd(int batch, int heads, int S, int d, const void* q_rm, const void* k_rm, const void* v_tr,
                         float* out, float* lse, cudaStream_t st) {
  Params p{};
  p.S = S; p.d = d; p.H = heads; p.BH = batch * heads; p.E = heads * d;
  p.out = out; p.lse = lse;
  return launch_attn<attn::MODE_FWD>(p, q_rm, nullptr, k_rm, nullptr, v_tr, nullptr, st);
}

int launch_attention_bwd(int batch, int heads, int S, int d, const void* q_rm, const void* k_rm, const void* v_rm,
                         const void* q_tr, const void* k_tr, const void* do_rm, const void* do_tr, float* lse,
                         const float* delta, float* dq, float* dk, float* dv, cudaStream_t st) {
  Params p{};
  p.S = S; p.d = d; p.H = heads; p.BH = batch * heads; p.E = heads * d;
  p.lse = lse; p.delta = delta;
  // q_rm / q_tr hold q * log2(e) / sqrt(d):  dK = (dS^T q_scaled) * ln 2,  dQ = (dS k) / sqrt(d)
  p.o1 = dv; p.o2 = dk; p.scale1 = 1.f; p.scale2 = 0.6931471805599453f;
  // The dK / dV pass and the dQ pass are independent.  Each is a grid of (row tiles x heads) CTAs, one per SM -- 152 at
  // the model's shape (19 x 8) on 148 SMs, i.e. two waves whose second holds four CTAs -- so the dQ pass runs on a side
  // stream forked off `st` and joined back: the 304 CTAs of both fill the SMs in three waves instead of four
  // (SELDQ_ATTN_BWD_FORK=0: one after the other).  The fork / join events are capturable into a CUDA graph.
  static const bool fork = [] { const char* e = getenv("SELDQ_ATTN_BWD_FORK"); return !(e && e[0] == '0'); }();
  if (!fork) {
    int rc = launch_attn<attn::MODE_DKV>(p, k_rm, v_rm, q_rm, do_rm, do_tr, q_tr, st);
    if (rc) return rc;
    p.o1 = dq; p.o2 = nullptr; p.scale1 = 1.f / sqrtf((float)d);
    return launch_attn<attn::MODE_DQ>(p, q_rm, do_rm, k_rm, v_rm, k_tr, nullptr, st);
  }
  ForkJoin fj(st, 0);
  int rc = launch_attn<attn::MODE_DKV>(p, k_rm, v_rm, q_rm, do_rm, do_tr, q_tr, st);
  if (rc == SELDQ_OK) {
    p.o1 = dq; p.o2 = nullptr; p.scale1 = 1.f / sqrtf((float)d);
    rc = launch_attn<attn::MODE_DQ>(p, q_rm, do_rm, k_rm, v_rm, k_tr, nullptr, fj.side());
  }
  if (!fj.join() && rc == SELDQ_OK) rc = fail(SELDQ_ERR_CUDA, "attention backward: join: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

}  // namespace seldq

// tcgen05 implicit-GEMM convolution for the quaternion / dual-quaternion layers, forward and dgrad,
// reading channels-last bf16 operands (conv_cl.h).
//
//   D[t, (a,o)] = sum_tap sum_b sum_i  sign[a][b] * X[t + off(tap), (b,i)] * W_{widx[a][b]}[o, i, tap]
//
//   * M = 128 consecutive w positions of one (n, h) row              -> TMEM lanes
//   * N = out channels of ONE out component a (padded to 16)         -> TMEM columns al*NBp ...
//   * K = (tap, 64-channel chunk) boxes streamed by TMA through an mbarrier ring; a box [128 w x 64 ch]
//     lands in the canonical K-major 128B-swizzled layout, each of its four 16-channel slabs belongs to
//     one in component b
//
// The Hamilton / dual-quaternion expansion (quaternion_ops.py:131-135, dual_quaternion_ops.py:122-140)
// is never materialised: the COMPACT weights, pre-packed once per optimiser step as bf16 UMMA tiles
// (pack_weights_kernel), are bulk-copied into shared memory and stay resident; every (a, b, slab) is one
// tcgen05.mma whose instruction descriptor carries sign[a][b] in the negate-B bit; the structural zero
// block of the dual quaternion is never issued, and channel chunks that only feed zero blocks of a CTA's
// out components are never loaded.
//
// Work unit = (position tile, group of out components).  Splitting the out components over CTAs fills the
// 148 SMs when there are few position tiles (batch 1: 38 tiles per TCN layer); heavier groups (the dual
// half of a DQ layer has twice the blocks) are scheduled first.
//
// Layers whose K side has < 8 channels per component (first CNN layer, Cin = 8 | 16) run "dense": the
// activation is one 16-channel-padded component and the signed expanded tile is built in shared memory
// only, one MMA spanning all out channels.
//
// Warp roles: warp 0 = TMA producer, warps 1 and 6.. = MMA issuers (kMmaWarps; warp 1 owns the TMEM allocation),
// warps 2..5 and kFirstExtraEpiWarp.. = epilogue (TMEM -> registers -> coalesced global stores, + bias): kEpiSets sets
// of four warps (one per TMEM lane quarter) that take alternate 16-channel chunks -- a single warp per quarter runs
// its dependent ld / convert / store chain at well under one instruction per cycle, which bounded the layers with
// little K per output (first CNN layer: 9 MMAs but 12 column chunks per tile).
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "launch.h"
#include "pdl.cuh"
#include "tensor_map.h"
#include "umma_ptx.cuh"

namespace seldq {
namespace cl {

// fp16 store of a convolution output, saturating (the value range of fp16 ends at 65504)
__device__ __forceinline__ __half to_f16_sat(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
// two values at once: cvt.rn.satfinite.f16x2.f32 clamps to +-65504 and packs {hi, lo} in one instruction
__device__ __forceinline__ uint32_t to_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Unit schedule of a CTA: units are ordered heaviest group first; round r hands unit r*G + b to CTA b in even rounds
// and r*G + (G-1-b) in odd rounds (snake order), so the few units of a last, partial round go to the CTAs that
// hold the lightest units of the round before instead of the heaviest.  Returns -1 when the CTA sits a round out.
__device__ __forceinline__ int unit_of_round(int round, int total_units) {
  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  const int u = round * G + ((round & 1) ? G - 1 - b : b);
  return u < total_units ? u : -1;
}

template <int GC>
__global__ void __launch_bounds__(kFpropThreads, 1)
qconv_cl_fprop_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  __shared__ uint2 op_tbl_s[kOpTableEntries];
  __shared__ __align__(16) uint8_t out_stage[4 * kEpiSets][8 * 64];    // per epilogue warp: [8 ch][32 w] fp16
  // warp index through a shuffle: provably warp-uniform, so role branches and the MMA loop compile to the
  // uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // layout: [activation ring][barriers, 1 KB][weight tiles][slack]
  uint8_t* a_ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  uint8_t* b_img = smem + (size_t)p.nstages * p.stage_bytes + 1024;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint64_t* w_bar = bars + 2 * kMaxStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

  // ---- one-time setup ---------------------------------------------------------------------------
  for (int i = threadIdx.x; i < p.op_entries; i += kFpropThreads) op_tbl_s[i] = p.op_tbl[i];
  if (p.dense) {
    // signed expanded weight -> bf16 B tiles.  Tile (tap, j) holds B[n][k], n = out channel, k = in channels
    // 16 j ... 16 j + 15; UMMA K-major / no swizzle: core matrix = 8 rows x 16 B, LBO (K step) = NBp*16,
    // SBO (8 rows) = 128
    const int items = p.ntaps * p.J * 2 * p.NBp;          // one item = 8 consecutive k of one row
    for (int it = threadIdx.x; it < items; it += kFpropThreads) {
      int r = it;
      const int n = r % p.NBp; r /= p.NBp;
      const int kc = r & 1; r >>= 1;
      const int j = r % p.J;
      const int tap = r / p.J;
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int ch = j * 16 + kc * 8 + jj;
        v[jj] = __float2bfloat16_rn((n < p.g.P && ch < p.g.R) ? expanded_weight(p.g, p.w, n, ch, tap) : 0.f);
      }
      uint8_t* dst = b_img + (size_t)(tap * p.J + j) * p.slab_bytes + (size_t)kc * (p.NBp * 16) + (n >> 3) * 128 +
                     (n & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
    ptx::fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    // two MMA-issuing warps: each arrives once per stage / accumulator
    for (int i = 0; i < p.nstages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], kMmaWarps); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], kMmaWarps); ptx::mbar_init(&tempty_bar[i], 4 * kEpiSets); }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_in);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = (uint32_t)p.acc_cols;

  if (warp == 0) {
    // ===== TMA producer ============================================================================
    if (ptx::elect_one()) {
      if (!p.dense) {
        // resident weight tiles: one bulk copy per <= 32 KB
        ptx::mbar_arrive_expect_tx(w_bar, p.w_bytes);
        for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
          const uint32_t n = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
          ptx::bulk_load(b_img + off, p.packed + off, n, w_bar);
        }
      }
      pdl_wait();            // activations of the previous kernel from here on (the weights above never race)
      uint32_t slot = 0, parity = 0;
      for (int round = 0; round * (int)gridDim.x < p.total_units; ++round) {
        const int u = unit_of_round(round, p.total_units);
        if (u < 0) continue;
        const int group = p.group_order[u / p.total_tiles];
        int r = u % p.total_tiles;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int h = r % p.OH;
        const int n = r / p.OH;
        const int w0 = wt * kTileM;
        const uint32_t mask = p.chunk_mask[group];
        for (int c = 0; c < p.chunks; ++c) {
          if (!((mask >> c) & 1u)) continue;
          for (int tap = 0; tap < p.ntaps; tap += p.tps) {
            ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[slot], p.stage_tx);
            if (p.rs)                                // one box of 128 + span rows serves the taps of a kernel row
              ptx::tma_load_4d(a_ring + (size_t)slot * p.stage_bytes, &tm_in, &full_bar[slot], c * p.BK,
                               w0 + p.rs_min_off, h + p.off_h[tap], n);
            else
              for (int tl = 0; tl < p.tps; ++tl)     // one stage = the boxes of p.tps taps
                ptx::tma_load_4d(a_ring + (size_t)slot * p.stage_bytes + (size_t)tl * p.box_bytes, &tm_in,
                                 &full_bar[slot], c * p.BK, w0 + p.off_w[tap + tl], h + p.off_h[tap + tl], n);
            if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || (warp >= 6 && warp < kFirstExtraEpiWarp)) {
    // ===== MMA issuers =============================================================================
    // Lane-parallel issue: the MMAs of one stage (<= 32: slabs x out components of the group, structural zero
    // blocks left out) form a dense list in (slab, component) order.  All lanes build their descriptors at once;
    // the tcgen05.mma under the lane predicate then costs one short elect-and-issue round per active lane (~40
    // cycles per MMA instead of ~100 for descriptor arithmetic on the uniform datapath; tools/umma_rate.py).
    // kMmaWarps warps share the list (entry i goes to warp i mod kMmaWarps) and issue concurrently: accumulating MMAs commute, so only
    // the accumulator-initialising ones (tap 0, one per out component, distinct columns) need an order -- they go
    // first, from all warps, and a named barrier separates them from the rest of that stage.
    // The accumulate flag of an issue site must be warp-uniform (ptxas derives the instruction's predicate with a
    // vote over the issuing lanes), hence the separate sites.  tcgen05.commit tracks the MMAs of the executing
    // thread; the tensor pipe retires in order, so each warp's commit comes from the lane that issued last.
    {
      const int me = warp == 1 ? 0 : warp - 5;
      const uint64_t a_hi = ptx::make_smem_desc_hi(16, p.a_sbo, p.a_swz);                 // K-major, swizzled
      const uint64_t b_hi = ptx::make_smem_desc_hi((uint32_t)p.NBmma * 16u, 128, ptx::kSwizzleNone);
      const uint32_t a_base = ptx::smem_u32(a_ring), b_lo16 = ptx::smem_u32(b_img) >> 4;
      const int lanes_per_chunk = p.slabs_per_chunk * p.mma_per_slab;   // op-table entries of one (group, chunk)
      const int lanes_per_stage = lanes_per_chunk * p.tps;         // a stage holds p.tps taps (1, or all of them)
      const int my_entry = kMmaWarps * lane + me;
      const int my_tl = my_entry / lanes_per_chunk;                // tap of the stage this lane's MMA belongs to
      const int my_rest = my_entry - my_tl * lanes_per_chunk;      // ... and its entry in the chunk's op list
      // where that tap's A operand starts inside a stage: its own box, or a row offset into the shared box
      const uint32_t my_tap_off16 = my_tl >= p.tps ? 0u
                                    : p.rs ? (uint32_t)p.rs_row[my_tl] * ((uint32_t)p.BK * 2u >> 4)
                                           : ((uint32_t)my_tl * p.box_bytes) >> 4;
      if (!p.dense) ptx::mbar_wait(w_bar, 0);
      uint32_t slot = 0, parity = 0, it = 0;
      for (int round = 0; round * (int)gridDim.x < p.total_units; ++round) {
        const int u = unit_of_round(round, p.total_units);
        if (u < 0) continue;
        const int gi = u / p.total_tiles;
        const uint32_t mask = p.chunk_mask[p.group_order[gi]];
        const uint32_t as = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t use = p.acc_stages == 2 ? (it >> 1) : it;
        ptx::mbar_wait(&tempty_bar[as], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_unit = tmem_base + as * acc_cols;
        const uint2* tbl_g = op_tbl_s + (size_t)p.group_order[gi] * p.chunks * lanes_per_chunk;
        int last_closer = -1;                                  // lane that issued this warp's last MMA of the unit
        for (int c = 0; c < p.chunks; ++c) {
          if (!((mask >> c) & 1u)) continue;
          // per chunk: this lane's MMA (the same for every tap up to the tap's weight-tile offset)
          uint2 e = make_uint2(0u, 0u);
          if (my_entry < lanes_per_stage) e = tbl_g[c * lanes_per_chunk + my_rest];
          const bool valid = (int)e.x < 0;
          const bool first = valid && (e.x & (1u << 30)) != 0u && my_tl == 0;
          const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
          const bool any = vmask != 0u;
          const int closer = any ? 31 - __clz((int)vmask) : -1;
          if (any) last_closer = closer;
          const uint32_t a_off16 = ((e.x >> 16) & 3u) * 2u + my_tap_off16;
          const uint32_t d_lane = d_unit + ((e.x >> 20) & 0x1ffu);
          uint32_t b16 = b_lo16 + (e.x & 0x3fffu) + (uint32_t)my_tl * p.tap_stride16;
          for (int tap = 0; tap < p.ntaps; tap += p.tps, b16 += p.tap_stride16 * (uint32_t)p.tps) {
            const uint64_t a_desc = a_hi | (uint64_t)((((a_base + slot * p.stage_bytes) >> 4) + a_off16) & 0x3fffu);
            const uint64_t b_desc = b_hi | (uint64_t)(b16 & 0x3fffu);
            ptx::mbar_wait(&full_bar[slot], parity);
            ptx::tc_fence_after();
            if (tap == 0) {
              if (first) ptx::umma_f16(d_lane, a_desc, b_desc, e.y, 0u);
              __syncwarp();
              asm volatile("bar.sync 2, %0;" ::"n"(32 * kMmaWarps) : "memory");      // every warp's initialising MMAs are issued
              if (valid && !first) ptx::umma_f16(d_lane, a_desc, b_desc, e.y, 1u);
            } else {
              if (valid) ptx::umma_f16(d_lane, a_desc, b_desc, e.y, 1u);
            }
            __syncwarp();
            if (any) {
              if (lane == closer) ptx::umma_commit(&empty_bar[slot]);
            } else if (lane == 0) {
              ptx::mbar_arrive(&empty_bar[slot]);               // nothing of this warp reads the slot
            }
            __syncwarp();
            if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
          }
        }
        // accumulator complete -> epilogue (one arrival per MMA warp)
        if (last_closer >= 0) {
          if (lane == last_closer) ptx::umma_commit(&tfull_bar[as]);
        } else if (lane == 0) {
          ptx::mbar_arrive(&tfull_bar[as]);
        }
        __syncwarp();
        ++it;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global ===================================================
    pdl_wait();
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int eset = warp >= kFirstExtraEpiWarp ? 1 + (warp - kFirstExtraEpiWarp) / 4 : 0;
    const int row = q * 32 + lane;
    // 16-byte bf16 stores need 8-element alignment of every row start
    const bool vec16 = p.out16 != nullptr && (p.OW & 7) == 0 && (p.out_sC & 7) == 0 && (p.out_sH & 7) == 0 &&
                       (p.out_sN & 7) == 0 && (reinterpret_cast<unsigned long long>(p.out16) & 15ull) == 0;
    uint32_t it = 0;
    for (int round = 0; round * (int)gridDim.x < p.total_units; ++round) {
      const int u = unit_of_round(round, p.total_units);
      if (u < 0) continue;
      const int group = p.group_order[u / p.total_tiles];
      const uint32_t as = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t use = p.acc_stages == 2 ? (it >> 1) : it;
      int r = u % p.total_tiles;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int h = r % p.OH;
      const int n = r / p.OH;
      const int w = wt * kTileM + row;
      const bool w_ok = w < p.OW;
      ptx::mbar_wait(&tfull_bar[as], use & 1);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols;
      const long long row_off = (long long)n * p.out_sN + (long long)h * p.out_sH + w;
      float* out_row = p.out + row_off;
      __half* out16_row = p.out16 + row_off;
      int piece_no = 0;
      for (int al = 0; al < GC; ++al) {
        const int ch_base = p.comp_of[group][al] * p.Pc;
        for (int c0 = 0; c0 < p.Pc; c0 += 16, ++piece_no) {
          if (piece_no % kEpiSets != eset) continue;
          uint32_t v[16];
          if (!p.fuse) {
            ptx::tmem_ld16(t_row + (uint32_t)(al * p.NBp + c0), v);
            ptx::tmem_ld_wait();
          } else {
            // fusion: channels [c0, c0 + 8) and [c0 + 8, c0 + 16) of this component are 8-column groups of the
            // (up to four) column sets of its pair / quad; sum them with the component's signs (conv_cl.h)
            const uint32_t og_stride = 8u * (uint32_t)p.fuse;
            // all loads of the piece in flight before the one wait (a wait per column set would serialise four
            // tensor-memory round trips); sets that do not exist load column 0 and are multiplied by 0
            uint32_t u[4][2][8];
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const bool on = p.epi_sgn[group][al][st] != 0 && c0 + hf * 8 < p.Pc;       // warp-uniform
                if (on)
                  ptx::tmem_ld8(t_row + (uint32_t)p.epi_col[group][al][st] + (uint32_t)((c0 >> 3) + hf) * og_stride, u[st][hf]);
                else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) u[st][hf][j] = 0u;
                }
              }
            ptx::tmem_ld_wait();
            const float s0 = (float)p.epi_sgn[group][al][0], s1 = (float)p.epi_sgn[group][al][1];
            const float s2 = (float)p.epi_sgn[group][al][2], s3 = (float)p.epi_sgn[group][al][3];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[hf * 8 + j] = __float_as_uint(fmaf(s3, __uint_as_float(u[3][hf][j]), fmaf(s2, __uint_as_float(u[2][hf][j]),
                                                fmaf(s1, __uint_as_float(u[1][hf][j]), s0 * __uint_as_float(u[0][hf][j])))));
          }
          if (p.out16 && vec16) {
            // fp16 output in the tensor's own NCHW order (fused CNN-block path, epilogue.cu): the warp's
            // [32 w x 16 ch] block is transposed through shared memory and leaves as 16-byte pieces
            const int lim = p.Pc - c0;
            uint8_t* stg = out_stage[eset * 4 + q];
#pragma unroll
            for (int r = 0; r < 2; ++r) {                      // 8 channels per pass: [8 ch][32 w] staging
              if (p.bias == nullptr) {        // the common case (--use_bias_conv=False): no per-element select / load / add
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  const uint32_t h2 = to_f16x2_sat(__uint_as_float(v[r * 8 + j]), __uint_as_float(v[r * 8 + j + 1]));
                  *reinterpret_cast<uint16_t*>(stg + j * 64 + lane * 2) = (uint16_t)(h2 & 0xffffu);
                  *reinterpret_cast<uint16_t*>(stg + (j + 1) * 64 + lane * 2) = (uint16_t)(h2 >> 16);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<__half*>(stg + j * 64 + lane * 2) = to_f16_sat(
                      __uint_as_float(v[r * 8 + j]) + (r * 8 + j < lim ? __ldg(p.bias + ch_base + c0 + r * 8 + j) : 0.f));
              }
              __syncwarp();
              const int ch = lane >> 2, piece = lane & 3;
              const int wp = wt * kTileM + q * 32 + piece * 8;
              if (r * 8 + ch < lim && wp < p.OW) {
                const uint4 val = *reinterpret_cast<const uint4*>(stg + ch * 64 + piece * 16);
                *reinterpret_cast<uint4*>(p.out16 + (long long)n * p.out_sN + (long long)h * p.out_sH +
                                          (long long)(ch_base + c0 + r * 8 + ch) * p.out_sC + wp) = val;
              }
              __syncwarp();
            }
          } else if (w_ok && p.out16) {
            __half* dst = out16_row + (long long)(ch_base + c0) * p.out_sC;
            const int lim = p.Pc - c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < lim)
                *dst = to_f16_sat(__uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + ch_base + c0 + j) : 0.f));
              dst += p.out_sC;
            }
          } else if (w_ok) {
            float* dst = out_row + (long long)(ch_base + c0) * p.out_sC;
            const int lim = p.Pc - c0;
            if (lim >= 16 && p.bias == nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) { *dst = __uint_as_float(v[j]); dst += p.out_sC; }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j < lim) *dst = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + ch_base + c0 + j) : 0.f);
                dst += p.out_sC;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      ++it;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// compact fp32 weights -> bf16 UMMA B tiles [img][tap][j][NBp x 16] (K-major, no swizzle; see the dense
// prologue above for the byte layout).  Row n / column k of tile (img, tap, j):
//   forward : W_img[o = n][i = 16 j + k][tap]        dgrad : W_img[o = 16 j + k][i = n][tap]
// Fusion (pair_xor = m != 0; F = 2 for m = 1 | 2, F = 4 for m = 3): tile (image set q, tap, j) is [F NB8 x 16]; its
// 8-row groups cycle through the F images {i0 | sub : sub a submask of m} of the set:
// row (n / 8) * 8 F + slot * 8 + n % 8, slot = the bits of the image index under m.
struct PackParams {
  const float* w[8];
  uint8_t* dst;
  int n_img, ntaps, J, NBp;
  int rows_real, k_real;        // real extent of the row (N side) and K side
  int transposed;
  int wsO, wsI, wsT;
  int pair_xor, NB8;            // pair fusion; rows per image = NB8 then, NBp otherwise
};
__device__ __forceinline__ int pack_rows(const PackParams& p) { return p.pair_xor ? p.NB8 : p.NBp; }
// position of image (or component) index i inside its fusion set, and the index of that set: the bits of i under
// the mask m, and the remaining bits squeezed together
__host__ __device__ __forceinline__ int fuse_slot(int i, int m) { return m == 1 ? (i & 1) : m == 2 ? ((i >> 1) & 1) : (i & 3); }
__host__ __device__ __forceinline__ int fuse_set(int i, int m) { return m == 1 ? (i >> 1) : m == 2 ? (((i >> 2) << 1) | (i & 1)) : (i >> 2); }
__device__ __forceinline__ void pack_item(const PackParams& p, int it) {
  const int rows = pack_rows(p);
  int r = it;
  const int n = r % rows; r /= rows;
  const int kc = r & 1; r >>= 1;
  const int j = r % p.J; r /= p.J;
  const int tap = r % p.ntaps;
  const int img = r / p.ntaps;
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int k = j * 16 + kc * 8 + jj;
    float x = 0.f;
    if (n < p.rows_real && k < p.k_real) {
      const int o = p.transposed ? k : n, i = p.transposed ? n : k;
      x = __ldg(p.w[img] + (long long)o * p.wsO + (long long)i * p.wsI + (long long)tap * p.wsT);
    }
    v[jj] = __float2bfloat16_rn(x);
  }
  uint8_t* dst;
  if (p.pair_xor) {
    const int m = p.pair_xor, F = m == 3 ? 4 : 2;
    const int slot = fuse_slot(img, m), q = fuse_set(img, m);
    const int nf = (n >> 3) * (8 * F) + slot * 8 + (n & 7), NBf = F * p.NB8;
    dst = p.dst + ((size_t)(q * p.ntaps + tap) * p.J + j) * ((size_t)NBf * 32) + (size_t)kc * (NBf * 16) +
          (nf >> 3) * 128 + (nf & 7) * 16;
  } else {
    dst = p.dst + ((size_t)(img * p.ntaps + tap) * p.J + j) * ((size_t)p.NBp * 32) + (size_t)kc * (p.NBp * 16) +
          (n >> 3) * 128 + (n & 7) * 16;
  }
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
}
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackParams p) {
  const int items = p.n_img * p.ntaps * p.J * 2 * pack_rows(p);
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) pack_item(p, it);
}

// the same for many layers in one launch: blockIdx.y selects an entry of a device-resident parameter table
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackParams* __restrict__ table) {
  const PackParams p = table[blockIdx.y];
  const int items = p.n_img * p.ntaps * p.J * 2 * pack_rows(p);
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) pack_item(p, it);
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      return 148;
  }
  return g_num_sms;
}

}  // namespace cl

// ---- host side: layouts, plan, launch -------------------------------------------------------------
using cl::FpropParams;

cl::OperandLayout x_operand_layout(const ConvGeom& fwd) {
  return cl::operand_layout(fwd.tab.nc, fwd.tab.nc * fwd.Ic, fwd.tab.nc == 1 || fwd.Ic < 8);
}
cl::OperandLayout gy_operand_layout(const ConvGeom& fwd) {
  return cl::operand_layout(fwd.tab.nc, fwd.tab.nc * fwd.Oc, fwd.tab.nc == 1 || fwd.Oc < 8);
}

// block (pass-out component a, pass-in component b) of the expanded weight: image index (-1: structural zero) and sign
static void pass_block(const ConvGeom& g, int a, int b, int* img, int* neg) {
  const int fa = g.transposed ? b : a, fb = g.transposed ? a : b;   // forward-sense (out, in) components
  *img = g.tab.widx[fa][fb];
  *neg = g.tab.sign[fa][fb] < 0;
}

// Fusion (conv_cl.h): with the out components grouped as {a0 | sub : sub a submask of m} (pairs for m = 1 | 2, quads
// for m = 3), does every group see at most `max_sets` (slot order, relative signs) classes over the in components?
// Fills set_of[a0][b] (a0 = lowest member) with the class index of block column b, or -1 where the group has a
// structural zero, and returns the number of classes needed (0: this grouping does not work).
static int fusion_sets(const ConvGeom& g, int m, int max_sets, int8_t set_of[8][8]) {
  const int nc = g.tab.nc, F = m == 3 ? 4 : 2;
  int need = 0;
  for (int a0 = 0; a0 < nc; ++a0) {
    if (a0 & m) continue;
    int nclass = 0, key[4] = {0, 0, 0, 0};
    for (int b = 0; b < nc; ++b) {
      int e[4], sg[4], nz = 0, used = 0, k = 0;
      for (int t = 0, sub = 0; t < F; ++t, sub = (sub - m) & m) {      // submasks of m in increasing order
        pass_block(g, a0 | sub, b, &e[t], &sg[t]);
        nz += e[t] >= 0;
      }
      set_of[a0][b] = -1;
      if (nz == 0) continue;
      if (nz != F) return 0;
      for (int t = 0; t < F; ++t) {
        if (cl::fuse_set(e[t], m) != cl::fuse_set(e[0], m)) return 0;
        used |= 1 << cl::fuse_slot(e[t], m);
        k = (k << 3) | (cl::fuse_slot(e[t], m) << 1) | (sg[t] ^ sg[0]);
      }
      if (used != (1 << F) - 1) return 0;
      int c = 0;
      while (c < nclass && key[c] != k) ++c;
      if (c == nclass) {
        if (nclass == max_sets) return 0;
        key[nclass++] = k;
      }
      set_of[a0][b] = (int8_t)c;
    }
    need = nclass > need ? nclass : need;
  }
  return need;
}

// geometry of the resident weight tiles of one pass
struct WeightPlan {
  int dense, n_img, ntaps, J, NBp, rows_real, k_real;
  int fuse, pair_xor, nsets, NB8, NBmma;   // fusion factor F (0 | 2 | 4): a tile holds the F images of a set, N = F * NB8
  int n_tilesets;                   // images, or image sets when fused
  size_t slab_bytes, img_bytes, total;
};
static WeightPlan weight_plan(const ConvGeom& g) {
  WeightPlan w;
  memset(&w, 0, sizeof(w));
  w.dense = cl::is_dense(g) ? 1 : 0;
  w.ntaps = g.KH * g.KW;
  if (w.dense) {
    const cl::OperandLayout l = cl::operand_layout(1, g.R, true);
    w.n_img = 1; w.J = l.Cp / 16; w.NBp = cl::round_up(g.P, 16); w.rows_real = g.P; w.k_real = g.R;
  } else {
    const int kc = g.transposed ? g.Oc : g.Ic, pc = g.transposed ? g.Ic : g.Oc;
    w.n_img = g.tab.nw; w.J = cl::round_up(kc, 16) / 16; w.NBp = cl::round_up(pc, 16); w.rows_real = pc; w.k_real = kc;
    static const bool enabled = getenv("SELDQ_PAIR_FUSE") == nullptr || atoi(getenv("SELDQ_PAIR_FUSE")) != 0;
    static const bool quads = getenv("SELDQ_QUAD_FUSE") == nullptr || atoi(getenv("SELDQ_QUAD_FUSE")) != 0;
    if (enabled && g.tab.nc >= 4 && (pc & 7) == 0 && pc <= 128 && (w.n_img & 3) == 0) {
      int8_t scratch[8][8];
      // quads need four column sets of 4 * pc accumulator columns each: 16 * pc <= 512.  SELDQ_QUAD_FUSE=2 takes them
      // only where that fits TWICE (pc <= 16), i.e. where MMA and epilogue phases of successive units still overlap
      static const int quad_cols = (getenv("SELDQ_QUAD_FUSE") && atoi(getenv("SELDQ_QUAD_FUSE")) == 2) ? 256 : 512;
      if (quads && 16 * pc <= quad_cols && (w.nsets = fusion_sets(g, 3, 4, scratch)) > 0) { w.fuse = 4; w.pair_xor = 3; }
      for (int m = 1; m <= 2 && !w.fuse; ++m)
        if ((w.nsets = fusion_sets(g, m, 2, scratch)) > 0) { w.fuse = 2; w.pair_xor = m; }
    }
  }
  w.NB8 = cl::round_up(w.rows_real, 8);
  w.NBmma = w.fuse ? w.fuse * w.NB8 : w.NBp;
  w.n_tilesets = w.fuse ? w.n_img / w.fuse : w.n_img;
  w.slab_bytes = (size_t)w.NBmma * 32;
  w.img_bytes = (size_t)w.ntaps * w.J * w.slab_bytes;
  w.total = (size_t)w.n_tilesets * w.img_bytes;
  return w;
}

size_t packed_weight_bytes(const ConvGeom& g) {
  const WeightPlan w = weight_plan(g);
  return w.dense ? 0 : w.total;
}

int launch_pack_weights(const ConvGeom& g, const float* const* host_w, void* packed, cudaStream_t st) {
  const WeightPlan w = weight_plan(g);
  if (w.dense) return SELDQ_OK;      // dense layers build their tile from the fp32 weights inside the kernel
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(SELDQ_ERR_INVALID, "packed weight buffer must be 16-byte aligned");
  cl::PackParams p{};
  for (int i = 0; i < w.n_img; ++i) p.w[i] = host_w[i];
  p.dst = reinterpret_cast<uint8_t*>(packed);
  p.n_img = w.n_img; p.ntaps = w.ntaps; p.J = w.J; p.NBp = w.NBp;
  p.rows_real = w.rows_real; p.k_real = w.k_real; p.transposed = g.transposed;
  p.wsO = g.wsO; p.wsI = g.wsI; p.wsT = g.wsT;
  p.pair_xor = w.fuse ? w.pair_xor : 0; p.NB8 = w.NB8;
  const int items = w.n_img * w.ntaps * w.J * 2 * (w.fuse ? w.NB8 : w.NBp);
  int blocks = (items + 255) / 256;
  if (blocks > 4 * cl::num_sms()) blocks = 4 * cl::num_sms();
  cl::pack_weights_kernel<<<blocks, 256, 0, st>>>(p);
  return check_launch("pack_weights_kernel");
}

size_t pack_table_entry_bytes() { return sizeof(cl::PackParams); }

// fills one host-side table entry; returns the number of 8-element items of that entry (0 for dense layers)
int fill_pack_table_entry(const ConvGeom& g, const float* const* host_w, void* packed, void* entry, int* items) {
  const WeightPlan w = weight_plan(g);
  cl::PackParams p{};
  *items = 0;
  if (!w.dense) {
    if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(SELDQ_ERR_INVALID, "packed weight buffer must be 16-byte aligned");
    for (int i = 0; i < w.n_img; ++i) p.w[i] = host_w[i];
    p.dst = reinterpret_cast<uint8_t*>(packed);
    p.n_img = w.n_img; p.ntaps = w.ntaps; p.J = w.J; p.NBp = w.NBp;
    p.rows_real = w.rows_real; p.k_real = w.k_real; p.transposed = g.transposed;
    p.wsO = g.wsO; p.wsI = g.wsI; p.wsT = g.wsT;
    p.pair_xor = w.fuse ? w.pair_xor : 0; p.NB8 = w.NB8;
    *items = w.n_img * w.ntaps * w.J * 2 * (w.fuse ? w.NB8 : w.NBp);
  }
  memcpy(entry, &p, sizeof(p));
  return SELDQ_OK;
}

int launch_pack_weights_multi(const void* dev_table, int count, int max_items, cudaStream_t st) {
  if (count <= 0 || max_items <= 0) return SELDQ_OK;
  int bx = (max_items + 255) / 256;
  if (bx > 16) bx = 16;
  dim3 grid((unsigned)bx, (unsigned)count);
  cl::pack_weights_multi_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const cl::PackParams*>(dev_table));
  return check_launch("pack_weights_multi_kernel");
}

int plan_cl_fprop(const ConvGeom& g, FpropParams* p, size_t* smem_bytes) {
  memset(p, 0, sizeof(*p));
  if (g.sh != 1 || g.sw != 1)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 tensor-core path implements stride 1 only (got %dx%d)", g.sh, g.sw);
  const int ntaps = g.KH * g.KW;
  if (ntaps > cl::kMaxTaps) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most %d taps", cl::kMaxTaps);
  const WeightPlan w = weight_plan(g);
  const int nc = g.tab.nc;
  p->g = g;
  p->ntaps = ntaps;
  p->N = g.N; p->OH = g.OH; p->OW = g.OW;
  for (int t = 0; t < ntaps; ++t) {
    const int kh = t / g.KW, kw = t % g.KW;
    // forward: in = out - pad + k*dil ; dgrad: gy position = gx position + pad - k*dil   (stride 1)
    p->off_h[t] = g.transposed ? g.ph - kh * g.dh : kh * g.dh - g.ph;
    p->off_w[t] = g.transposed ? g.pw - kw * g.dw : kw * g.dw - g.pw;
  }
  p->dense = w.dense;
  cl::OperandLayout l;
  for (int b = 0; b < 8; ++b)
    for (int a = 0; a < 8; ++a) p->op_img[b][a] = -1;
  if (w.dense) {
    if (g.P > 256)
      return fail(SELDQ_ERR_UNSUPPORTED, "bf16 dense mode (K side < 8 channels per component) needs <= 256 out channels, got %d",
                  g.P);
    l = cl::operand_layout(1, g.R, true);
    p->ncomp_out = 1; p->Pc = g.P;
    p->op_img[0][0] = 0; p->op_neg[0][0] = 0;
  } else {
    l = cl::operand_layout(nc, g.R, false);
    p->ncomp_out = nc; p->Pc = g.transposed ? g.Ic : g.Oc;
    for (int b = 0; b < nc; ++b)
      for (int a = 0; a < nc; ++a) {
        const int fa = g.transposed ? b : a, fb = g.transposed ? a : b;   // forward-sense (out, in) components
        const int e = g.tab.widx[fa][fb];
        if (e < 0) continue;
        p->op_img[b][a] = (int8_t)e;
        p->op_neg[b][a] = (int8_t)(g.tab.sign[fa][fb] < 0);
      }
  }
  p->NBp = w.NBp; p->J = w.J; p->n_img = w.n_img;
  p->slab_bytes = (uint32_t)w.slab_bytes; p->img_bytes = (uint32_t)w.img_bytes; p->w_bytes = (uint32_t)w.total;
  p->BK = l.BK; p->chunks = l.Cp / l.BK; p->slabs_per_chunk = l.BK / 16; p->cpad_in = l.cpad;
  if (p->chunks > 32) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most 2048 padded input channels");
  p->box_bytes = (uint32_t)cl::kTileM * l.BK * 2;
  p->tps = 1;
  p->a_sbo = 8u * l.BK * 2;
  p->a_swz = l.BK == 64 ? ptx::kSwizzle128B : (l.BK == 32 ? ptx::kSwizzle64B : ptx::kSwizzle32B);

  p->tiles_w = (g.OW + cl::kTileM - 1) / cl::kTileM;
  const long long tiles = (long long)g.N * g.OH * p->tiles_w;
  if (tiles > 0x0fffffffLL) return fail(SELDQ_ERR_UNSUPPORTED, "too many tiles");
  p->total_tiles = (int)tiles;
  // out-component groups: as few as TMEM allows, more while that helps to fill the SMs.  A component costs NBp
  // accumulator columns, or 2 * NB8 when pairs are fused (two column sets per pair); fused groups hold whole pairs.
  p->fuse = w.fuse; p->pair_xor = w.pair_xor; p->NB8 = w.NB8; p->NBmma = w.NBmma;
  const int F = w.fuse ? w.fuse : 1, S = w.fuse ? w.nsets : 1;
  const int cols_per_comp = w.fuse ? S * w.NB8 : w.NBp;                // a fused set of F components: S * F * NB8
  const int max_groups = p->ncomp_out / F;
  int ngroups = 1;
  while (p->ncomp_out / ngroups * cols_per_comp > 512 && ngroups < max_groups) ngroups *= 2;
  if (p->ncomp_out / ngroups * cols_per_comp > 512)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most 512 out channels per component");
  while (ngroups < max_groups && tiles * ngroups * 2 <= cl::num_sms() + cl::num_sms() / 16) ngroups *= 2;
  // SELDQ_ACC_DOUBLE=1: split once more if that lets the accumulators double-buffer (the epilogue of a unit then
  // overlaps the MMAs of the next; costs re-loading the shared channel chunks per group)
  if (getenv("SELDQ_ACC_DOUBLE") && atoi(getenv("SELDQ_ACC_DOUBLE")) != 0 && ngroups < max_groups &&
      p->ncomp_out / ngroups * cols_per_comp * 2 > 512 && p->ncomp_out / (ngroups * 2) * cols_per_comp * 2 <= 512 &&
      tiles * ngroups >= cl::num_sms())
    ngroups *= 2;
  p->ngroups = ngroups;
  p->gc = p->ncomp_out / ngroups;
  p->mma_per_slab = p->gc / F;
  p->acc_cols = p->gc * cols_per_comp;
  p->total_units = (int)(tiles * ngroups);
  // members of the groups.  Unfused: consecutive components.  Fused: consecutive PAIRS {a0, a0 ^ pair_xor} in the
  // order of their lower members; local component 2 q + t is member t of the group's pair q.
  int8_t set_of[8][8];
  if (w.fuse) {
    if (fusion_sets(g, w.pair_xor, S, set_of) != S) return fail(SELDQ_ERR_INVALID, "fusion: inconsistent grouping");
    int lower[4], nl = 0;
    for (int a = 0; a < p->ncomp_out; ++a)
      if (!(a & w.pair_xor)) lower[nl++] = a;
    const int spg = p->gc / F;                       // fused sets per group; local component F q + t = member t of set q
    for (int gi = 0; gi < ngroups; ++gi)
      for (int q = 0; q < spg; ++q)
        for (int t = 0, sub = 0; t < F; ++t, sub = (sub - w.pair_xor) & w.pair_xor)
          p->comp_of[gi][F * q + t] = (int8_t)(lower[gi * spg + q] | sub);
  } else {
    for (int gi = 0; gi < ngroups; ++gi)
      for (int al = 0; al < p->gc; ++al) p->comp_of[gi][al] = (int8_t)(gi * p->gc + al);
  }
  // cost (number of non-zero blocks) and needed chunks per group
  int cost[8];
  for (int gi = 0; gi < ngroups; ++gi) {
    cost[gi] = 0;
    uint32_t mask = 0;
    for (int c = 0; c < p->chunks; ++c)
      for (int s = 0; s < p->slabs_per_chunk; ++s) {
        const int b = (c * l.BK + s * 16) / l.cpad;
        for (int al = 0; al < p->gc; ++al)
          if (p->op_img[b][p->comp_of[gi][al]] >= 0) { mask |= 1u << c; ++cost[gi]; }
      }
    p->chunk_mask[gi] = mask;
    p->group_order[gi] = gi;
  }
  for (int i = 1; i < ngroups; ++i)                  // insertion sort, heaviest first
    for (int j = i; j > 0 && cost[p->group_order[j]] > cost[p->group_order[j - 1]]; --j) {
      const int t = p->group_order[j]; p->group_order[j] = p->group_order[j - 1]; p->group_order[j - 1] = t;
    }

  // MMA op table (conv_cl.h): per (group, chunk) the valid MMAs in (slab, component | pair) order, zero-padded to
  // slabs_per_chunk * mma_per_slab entries.  `first` = the first MMA of a unit into its accumulator columns
  // (chunks outside the group's mask hold no valid entry for its components by construction).
  const int lps = p->slabs_per_chunk * p->mma_per_slab;
  p->op_entries = ngroups * p->chunks * lps;
  if (p->op_entries > cl::kOpTableEntries)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: too many (slab, component) pairs for the MMA op table");
  if (lps > 32) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: more than 32 MMAs per stage");
  if (p->acc_cols > 512) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: accumulators do not fit in tensor memory");
  // Narrow K side (first CNN layer: one 16-channel slab per tap): a stage per tap is a 4 KB box and one MMA, so the
  // producer's and the issuers' per-stage bookkeeping (~0.3 us) would bound the layer; all taps share one stage then.
  if (l.BK <= 16 && ntaps * lps <= 32 && !getenv("SELDQ_NO_TPS")) p->tps = ntaps;
  p->stage_bytes = p->box_bytes * (uint32_t)p->tps;
  p->box_rows = cl::kTileM;
  // Row-shared taps (conv_cl.h): the KW taps of a kernel row from one box, if their shifts span at most 8 rows
  if (p->tps == 1 && l.BK == 64 && g.KW > 1 && (g.KW - 1) * g.dw <= 8 && g.KW * lps <= 32 * cl::kMmaWarps &&
      !getenv("SELDQ_NO_RS")) {
    int mn = p->off_w[0];
    for (int t = 0; t < g.KW; ++t) mn = p->off_w[t] < mn ? p->off_w[t] : mn;
    p->rs = 1; p->tps = g.KW; p->rs_min_off = mn;
    for (int t = 0; t < g.KW; ++t) p->rs_row[t] = p->off_w[t] - mn;        // the same for every kernel row
    p->box_rows = cl::kTileM + (g.KW - 1) * g.dw;
    p->stage_bytes = (uint32_t)cl::round_up(p->box_rows * l.BK * 2, 1024);
  }
  p->stage_tx = p->rs ? (uint32_t)(p->box_rows * l.BK * 2) : p->stage_bytes;
  p->tap_stride16 = (uint32_t)(((size_t)p->J * p->slab_bytes) >> 4);
  {
    const uint32_t idesc = ptx::make_idesc_bf16(cl::kTileM, (uint32_t)p->NBmma, 0, 0, 0, 0);
    for (int gi = 0; gi < ngroups; ++gi) {
      bool seen[8][4];
      memset(seen, 0, sizeof(seen));
      for (int al = 0; al < 8; ++al)
        for (int st = 0; st < 4; ++st) { p->epi_col[gi][al][st] = 0; p->epi_sgn[gi][al][st] = 0; }
      for (int c = 0; c < p->chunks; ++c) {
        uint2* dst = p->op_tbl + ((size_t)gi * p->chunks + c) * lps;
        int n = 0;
        for (int s = 0; s < p->slabs_per_chunk; ++s)
          for (int ml = 0; ml < p->mma_per_slab; ++ml) {
            const int ch0 = (c * p->slabs_per_chunk + s) * 16;
            const int b = ch0 / l.cpad;
            const int j = (ch0 - b * l.cpad) >> 4;
            uint32_t col, tile16, neg;
            int fi, fs;                                       // which `seen` flag this MMA initialises
            if (!w.fuse) {
              const int a = p->comp_of[gi][ml];
              const int img = p->op_img[b][a];
              if (img < 0) continue;
              col = (uint32_t)(ml * p->NBp);
              tile16 = (uint32_t)(((size_t)img * p->img_bytes + (size_t)j * p->slab_bytes) >> 4);
              neg = (uint32_t)p->op_neg[b][a];
              fi = ml; fs = 0;
            } else {
              const int a0 = p->comp_of[gi][F * ml], m = w.pair_xor;
              const int e0 = p->op_img[b][a0];
              if (e0 < 0) continue;
              const int st = set_of[a0][b];
              col = (uint32_t)((ml * S + st) * p->NBmma);
              tile16 = (uint32_t)(((size_t)cl::fuse_set(e0, m) * p->img_bytes + (size_t)j * p->slab_bytes) >> 4);
              neg = (uint32_t)p->op_neg[b][a0];                // a0's product enters with sign +
              fi = ml; fs = st;
              // epilogue: member t reads the 8-column slot of ITS image in this column set, with its sign relative to a0's
              for (int t = 0; t < F; ++t) {
                const int at = p->comp_of[gi][F * ml + t];
                p->epi_col[gi][F * ml + t][st] = (uint16_t)(col + cl::fuse_slot(p->op_img[b][at], m) * 8);
                p->epi_sgn[gi][F * ml + t][st] = (int8_t)((p->op_neg[b][a0] ^ p->op_neg[b][at]) ? -1 : 1);
              }
            }
            if (col > 0x1ffu || tile16 > 0x3fffu) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: op table field overflow");
            uint2 e;
            e.x = (1u << 31) | (seen[fi][fs] ? 0u : (1u << 30)) | (col << 20) | ((uint32_t)s << 16) | tile16;
            e.y = idesc | (neg << 14);
            seen[fi][fs] = true;
            dst[n++] = e;
          }
        if (n > 0) dst[n - 1].x |= 1u << 29;                 // the lane that issues last commits the stage
        for (; n < lps; ++n) dst[n] = make_uint2(0u, 0u);
      }
    }
  }
  const int acc_cols = p->acc_cols;
  p->acc_stages = acc_cols * 2 <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < acc_cols * p->acc_stages) cols <<= 1;
  p->tmem_cols = cols;
  const size_t fixed = 1024 /* barriers */ + w.total + 4096 /* slack behind the tiles */;
  const size_t budget = 218 * 1024;   // + 8 KB static (op table, output staging) stays under the 227 KB limit
  if (fixed + 2 * (size_t)p->stage_bytes > budget)
    return fail(SELDQ_ERR_UNSUPPORTED, "compact weights (%zu B as bf16 tiles) do not fit in shared memory", w.total);
  size_t ns = (budget - fixed) / p->stage_bytes;
  if (ns > (size_t)cl::kMaxStages) ns = cl::kMaxStages;
  p->nstages = (int)ns;
  *smem_bytes = ns * p->stage_bytes + fixed;
  return SELDQ_OK;
}

// (C, W, H, N) view of a CL operand, box {BK, 128, 1, 1}
static int encode_cl_map(CUtensorMap* tm, const void* data, const cl::OperandLayout& l, int w, int h, int n,
                         int box_rows) {
  const uint64_t dims[4] = {(uint64_t)l.Cp, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)l.Cp * 2, (uint64_t)l.Cp * 2 * w, (uint64_t)l.Cp * 2 * w * h};
  const uint32_t box[4] = {(uint32_t)l.BK, (uint32_t)box_rows, 1, 1};
  return encode_tensor_map(tm, data, 2, 4, dims, strides, box, l.BK == 64 ? 3 : (l.BK == 32 ? 2 : 1));
}

int launch_cl_fprop(const ConvGeom& g, const void* in_cl, const float* const* host_w, const void* packed,
                    const float* bias, float* out, void* out_f16, cudaStream_t st) {
  FpropParams p;
  size_t smem = 0;
  int rc = plan_cl_fprop(g, &p, &smem);
  if (rc) return rc;
  if (p.dense) {
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
  } else {
    if (!packed) return fail(SELDQ_ERR_INVALID, "bf16 path needs the packed weight tiles");
    if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(SELDQ_ERR_INVALID, "packed weights must be 16-byte aligned");
    p.packed = reinterpret_cast<const uint8_t*>(packed);
  }
  p.bias = bias;
  p.out = out;
  p.out16 = reinterpret_cast<__half*>(out_f16);
  p.out_sN = g.out_sN; p.out_sC = g.out_sC; p.out_sH = g.out_sH;
  const cl::OperandLayout l = cl::operand_layout(g.tab.nc, g.R, p.dense != 0);
  alignas(64) CUtensorMap tm;
  rc = encode_cl_map(&tm, in_cl, l, g.IW, g.IH, g.N, p.box_rows);
  if (rc) return rc;
  const int grid = p.total_units < cl::num_sms() ? p.total_units : cl::num_sms();
  void (*kern)(const CUtensorMap, const FpropParams) = nullptr;
  switch (p.gc) {
    case 1: kern = cl::qconv_cl_fprop_kernel<1>; break;
    case 2: kern = cl::qconv_cl_fprop_kernel<2>; break;
    case 4: kern = cl::qconv_cl_fprop_kernel<4>; break;
    case 8: kern = cl::qconv_cl_fprop_kernel<8>; break;
    default: return fail(SELDQ_ERR_UNSUPPORTED, "unexpected out-component group size %d", p.gc);
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "fprop smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(cl::kFpropThreads), smem, st, tm, p);
  if (le != cudaSuccess) return fail(SELDQ_ERR_CUDA, "qconv_cl_fprop_kernel: %s", cudaGetErrorString(le));
  return check_launch("qconv_cl_fprop_kernel");
}

}  // namespace seldq

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
#include "conv_cl_plan.h"
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

// Per-channel sums over the 32 rows a warp holds: every lane owns v[0..16) = 16 channels of ONE row.  Butterfly
// "halve the channels, double the rows": after the exchanges with lanes ^16, ^8, ^4, ^2 a lane holds the partial sum
// of ONE channel, ch = bit4 * 8 + bit3 * 4 + bit2 * 2 + bit1 of its lane number; the last exchange (^1) completes it
// in both lanes of a pair.  15 shuffles instead of 16 x 5.
__device__ __forceinline__ float warp_channel_sums(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mine = hi ? v[8 + j] : v[j], other = hi ? v[j] : v[8 + j];
      a[j] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mine = hi ? a[4 + j] : a[j], other = hi ? a[j] : a[4 + j];
      b[j] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
    }
  }
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float mine = hi ? b[2 + j] : b[j], other = hi ? b[j] : b[2 + j];
      c[j] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
    }
  }
  const bool hi = (lane & 2) != 0;
  const float mine = hi ? c[1] : c[0], other = hi ? c[0] : c[1];
  float d = mine + __shfl_xor_sync(0xffffffffu, other, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}
__device__ __forceinline__ int warp_channel_of_lane(int lane) { return (lane >> 1) & 15; }

// GLUE = the fused-glue epilogue of the fp32 output path (FpropParams::epi_mode / stats) is compiled in; the plain
// instantiations carry none of its registers or branches
// trace slots (CTA 0 only): 0 kernel entry, 1 set-up done, 2 weights resident (MMA warp), 3 first activation stage landed,
// 8 + 4 u + {0: MMA unit start, 1: MMA unit issued, 2: epilogue sees the accumulator, 3: epilogue unit done}, 7 exit
__device__ __forceinline__ void trace_stamp(const FpropParams& p, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

template <int GC, bool GLUE>
__global__ void __launch_bounds__(kFpropThreads, 1)
qconv_cl_fprop_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                      const __grid_constant__ FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  if (threadIdx.x == 0) trace_stamp(p, 0);
  // sibling launch: even CTAs serve problem 0, odd CTAs problem 1; G CTAs share a problem's units
  const int prob = p.nprob == 2 ? (int)(blockIdx.x & 1u) : 0;
  const int G = p.nprob == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int cta = p.nprob == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const CUtensorMap& tm_in = prob ? tm_in1 : tm_in0;
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
    // a stage is consumed by ONE MMA warp (its owner, whose tcgen05.commit releases the slot), but EVERY MMA warp
    // arrives on the slot's empty barrier once it has observed the stage's full phase: a slot is refilled only after
    // all of them have seen it, so a warp that lags (a profiler's instrumentation is enough) can never find the
    // barrier two phases on and mistake the refilled slot's phase for the one it still has to observe
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
  if (threadIdx.x == 0) trace_stamp(p, 1);

  if (warp == 0) {
    // ===== TMA producer ============================================================================
    if (ptx::elect_one()) {
      if (!p.dense) {
        // resident weight tiles: one bulk copy per <= 32 KB
        ptx::mbar_arrive_expect_tx(w_bar, p.w_bytes);
        for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
          const uint32_t n = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
          ptx::bulk_load(b_img + off, (prob ? p.packed1 : p.packed) + off, n, w_bar);
        }
      }
      pdl_wait();            // activations of the previous kernel from here on (the weights above never race)
      uint32_t slot = 0, parity = 0;
      for (int round = 0; round * G < p.total_units; ++round) {
        const int u = unit_of_round(round, p.total_units, G, cta);
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
    // Lane-parallel issue: the MMAs of one stage (<= 32: taps x slabs x out components of the group, structural zero
    // blocks left out) form a dense list; all lanes of a warp build their descriptors at once and the tcgen05.mma
    // under the lane predicate then costs one short elect-and-issue round per active lane (~40 cycles per MMA
    // instead of ~100 for descriptor arithmetic on the uniform datapath; tools/umma_rate.py).
    // Every accumulator starts from ZERO -- the epilogue warps clear it with tcgen05.st after reading it -- so all
    // MMAs accumulate and commute, nothing orders them, and the kMmaWarps warps take whole STAGES in turn (stage n
    // of the CTA belongs to warp n mod kMmaWarps).  Measured with the CTA timeline (tools/fprop_trace.py): when the
    // four warps shared every stage (entry i to warp i mod 4, a named barrier behind the initialising MMAs), a
    // stage cost ~0.35 us of wait / barrier / commit bookkeeping on top of its MMAs and a TCN unit took 2.7-4.4 us
    // for 1.0-2.1 us of tensor-core work; with a stage per warp that bookkeeping overlaps four ways.
    // tcgen05.commit tracks the MMAs of the executing thread; the tensor pipe retires in order, so a warp's commit
    // comes from the lane that issued last.
    {
      const int me = warp == 1 ? 0 : warp - 5;
      const uint64_t a_hi = ptx::make_smem_desc_hi(16, p.a_sbo, p.a_swz);                 // K-major, swizzled
      const uint64_t b_hi = ptx::make_smem_desc_hi((uint32_t)p.NBmma * 16u, 128, ptx::kSwizzleNone);
      const uint32_t a_base = ptx::smem_u32(a_ring), b_lo16 = ptx::smem_u32(b_img) >> 4;
      const int lanes_per_chunk = p.slabs_per_chunk * p.mma_per_slab;   // op-table entries of one (group, chunk)
      const int lanes_per_stage = lanes_per_chunk * p.tps;         // a stage holds p.tps taps (1, or all of them)
      const bool lane_on = lane < lanes_per_stage;
      const int my_tl = lane_on ? lane / lanes_per_chunk : 0;      // tap of the stage this lane's MMA belongs to
      const int my_rest = lane - my_tl * lanes_per_chunk;          // ... and its entry in the chunk's op list
      // where that tap's A operand starts inside a stage: its own box, or a row offset into the shared box
      const uint32_t my_tap_off16 = !lane_on ? 0u
                                    : p.rs ? (uint32_t)p.rs_row[my_tl] * ((uint32_t)p.BK * 2u >> 4)
                                           : ((uint32_t)my_tl * p.box_bytes) >> 4;
      if (!p.dense) ptx::mbar_wait(w_bar, 0);
      if (warp == 1 && lane == 0) trace_stamp(p, 2);
      uint32_t slot = 0, parity = 0, it = 0, sc = 0;               // sc: stages of this CTA so far
      for (int round = 0; round * G < p.total_units; ++round) {
        const int u = unit_of_round(round, p.total_units, G, cta);
        if (u < 0) continue;
        const int gi = u / p.total_tiles;
        const uint32_t mask = p.chunk_mask[p.group_order[gi]];
        const uint32_t as = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t use = p.acc_stages == 2 ? (it >> 1) : it;
        ptx::mbar_wait(&tempty_bar[as], use & 1);                  // read AND cleared by the epilogue warps
        ptx::tc_fence_after();
        if (warp == 1 && lane == 0 && it < 6) trace_stamp(p, 8 + 4 * (int)it);
        const uint32_t d_unit = tmem_base + as * acc_cols;
        const uint2* tbl_g = op_tbl_s + (size_t)p.group_order[gi] * p.chunks * lanes_per_chunk;
        int last_closer = -1;                                  // lane that issued this warp's last MMA of the unit
        for (int c = 0; c < p.chunks; ++c) {
          if (!((mask >> c) & 1u)) continue;
          for (int tap = 0; tap < p.ntaps; tap += p.tps, ++sc) {
            // EVERY warp observes every phase of every stage barrier (a parity wait tells the current phase from the
            // previous one only: a warp that skipped a lap of the ring would take an unfilled slot for a filled one);
            // for a stage that has landed this is one try_wait
            ptx::mbar_wait(&full_bar[slot], parity);
            if ((int)(sc % (uint32_t)kMmaWarps) == me) {
              // this lane's MMA of the stage
              uint2 e = make_uint2(0u, 0u);
              if (lane_on) e = tbl_g[c * lanes_per_chunk + my_rest];
              const bool valid = (int)e.x < 0;
              const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
              const uint32_t a_off16 = ((e.x >> 16) & 3u) * 2u + my_tap_off16;
              const uint32_t d_lane = d_unit + ((e.x >> 20) & 0x1ffu);
              const uint32_t b16 = b_lo16 + (e.x & 0x3fffu) + (uint32_t)(tap + my_tl) * p.tap_stride16;
              const uint64_t a_desc = a_hi | (uint64_t)((((a_base + slot * p.stage_bytes) >> 4) + a_off16) & 0x3fffu);
              const uint64_t b_desc = b_hi | (uint64_t)(b16 & 0x3fffu);
              ptx::tc_fence_after();
              if (warp == 1 && lane == 0 && it == 0 && sc == 0) trace_stamp(p, 3);
              if (valid) ptx::umma_f16(d_lane, a_desc, b_desc, e.y, 1u);
              __syncwarp();
              if (vmask != 0u) {
                const int closer = 31 - __clz((int)vmask);
                last_closer = closer;
                if (lane == closer) ptx::umma_commit(&empty_bar[slot]);
              } else if (lane == 0) {
                ptx::mbar_arrive(&empty_bar[slot]);             // nothing reads the slot
              }
              __syncwarp();
            } else {
              __syncwarp();                                     // every lane is past its wait on this phase
              if (lane == 0) ptx::mbar_arrive(&empty_bar[slot]);
            }
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
        if (warp == 1 && lane == 0 && it < 6) trace_stamp(p, 9 + 4 * (int)it);
        ++it;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global ===================================================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int eset = warp >= kFirstExtraEpiWarp ? 1 + (warp - kFirstExtraEpiWarp) / 4 : 0;
    const int row = q * 32 + lane;
    // The accumulators start from zero (every MMA accumulates: see the issuers): the warps of a lane quarter clear
    // alternate 16-column groups of all buffers here; behind a unit every warp clears the columns it has read.
    auto clear_acc = [&](uint32_t first_col, uint32_t ncols) {
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + first_col;
      for (uint32_t col = 16u * (uint32_t)eset; col < ncols; col += 16u * kEpiSets) ptx::tmem_st16_zero(t0 + col);
      ptx::tmem_st_wait();
    };
    clear_acc(0u, acc_cols * (uint32_t)p.acc_stages);
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      ptx::mbar_arrive(&tempty_bar[0]);
      ptx::mbar_arrive(&tempty_bar[1]);
    }
    pdl_wait();
    // 16-byte bf16 stores need 8-element alignment of every row start
    const bool vec16 = p.out16 != nullptr && (p.OW & 7) == 0 && (p.out_sC & 7) == 0 && (p.out_sH & 7) == 0 &&
                       (p.out_sN & 7) == 0 && (reinterpret_cast<unsigned long long>(p.out16) & 15ull) == 0;
    uint32_t it = 0;
    for (int round = 0; round * G < p.total_units; ++round) {
      const int u = unit_of_round(round, p.total_units, G, cta);
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
      const long long row_off = (long long)n * p.out_sN + (long long)h * p.out_sH + w;
      float* out_row = (prob ? p.out1 : p.out) + row_off;
      __half* out16_row = p.out16 + row_off;
      const int epi_mode = GLUE ? p.epi_mode[prob] : 0;
      const float* add_row = GLUE && p.addend[prob] ? p.addend[prob] + row_off : nullptr;
      double* stats = GLUE ? p.stats[prob] : nullptr;
      // fused glue: the tensor this warp adds to (skip sum or x) is requested into L1 one 16-channel piece ahead --
      // the first piece while the unit's MMAs are still running -- so the epilogue's loads do not pay an L2 round
      // trip per piece
      auto prefetch_piece = [&](int pn) {
        if (!GLUE || !w_ok || !(epi_mode == 1 || epi_mode == 2)) return;
        const int ppc = (p.Pc + 15) >> 4;
        if (pn >= GC * ppc) return;
        const int pal = pn / ppc, pc0 = (pn - pal * ppc) * 16;
        const float* src = (epi_mode == 1 ? out_row : add_row) + (long long)(p.comp_of[group][pal] * p.Pc + pc0) * p.out_sC;
        const int plim = p.Pc - pc0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < plim) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (long long)j * p.out_sC));
      };
      prefetch_piece(eset);
      ptx::mbar_wait(&tfull_bar[as], use & 1);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0 && it < 6) trace_stamp(p, 10 + 4 * (int)it);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols;
      int piece_no = 0;
      for (int al = 0; al < GC; ++al) {
        const int ch_base = p.comp_of[group][al] * p.Pc;
        for (int c0 = 0; c0 < p.Pc; c0 += 16, ++piece_no) {
          if (piece_no % kEpiSets != eset) continue;
          const bool tr = warp == 2 && lane == 0 && it == 0 && piece_no < 3 * kEpiSets;
          if (tr) trace_stamp(p, 40 + 4 * (piece_no / kEpiSets));
          uint32_t v[16];
          if (!p.fuse) {
            ptx::tmem_ld16(t_row + (uint32_t)(al * p.NBp + c0), v);
            ptx::tmem_ld_wait();
            ptx::tmem_st16_zero(t_row + (uint32_t)(al * p.NBp + c0));      // read once, by this warp only: clear it
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
            if (tr) trace_stamp(p, 41 + 4 * (piece_no / kEpiSets));
            // every 8-column group belongs to one (component, channel group), i.e. to this piece alone: clear it
            // for the next unit right away (no pass over the whole accumulator, no barrier between the warps)
#pragma unroll
            for (int st = 0; st < 4; ++st)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
                if (p.epi_sgn[group][al][st] != 0 && c0 + hf * 8 < p.Pc)
                  ptx::tmem_st8_zero(t_row + (uint32_t)p.epi_col[group][al][st] + (uint32_t)((c0 >> 3) + hf) * og_stride);
            const float s0 = (float)p.epi_sgn[group][al][0], s1 = (float)p.epi_sgn[group][al][1];
            const float s2 = (float)p.epi_sgn[group][al][2], s3 = (float)p.epi_sgn[group][al][3];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
              for (int j = 0; j < 8; ++j)
                v[hf * 8 + j] = __float_as_uint(fmaf(s3, __uint_as_float(u[3][hf][j]), fmaf(s2, __uint_as_float(u[2][hf][j]),
                                                fmaf(s1, __uint_as_float(u[1][hf][j]), s0 * __uint_as_float(u[0][hf][j])))));
            if (tr) trace_stamp(p, 42 + 4 * (piece_no / kEpiSets));
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
          } else if (GLUE && (epi_mode != 0 || stats != nullptr)) {
            // fused glue (conv_cl.h): running skip sum / x + residual, and the next BatchNorm's batch statistics.
            // The epilogue warps run a dependent chain at a fraction of an instruction per cycle, so the common case
            // (a full 16-channel piece, no bias) is straight-line code without per-channel predicates; loads first,
            // all in flight together, then the stores (the lines were requested into L1 one piece ahead).
            const int lim = p.Pc - c0;
            const long long off0 = (long long)(ch_base + c0) * p.out_sC;
            const bool loads = epi_mode == 1 || epi_mode == 2;
            prefetch_piece(piece_no + kEpiSets);
            float f[16];
            if (lim >= 16 && p.bias == nullptr) {
              if (w_ok) {
                float* dst = out_row + off0;
                if (loads) {
                  const float* src = (epi_mode == 1 ? out_row : add_row) + off0;
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = src[(long long)j * p.out_sC];
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] += __uint_as_float(v[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[(long long)j * p.out_sC] = f[j];
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = 0.f;
              if (w_ok && loads) {
                const float* src = (epi_mode == 1 ? out_row : add_row) + off0;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < lim) f[j] = src[(long long)j * p.out_sC];
              }
              if (w_ok) {
                float* dst = out_row + off0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (j < lim) {
                    f[j] += __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + ch_base + c0 + j) : 0.f);
                    *dst = f[j];
                  }
                  dst += p.out_sC;
                }
              }
            }
            if (stats != nullptr) {                 // warp-uniform: every lane takes part in the shuffles
              float q[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) q[j] = f[j] * f[j];
              const float s1 = warp_channel_sums(f, lane), s2 = warp_channel_sums(q, lane);
              const int j = warp_channel_of_lane(lane);
              if (j < lim) atomicAdd(stats + 2 * (ch_base + c0 + j) + (lane & 1), (double)((lane & 1) ? s2 : s1));
            }
          } else if (w_ok) {
            float* dst = out_row + (long long)(ch_base + c0) * p.out_sC;
            const int lim = p.Pc - c0;
            if (lim >= 16 && p.bias == nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) { *dst = __uint_as_float(v[j]); dst += p.out_sC; }
              if (tr) trace_stamp(p, 43 + 4 * (piece_no / kEpiSets));
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
      ptx::tmem_st_wait();                         // this warp's clears of the columns it read have landed
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (warp == 2 && lane == 0 && it < 6) trace_stamp(p, 11 + 4 * (int)it);
      ++it;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_stamp(p, 7);
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

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
  uint8_t* dst = p.dst + pack_dst_offset(p, img, tap, j, kc, n);
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

static unsigned long long* g_fprop_trace = nullptr;
void set_fprop_trace(void* dev_buf) { g_fprop_trace = reinterpret_cast<unsigned long long*>(dev_buf); }
unsigned long long* fprop_trace() { return g_fprop_trace; }

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
using cl::WeightPlan;
using cl::weight_plan;

cl::OperandLayout x_operand_layout(const ConvGeom& fwd) {
  return cl::operand_layout(fwd.tab.nc, fwd.tab.nc * fwd.Ic, fwd.tab.nc == 1 || fwd.Ic < 8);
}
cl::OperandLayout gy_operand_layout(const ConvGeom& fwd) {
  return cl::operand_layout(fwd.tab.nc, fwd.tab.nc * fwd.Oc, fwd.tab.nc == 1 || fwd.Oc < 8);
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
  cl::fill_pack_params(g, w, &p);
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
    cl::fill_pack_params(g, w, &p);
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
  return cl::plan_fprop(g, p, smem_bytes, cl::num_sms());
}

// (C, W, H, N) view of a CL operand, box {BK, 128, 1, 1}
static int encode_cl_map(CUtensorMap* tm, const void* data, const cl::OperandLayout& l, int w, int h, int n,
                         int box_rows) {
  const uint64_t dims[4] = {(uint64_t)l.Cp, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  const uint64_t strides[3] = {(uint64_t)l.Cp * 2, (uint64_t)l.Cp * 2 * w, (uint64_t)l.Cp * 2 * w * h};
  const uint32_t box[4] = {(uint32_t)l.BK, (uint32_t)box_rows, 1, 1};
  return encode_tensor_map(tm, data, 2, 4, dims, strides, box, l.BK == 64 ? 3 : (l.BK == 32 ? 2 : 1));
}

typedef void (*FpropKernel)(const CUtensorMap, const CUtensorMap, const FpropParams);
static int fprop_kernel_for(int gc, bool glue, FpropKernel* kern) {
  switch (gc) {
    case 1: *kern = glue ? cl::qconv_cl_fprop_kernel<1, true> : cl::qconv_cl_fprop_kernel<1, false>; break;
    case 2: *kern = glue ? cl::qconv_cl_fprop_kernel<2, true> : cl::qconv_cl_fprop_kernel<2, false>; break;
    case 4: *kern = glue ? cl::qconv_cl_fprop_kernel<4, true> : cl::qconv_cl_fprop_kernel<4, false>; break;
    case 8: *kern = glue ? cl::qconv_cl_fprop_kernel<8, true> : cl::qconv_cl_fprop_kernel<8, false>; break;
    default: return fail(SELDQ_ERR_UNSUPPORTED, "unexpected out-component group size %d", gc);
  }
  return SELDQ_OK;
}
static bool wants_glue(const FpropParams& p) {
  return p.epi_mode[0] != 0 || p.epi_mode[1] != 0 || p.stats[0] != nullptr || p.stats[1] != nullptr;
}

static int set_fprop_epilogue(FpropParams& p, int prob, const cl::FpropEpilogue* e, const float* out) {
  if (!e) return SELDQ_OK;
  if (e->mode < 0 || e->mode > 3) return fail(SELDQ_ERR_INVALID, "convolution epilogue mode %d", e->mode);
  if (e->mode == 2 && !e->addend) return fail(SELDQ_ERR_INVALID, "convolution epilogue mode 2 needs the addend tensor");
  if (e->mode == 2 && e->addend == out) return fail(SELDQ_ERR_INVALID, "the addend of a convolution epilogue may not alias its output");
  p.epi_mode[prob] = e->mode;
  p.addend[prob] = e->mode == 2 ? e->addend : nullptr;
  p.stats[prob] = e->stats;
  return SELDQ_OK;
}

int launch_cl_fprop(const ConvGeom& g, const void* in_cl, const float* const* host_w, const void* packed,
                    const float* bias, float* out, void* out_f16, cudaStream_t st, const cl::FpropEpilogue* epi) {
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
  if (epi && (epi->mode != 0 || epi->stats) && out_f16)
    return fail(SELDQ_ERR_UNSUPPORTED, "fused convolution epilogues exist for the fp32 output only");
  if ((rc = set_fprop_epilogue(p, 0, epi, out))) return rc;
  p.trace = cl::fprop_trace();
  const cl::OperandLayout l = cl::operand_layout(g.tab.nc, g.R, p.dense != 0);
  alignas(64) CUtensorMap tm;
  rc = encode_cl_map(&tm, in_cl, l, g.IW, g.IH, g.N, p.box_rows);
  if (rc) return rc;
  const int grid = p.total_units < cl::num_sms() ? p.total_units : cl::num_sms();
  FpropKernel kern = nullptr;
  if ((rc = fprop_kernel_for(p.gc, wants_glue(p), &kern))) return rc;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "fprop smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(cl::kFpropThreads), smem, st, tm, tm, p);
  if (le != cudaSuccess) return fail(SELDQ_ERR_CUDA, "qconv_cl_fprop_kernel: %s", cudaGetErrorString(le));
  return check_launch("qconv_cl_fprop_kernel");
}

int plan_cl_fprop_pair(const ConvGeom& g, FpropParams* p, size_t* smem_bytes) {
  return cl::plan_fprop(g, p, smem_bytes, cl::num_sms(), 2);
}

int launch_cl_fprop_pair(const ConvGeom& g, const void* const in_cl[2], const void* const packed[2], float* const out[2],
                         const cl::FpropEpilogue epi[2], cudaStream_t st) {
  FpropParams p;
  size_t smem = 0;
  int rc = plan_cl_fprop_pair(g, &p, &smem);
  if (rc) return rc;
  if (p.dense) return fail(SELDQ_ERR_UNSUPPORTED, "sibling launches serve layers with >= 8 channels per component");
  for (int k = 0; k < 2; ++k) {
    if (!in_cl[k] || !packed[k] || !out[k]) return fail(SELDQ_ERR_INVALID, "sibling convolution launch: null pointer");
    if (reinterpret_cast<uintptr_t>(packed[k]) & 15) return fail(SELDQ_ERR_INVALID, "packed weights must be 16-byte aligned");
  }
  if (out[0] == out[1]) return fail(SELDQ_ERR_INVALID, "sibling convolutions need distinct outputs");
  p.packed = reinterpret_cast<const uint8_t*>(packed[0]);
  p.packed1 = reinterpret_cast<const uint8_t*>(packed[1]);
  p.out = out[0];
  p.out1 = out[1];
  p.out_sN = g.out_sN; p.out_sC = g.out_sC; p.out_sH = g.out_sH;
  for (int k = 0; k < 2; ++k)
    if ((rc = set_fprop_epilogue(p, k, epi ? &epi[k] : nullptr, out[k]))) return rc;
  p.trace = cl::fprop_trace();
  const cl::OperandLayout l = cl::operand_layout(g.tab.nc, g.R, false);
  alignas(64) CUtensorMap tm[2];
  for (int k = 0; k < 2; ++k)
    if ((rc = encode_cl_map(&tm[k], in_cl[k], l, g.IW, g.IH, g.N, p.box_rows))) return rc;
  const int per = cl::num_sms() / 2;
  const int grid = 2 * (p.total_units < per ? p.total_units : per);
  FpropKernel kern = nullptr;
  if ((rc = fprop_kernel_for(p.gc, wants_glue(p), &kern))) return rc;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "fprop smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(cl::kFpropThreads), smem, st, tm[0], tm[1], p);
  if (le != cudaSuccess) return fail(SELDQ_ERR_CUDA, "qconv_cl_fprop_kernel (pair): %s", cudaGetErrorString(le));
  return check_launch("qconv_cl_fprop_kernel");
}

}  // namespace seldq

// tcgen05 weight-gradient kernel for the quaternion / dual-quaternion convolutions (channels-last x).
//
//   dWq[(a,o), (b,i), tap] = sum_{n,h,w} gy[n, a*Oc+o, h, w] * x[n, b*Ic+i, h + off_h(tap), w + off_w(tap)]
//   gW_e[o, i, tap]        = sum_{(a,b): widx[a][b] = e} sign[a][b] * dWq[(a,o), (b,i), tap]     
//
// The contraction runs over positions.  A = gy from its pitched NCHW bf16 copy: time is contiguous, so a TMA
// box [OS channels x 64 t] is a K-major operand and the rows of the M = 128 tile are gathered per component
// (all components a x OS out channels).  B = x from its channels-last operand: a box [64 t x 64 ch] is an
// MN-major operand, all Cp padded input channels form the N extent (one or two MMAs of <= 256 columns), and
// the convolution tap is just an offset on the box's w / h coordinates.  One CTA owns (OS out channels of
// every component) x (all in channels) x (a group of taps) and reduces a contiguous slice of the position
// axis (split-K).  The dense 128 x Cp accumulator lives in TMEM; the epilogue folds it onto the COMPACT
// gradients inside the CTA (sign-weighted sum over the (a,b) pairs of each compact tensor, through shared
// memory) and adds the result with one atomicAdd per compact element -- the expanded gradient never
// reaches HBM.
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "conv_cl_plan.h"
#include "launch.h"
#include "pdl.cuh"
#include "tensor_map.h"
#include "umma_ptx.cuh"

namespace seldq {
namespace cl {

// trace slots (CTA (0, 0) only): 0 entry, 1 set-up done, 2 first stage requested, 3 first stage landed, 4 last MMA issued,
// 5 epilogue sees the accumulator, 6 fold done, 7 exit; 8 + ks: MMA warp saw stage ks (first 24 K steps)
__device__ __forceinline__ void wtrace(const WgradParams& p, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[slot] = t;
  }
}

// one (out channel, in channel) position of the staged 128 x (ncomp * 16) block folded onto the compact tensors:
// gw_e[off] += sum_k sign(e, k) * block(a(e, k), b(e, k)); table entries beyond a tensor's pair count have sign 0.
// Two tensors per round: their 2 * KMAX table entries, then their 2 * KMAX values, are in flight together (compile-time
// trip counts, no predicate between the loads) while the register footprint stays small -- the kernel's CTAs share
// their SMs with the glue kernels of the main stream, which need the registers (all eight tensors in one round
// took 121 registers per thread and cost the training step 0.06 ms).
template <int NW, int KMAX>
__device__ __forceinline__ void fold_block(const float* col, const int2* tbl, float* const* gw, long long off) {
  constexpr int EB = NW >= 2 ? 2 : 1;
#pragma unroll 1
  for (int e0 = 0; e0 < NW; e0 += EB) {
    int2 f[EB][KMAX];
#pragma unroll
    for (int e = 0; e < EB; ++e)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) f[e][k] = tbl[(e0 + e) * 8 + k];
    float v[EB][KMAX];
#pragma unroll
    for (int e = 0; e < EB; ++e)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) v[e][k] = col[f[e][k].x];
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) acc = fmaf(__int_as_float(f[e][k].y), v[e][k], acc);
      atomicAdd(gw[e0 + e] + off, acc);
    }
  }
}

// (register cap of 64: one CTA per SM by shared memory, but the glue kernels of the main stream co-reside in what is left)
__global__ void __launch_bounds__(kThreads, 3)
qconv_cl_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g0, const __grid_constant__ CUtensorMap tm_g1,
                      const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  if (threadIdx.x == 0) wtrace(p, 0);
  __shared__ __align__(8) uint64_t full_bar[kWgradStages], empty_bar[kWgradStages], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int2 fold_tbl[64];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // tile decode: blockIdx.x = (problem * tap_groups + tap_group) * o_tiles + o_tile, blockIdx.y = split
  const int o_tile = blockIdx.x % p.o_tiles;
  const int prob = (blockIdx.x / p.o_tiles) / p.tap_groups;
  const int tg = (blockIdx.x / p.o_tiles) % p.tap_groups;
  const CUtensorMap* tm_g = prob ? &tm_g1 : &tm_g0;
  const int tap0 = tg * p.taps_per_group;
  const int ntap = min(p.taps_per_group, p.ntaps - tap0);
  const int o0 = o_tile * p.OS;
  const long long per = (p.ksteps + p.splits - 1) / p.splits;
  const long long k_begin = (long long)blockIdx.y * per;
  const long long k_end = min(p.ksteps, k_begin + per);
  const int nk = (int)max(0LL, k_end - k_begin);

  const uint32_t a_bytes = 128u * 128u;                    // [128 rows x 64 t] bf16
  const int nstages = p.nstages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(tm_g);
    ptx::prefetch_tensormap(&tm_x);
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                   // operands and the gradient buffers belong to earlier kernels up to here
  if (threadIdx.x == 0) wtrace(p, 1);

  if (nk > 0) {
    if (warp == 0) {
      if (ptx::elect_one()) {
        uint32_t slot = 0, parity = 0;
        for (int ks = 0; ks < nk; ++ks) {
          long long u = k_begin + ks;
          const int wc = (int)(u % p.chunks_w); u /= p.chunks_w;
          const int h = (int)(u % p.OH);
          const int n = (int)(u / p.OH);
          const int w0 = wc * 64;
          ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[slot], a_bytes + (uint32_t)ntap * p.b_tap_bytes);
          uint8_t* st = smem + (size_t)slot * p.stage_bytes;
          // one box per operand and tap: the component / channel-chunk axes are box dimensions of rank-5 maps (a stage
          // was 8 + 6 boxes of 128-byte rows before; the K loop is bound by the latency of its TMA traffic)
          ptx::tma_load_5d(st, tm_g, &full_bar[slot], w0, h, o0, 0, n);
          for (int t = 0; t < ntap; ++t) {
            if (p.x_one_box)
              ptx::tma_load_5d(st + a_bytes + (size_t)t * p.b_tap_bytes, &tm_x, &full_bar[slot], 0,
                               w0 + p.off_w[tap0 + t], 0, h + p.off_h[tap0 + t], n);
            else
              for (int c = 0; c < p.nchunks; ++c)
                ptx::tma_load_4d(st + a_bytes + (size_t)t * p.b_tap_bytes + (size_t)c * 8192, &tm_x, &full_bar[slot],
                                 c * 64, w0 + p.off_w[tap0 + t], h + p.off_h[tap0 + t], n);
          }
          if (ks == 0) wtrace(p, 2);
          if (++slot == (uint32_t)nstages) { slot = 0; parity ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
        // A: K-major, 128B swizzle: 8-row groups 1024 B apart (SBO), K advances inside the swizzled row
        // B: MN-major, 128B swizzle: 64-channel chunks 8192 B apart (LBO), 8 t-rows 1024 B apart (SBO),
        //    a K = 16 slab is 16 rows = 2048 B
        const uint64_t a_hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
        const uint64_t b_hi = ptx::make_smem_desc_hi(8192, 1024, ptx::kSwizzle128B);
        const uint32_t base = ptx::smem_u32(smem);
        uint32_t slot = 0, parity = 0;
        for (int ks = 0; ks < nk; ++ks) {
          ptx::mbar_wait(&full_bar[slot], parity);
          ptx::tc_fence_after();
          if (ks == 0) wtrace(p, 3);
          if (ks < 24) wtrace(p, 8 + ks);
          const uint32_t st = base + slot * p.stage_bytes;
          for (int t = 0; t < ntap; ++t)
            for (int n0 = 0; n0 < p.Cp; n0 += 256) {
              const uint32_t nn = (uint32_t)min(256, p.Cp - n0);
              const uint32_t idesc = ptx::make_idesc_bf16(128, nn, 0, 1, 0, 0);
              const uint32_t b0 = st + a_bytes + (uint32_t)t * p.b_tap_bytes + (uint32_t)(n0 / 64) * 8192u;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16(tmem_base + (uint32_t)(t * p.Cp + n0), ptx::smem_desc(a_hi, st + k * 32u),
                              ptx::smem_desc(b_hi, b0 + k * 2048u), idesc, (ks > 0 || k > 0) ? 1u : 0u);
            }
          ptx::umma_commit(&empty_bar[slot]);
          if (++slot == (uint32_t)nstages) { slot = 0; parity ^= 1; }
        }
        ptx::umma_commit(&done_bar);
        wtrace(p, 4);
      }
    } else {
      // ===== epilogue ==============================================================================
      // Measured (tools/wgrad_bench.py: time versus split-K factor): a launch costs ~21 us of fixed time and only
      // 0.5 us per K step, and most of the fixed time was this fold when four warps walked a per-target loop with
      // run-time divisions.  Now eight warps: each stages the columns of half of the in components of its lane
      // quarter, and every thread owns one (out channel, in channel) position of the 16 x 16 block and folds it for
      // all compact tensors -- the (a, b, sign) table is then uniform across the warp.
      const int q = warp & 3, hf = (warp - 2) >> 2;
      const int row = q * 32 + lane;                 // accumulator row = TMEM lane = (a, ol)
      const int et = threadIdx.x - 64;               // 0..255 among the epilogue threads
      const int il = et & 15;
      // fold table: entry (e, k) = {offset of block (a, b) inside the staging tile, +-1.f | 0.f for an unused entry}.
      // Measured (tools/wgrad_trace.py): with the (a, b, sign) bytes read from the parameter block inside a rolled loop
      // the fold was a serial chain of constant-bank and shared-memory latencies, 8 us of a 17 us launch; with the
      // table in shared memory and the loops unrolled all loads of a fold are in flight together.
      const int pitch = p.ncomp * 16 + 1;
      if (et < 64) {
        const int e = et >> 3, k = et & 7;
        const bool on = e < p.g.tab.nw && k < p.pair_n[e];
        fold_tbl[et] = make_int2(on ? p.pair_a[e][k] * p.OS * pitch + p.pair_b[e][k] * 16 : 0,
                                 on ? __float_as_int(p.pair_neg[e][k] ? -1.f : 1.f) : 0);
      }
      int kmax = 0;
      for (int e = 0; e < p.g.tab.nw; ++e) kmax = max(kmax, (int)p.pair_n[e]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ptx::mbar_wait(&done_bar, 0);
      ptx::tc_fence_after();
      if (et == 0) wtrace(p, 5);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const ConvGeom& g = p.g;
      // all MMAs have retired: the operand ring is free and is reused as a [128][ncomp*16 + 1] fp32 staging tile
      float* stg = reinterpret_cast<float*>(smem);
      const int bper = (p.ncomp + 1) >> 1;
      const int b_lo = hf * bper, b_hi = min(p.ncomp, b_lo + bper);
      for (int t = 0; t < ntap; ++t)
        for (int ic = 0; ic < p.cpad_in; ic += 16) {
          for (int b0 = b_lo; b0 < b_hi; b0 += 2) {               // two 16-column loads in flight per wait
            uint32_t v[2][16];
            const bool two = b0 + 1 < b_hi;
            ptx::tmem_ld16(t_row + (uint32_t)(t * p.Cp + b0 * p.cpad_in + ic), v[0]);
            if (two) ptx::tmem_ld16(t_row + (uint32_t)(t * p.Cp + (b0 + 1) * p.cpad_in + ic), v[1]);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) stg[row * pitch + b0 * 16 + j] = __uint_as_float(v[0][j]);
            if (two) {
#pragma unroll
              for (int j = 0; j < 16; ++j) stg[row * pitch + (b0 + 1) * 16 + j] = __uint_as_float(v[1][j]);
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (ic + il < g.Ic) {
            for (int ol = et >> 4; ol < p.OS; ol += 16) {
              if (o0 + ol >= g.Oc) break;
              const float* col = stg + ol * pitch + il;
              const long long off = (long long)(o0 + ol) * g.wsO + (long long)(ic + il) * g.wsI + (long long)(tap0 + t) * g.wsT;
              if (g.tab.nw == 8 && kmax <= 6) fold_block<8, 6>(col, fold_tbl, p.gw[prob], off);          // dual quaternion
              else if (g.tab.nw == 4 && kmax <= 4) fold_block<4, 4>(col, fold_tbl, p.gw[prob], off);     // quaternion
              else if (g.tab.nw == 1 && kmax <= 1) fold_block<1, 1>(col, fold_tbl, p.gw[prob], off);     // real
              else {
                for (int e = 0; e < g.tab.nw; ++e) {
                  float acc = 0.f;
                  for (int k = 0; k < kmax; ++k) {
                    const int2 f = fold_tbl[e * 8 + k];
                    acc = fmaf(__int_as_float(f.y), col[f.x], acc);
                  }
                  atomicAdd(p.gw[prob][e] + off, acc);
                }
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      if (et == 0) wtrace(p, 6);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) wtrace(p, 7);
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace cl

int launch_cl_wgrad(const ConvGeom& g, const void* x_cl, const void* gy_nchw16, float* const* host_gw,
                    cudaStream_t st, const void* gy2_nchw16, float* const* host_gw2) {
  using namespace cl;
  WgradParams p;
  size_t smem = 0;
  {
    const int rc = plan_wgrad(g, (gy2_nchw16 && host_gw2) ? 2 : 1, num_sms(), &p, &smem);
    if (rc) return rc;
  }
  const int nc = g.tab.nc;
  const OperandLayout lx = x_operand_layout(g);
  p.trace = cl::fprop_trace();
  for (int i = 0; i < g.tab.nw; ++i) {
    p.gw[0][i] = host_gw[i];
    p.gw[1][i] = p.nprob == 2 ? host_gw2[i] : nullptr;
  }

  // gy: (pitch, H, C, N) view of the pitched NCHW bf16 copy, box {64 t, 1, OS channels, 1}
  alignas(64) CUtensorMap tm_g, tm_g2, tm_x;
  for (int k = 0; k < 2; ++k) {
    const uint64_t pitch = (uint64_t)nchw16_pitch(g.OW);
    // (t, h, channel inside a component, component, n): box {64 t, 1, OS channels, all components, 1}
    const uint64_t dims[5] = {pitch, (uint64_t)g.OH, (uint64_t)g.Oc, (uint64_t)nc, (uint64_t)g.N};
    const uint64_t strides[4] = {pitch * 2, pitch * g.OH * 2, pitch * g.OH * g.Oc * 2, pitch * g.OH * g.P * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)p.OS, (uint32_t)nc, 1};
    const int rc = encode_tensor_map(k ? &tm_g2 : &tm_g, (k && p.nprob == 2) ? gy2_nchw16 : gy_nchw16, 2, 5, dims, strides,
                                     box, 3);
    if (rc) return rc;
  }
  {
    // (channel inside a 64-channel chunk, w, chunk, h, n): box {64 ch, 64 t, all chunks, 1, 1} lands as nchunks
    // [64 t x 128 B] tiles, the layout the six separate boxes produced
    const uint64_t dims[5] = {64, (uint64_t)g.IW, (uint64_t)p.nchunks, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[4] = {(uint64_t)lx.Cp * 2, 128, (uint64_t)lx.Cp * 2 * g.IW, (uint64_t)lx.Cp * 2 * g.IW * g.IH};
    const uint32_t box[5] = {64, 64, (uint32_t)p.nchunks, 1, 1};
    p.x_one_box = getenv("SELDQ_WGRAD_X_BOXES") == nullptr && encode_tensor_map(&tm_x, x_cl, 2, 5, dims, strides, box, 3) == SELDQ_OK;
    int rc = SELDQ_OK;
    if (!p.x_one_box) {       // the driver refused the map whose chunk stride is smaller than its w stride: one box per chunk
      const uint64_t dims4[4] = {(uint64_t)lx.Cp, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
      const uint64_t strides4[3] = {(uint64_t)lx.Cp * 2, (uint64_t)lx.Cp * 2 * g.IW, (uint64_t)lx.Cp * 2 * g.IW * g.IH};
      const uint32_t box4[4] = {64, 64, 1, 1};
      rc = encode_tensor_map(&tm_x, x_cl, 2, 4, dims4, strides4, box4, 3);
    }
    if (rc) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(qconv_cl_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "wgrad smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)(p.nprob * p.tap_groups * p.o_tiles), (unsigned)p.splits);
  const cudaError_t le = launch_pdl(qconv_cl_wgrad_kernel, grid, dim3(kThreads), smem, st, tm_g, tm_g2, tm_x, p);
  if (le != cudaSuccess) return fail(SELDQ_ERR_CUDA, "qconv_cl_wgrad_kernel: %s", cudaGetErrorString(le));
  return check_launch("qconv_cl_wgrad_kernel");
}

}  // namespace seldq

// tcgen05 weight-gradient kernel for the quaternion / dual-quaternion convolutions.
//
//   dWq[(a,o), (b,i), tap] = sum_{n,h,w} gy[n, a*Oc+o, h, w] * x[n, b*Ic+i, h + off_h(tap), w + off_w(tap)]
//   gW_e[o, i, tap]        = sum_{(a,b): widx[a][b] = e} sign[a][b] * dWq[(a,o), (b,i), tap]     
//
// The contraction runs over time, which is the contiguous axis of both tensors, so both operands
// are K-major: TMA boxes [64 t x rows] land directly in the 128B-swizzled canonical layout.
// One CTA owns  M = 128 rows = (all components a) x (OS out channels)  and, for a group of taps,
// N = (all components b) x (IS in channels) columns per tap, and reduces a contiguous slice of the
// (n, h, w-chunk) axis (split-K).  The dense 128 x N accumulator lives in TMEM; the epilogue folds
// it onto the COMPACT gradients inside the CTA (sign-weighted sum over the (a,b) pairs of each
// compact tensor, through shared memory) and adds the result with one atomicAdd per compact
// element -- the expanded gradient never reaches HBM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "launch.h"
#include "tensor_map.h"
#include "umma_ptx.cuh"

namespace seldq {
namespace umma {

struct WgradPairs {           // which (a,b) blocks feed compact tensor e
  int8_t n[8];
  int8_t a[8][8], b[8][8], neg[8][8];
};

__global__ void __launch_bounds__(kThreads, 1)
qconv_umma_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x,
                        const __grid_constant__ WgradParams p, const __grid_constant__ WgradPairs pairs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kWgradStages], empty_bar[kWgradStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode: blockIdx.x = ((tap_group * o_tiles + o_tile) * i_tiles + i_tile), blockIdx.y = split
  int r = blockIdx.x;
  const int i_tile = r % p.i_tiles; r /= p.i_tiles;
  const int o_tile = r % p.o_tiles;
  const int tg = r / p.o_tiles;
  const int tap0 = tg * p.taps_per_group;
  const int ntap = min(p.taps_per_group, p.ntaps - tap0);
  const int o0 = o_tile * p.OS, i0 = i_tile * p.IS;
  const long long per = (p.ksteps + p.splits - 1) / p.splits;
  const long long k_begin = (long long)blockIdx.y * per;
  const long long k_end = min(p.ksteps, k_begin + per);
  const int nk = (int)max(0LL, k_end - k_begin);

  const uint32_t a_bytes = 128u * 128u;                    // [128 rows x 64 t] bf16
  const uint32_t b_bytes = (uint32_t)p.NW * 128u;          // per tap
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.taps_per_group * b_bytes;
  const int nstages = p.nstages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_g);
    ptx::prefetch_tensormap(&tm_x);
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

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
          ptx::mbar_arrive_expect_tx(&full_bar[slot], a_bytes + (uint32_t)ntap * b_bytes);
          uint8_t* st = smem + (size_t)slot * stage_bytes;
          for (int a = 0; a < p.ncomp; ++a)
            ptx::tma_load_5d(st + (size_t)a * p.OS * 128, &tm_g, &full_bar[slot], w0, h,
                             p.dense ? o0 : a * p.g.Oc + o0, n, 0);
          // off_w already contains the tap's mirror shift: the inner coordinate is a multiple of 8
          for (int t = 0; t < ntap; ++t)
            for (int b = 0; b < p.ncomp; ++b)
              ptx::tma_load_5d(st + a_bytes + (size_t)t * b_bytes + (size_t)b * p.IS * 128, &tm_x, &full_bar[slot],
                               w0 + p.off_w[tap0 + t], h + p.off_h[tap0 + t], p.dense ? i0 : b * p.g.Ic + i0, n,
                               p.tap_sidx[tap0 + t]);
          if (++slot == (uint32_t)nstages) { slot = 0; parity ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
        // K-major, 128B swizzle: 8-row groups are 1024 B apart (SBO); K advances inside the swizzled row
        const uint64_t hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
        const uint32_t idesc = ptx::make_idesc_bf16(128, (uint32_t)p.NW, 0, 0, 0, 0);
        const uint32_t base = ptx::smem_u32(smem);
        uint32_t slot = 0, parity = 0;
        for (int ks = 0; ks < nk; ++ks) {
          ptx::mbar_wait(&full_bar[slot], parity);
          ptx::tc_fence_after();
          const uint32_t st = base + slot * stage_bytes;
          for (int t = 0; t < ntap; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_base + (uint32_t)(t * p.NW), ptx::smem_desc(hi, st + k * 32u),
                            ptx::smem_desc(hi, st + a_bytes + (uint32_t)t * b_bytes + k * 32u), idesc,
                            (ks > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(&empty_bar[slot]);
          if (++slot == (uint32_t)nstages) { slot = 0; parity ^= 1; }
        }
        ptx::umma_commit(&done_bar);
      }
    } else {
      // ===== epilogue ==============================================================================
      const int q = warp & 3;
      const int row = q * 32 + lane;                 // accumulator row = TMEM lane
      const int et = threadIdx.x - 64;               // 0..127 among the epilogue threads
      ptx::mbar_wait(&done_bar, 0);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
      const ConvGeom& g = p.g;
      if (p.dense) {
        const int co = o0 + row;
        const bool row_ok = co < g.P;
        const int a = row_ok ? co / g.Oc : 0, o = row_ok ? co - a * g.Oc : 0;
        for (int t = 0; t < ntap; ++t)
          for (int c0 = 0; c0 < p.NW; c0 += 8) {
            uint32_t v[8];
            ptx::tmem_ld8(t_row + (uint32_t)(t * p.NW + c0), v);   // warp-collective: never under divergence
            ptx::tmem_ld_wait();
            if (!row_ok) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int ci = i0 + c0 + j;
              if (ci < g.R) {
                const int b = ci / g.Ic, i = ci - b * g.Ic;
                const int e = g.tab.widx[a][b];
                if (e >= 0) {
                  const float val = __uint_as_float(v[j]);
                  atomicAdd(p.gw[e] + (long long)o * g.wsO + (long long)i * g.wsI + (long long)(tap0 + t) * g.wsT,
                            g.tab.sign[a][b] > 0 ? val : -val);
                }
              }
            }
          }
      } else {
        // all MMAs have retired: the operand ring is free and is reused as a [128][NW+1] fp32 staging tile
        float* stg = reinterpret_cast<float*>(smem);
        const int pitch = p.NW + 1;
        const int targets = g.tab.nw * p.OS * p.IS;
        for (int t = 0; t < ntap; ++t) {
          for (int c0 = 0; c0 < p.NW; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(t_row + (uint32_t)(t * p.NW + c0), v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) stg[row * pitch + c0 + j] = __uint_as_float(v[j]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int tg_i = et; tg_i < targets; tg_i += 128) {
            int rr = tg_i;
            const int il = rr % p.IS; rr /= p.IS;
            const int ol = rr % p.OS;
            const int e = rr / p.OS;
            if (o0 + ol < g.Oc && i0 + il < g.Ic) {
              float acc = 0.f;
              for (int k = 0; k < pairs.n[e]; ++k) {
                const float val = stg[(pairs.a[e][k] * p.OS + ol) * pitch + pairs.b[e][k] * p.IS + il];
                acc += pairs.neg[e][k] ? -val : val;
              }
              atomicAdd(p.gw[e] + (long long)(o0 + ol) * g.wsO + (long long)(i0 + il) * g.wsI +
                            (long long)(tap0 + t) * g.wsT, acc);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

}  // namespace umma

int launch_umma_wgrad(const ConvGeom& g, const MirrorSet& x, const MirrorSet& gy, float* const* host_gw,
                      cudaStream_t st) {
  using namespace umma;
  if (g.sh != 1 || g.sw != 1) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 tensor-core path implements stride 1 only");
  const int ntaps = g.KH * g.KW;
  if (ntaps > kMaxTaps) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most %d taps", kMaxTaps);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  WgradPairs pairs;
  memset(&pairs, 0, sizeof(pairs));
  p.g = g;
  for (int i = 0; i < g.tab.nw; ++i) p.gw[i] = host_gw[i];
  const int nc = g.tab.nc;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    p.off_h[t] = (t / g.KW) * g.dh - g.ph;
    p.off_w[t] = (t % g.KW) * g.dw - g.pw;
  }
  p.OH = g.OH; p.OW = g.OW; p.N = g.N;
  if (g.Ic % 8 == 0 && g.Oc % 8 == 0) {
    p.dense = 0;
    p.ncomp = nc;
    p.OS = 128 / nc; p.IS = 128 / nc;
    if (p.IS > ((g.Ic + 15) / 16) * 16) p.IS = ((g.Ic + 15) / 16) * 16;   // narrow layers: fewer wasted columns
    p.NW = nc * p.IS;
    p.o_tiles = (g.Oc + p.OS - 1) / p.OS;
    p.i_tiles = (g.Ic + p.IS - 1) / p.IS;
    for (int a = 0; a < nc; ++a)
      for (int b = 0; b < nc; ++b) {
        const int e = g.tab.widx[a][b];
        if (e < 0) continue;
        const int k = pairs.n[e]++;
        pairs.a[e][k] = (int8_t)a; pairs.b[e][k] = (int8_t)b; pairs.neg[e][k] = (int8_t)(g.tab.sign[a][b] < 0);
      }
  } else if (g.R <= 64) {
    p.dense = 1;
    p.ncomp = 1;
    p.OS = 128;
    p.IS = ((g.R + 15) / 16) * 16;
    p.NW = p.IS;
    p.o_tiles = (g.P + 127) / 128;
    p.i_tiles = 1;
  } else {
    return fail(SELDQ_ERR_UNSUPPORTED,
                "bf16 wgrad needs channels per component to be a multiple of 8, or <= 64 input channels; got %d -> %d",
                g.R, g.P);
  }
  p.taps_per_group = 512 / p.NW;
  if (p.taps_per_group > ntaps) p.taps_per_group = ntaps;
  // keep at least two pipeline stages inside 200 KB
  while (p.taps_per_group > 1 && 2 * (128 * 128 + p.taps_per_group * p.NW * 128) > 200 * 1024) --p.taps_per_group;
  p.tap_groups = (ntaps + p.taps_per_group - 1) / p.taps_per_group;
  int cols = 32;
  while (cols < p.taps_per_group * p.NW) cols <<= 1;
  p.tmem_cols = cols;
  p.chunks_w = (g.OW + 63) / 64;
  p.ksteps = (long long)g.N * g.OH * p.chunks_w;
  const int tiles = p.tap_groups * p.o_tiles * p.i_tiles;
  long long splits = (2LL * num_sms() + tiles - 1) / tiles;
  if (splits > p.ksteps) splits = p.ksteps;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  const size_t stage_bytes = 128 * 128 + (size_t)p.taps_per_group * p.NW * 128;
  size_t ns = (200 * 1024) / stage_bytes;
  if (ns > (size_t)kWgradStages) ns = kWgradStages;
  p.nstages = (int)ns;
  size_t smem = ns * stage_bytes;
  const size_t stg = (size_t)128 * (p.NW + 1) * 4;
  if (smem < stg) smem = stg;
  smem += 8192;   // slack: the tensor core may fetch past the logical end of the last operand tile

  // x is read at the taps' offsets through its shifted mirrors, gy unshifted (mirror 0)
  for (int t = 0; t < ntaps; ++t) {
    const int s = ((-p.off_w[t]) % 8 + 8) % 8;
    int idx = -1;
    for (int i = 0; i < x.nshifts; ++i)
      if (x.shifts[i] == s) idx = i;
    if (idx < 0) return fail(SELDQ_ERR_INVALID, "bf16 mirror set of x lacks shift %d needed by tap %d", s, t);
    p.tap_sidx[t] = idx;
    p.off_w[t] += s;
  }
  if (gy.nshifts < 1 || gy.shifts[0] != 0) return fail(SELDQ_ERR_INVALID, "bf16 mirror set of gy must start with shift 0");
  alignas(64) CUtensorMap tm_g, tm_x;
  int rc = encode_mirror_map(&tm_g, gy, g.OW, g.OH, g.P, g.N, p.OS);
  if (rc) return rc;
  rc = encode_mirror_map(&tm_x, x, g.IW, g.IH, g.R, g.N, p.IS);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(qconv_umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "wgrad smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  dim3 grid(tiles, p.splits);
  qconv_umma_wgrad_kernel<<<grid, kThreads, smem, st>>>(tm_g, tm_x, p, pairs);
  return check_launch("qconv_umma_wgrad_kernel");
}

}  // namespace seldq

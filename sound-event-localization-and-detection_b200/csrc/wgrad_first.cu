// First CNN block, backward: BatchNorm-backward "apply" fused into the tcgen05 weight-gradient kernel.
//
// The first convolution (model.py:270-283, Cin = 8 | 16) needs no input gradient, so d(conv out)
//     d = a (g - mean(dy) - xhat mean(dy xhat)),  g = pooled gradient routed to the arg-max row (epilogue.cu)
// has exactly one consumer: this layer's weight gradient.  Writing it to HBM and reading it back costs 2 x 472 MB
// per sample (the largest tensor of the model) plus a kernel; here the gradient tile is produced in shared memory
// instead: TMA drops the conv-output tile (fp16, K-major, 128B swizzle) where the B operand is expected, and
// "transform" warps turn it into d in place, from the arg-max flags and the pooled gradient they hold in registers.
//
//   D[(tap, c), (a, o)] = sum_{n, h, w} x[n, c, h + off_h(tap), w + off_w(tap)] * d[n, a*Oc + o, h, w]
//
// M = taps x input channels (72 of 128 rows for the 8-channel model; the narrow input is why the roles are swapped
// with respect to wgrad_umma.cu: one MMA of N = 192 per 16 positions instead of 18 MMAs of N = 8), N = all output
// channels, K = positions (time contiguous in both tensors -> both operands K-major).  Each CTA reduces a
// contiguous slice of the (n, h, w-chunk) axis; the epilogue folds D onto the compact gradients
// (sign-weighted sum over the (a, b) pairs of each compact tensor) and adds them with atomicAdd.
//
// Warp roles (320 threads): warp 0 = TMA producer of the x taps (bf16 mirror set, conv_umma.h), warp 1 = TMEM
// owner + MMA issuer, warps 2..9 = transform producers of d, then epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "epilogue.h"
#include "launch.h"
#include "pdl.cuh"
#include "tensor_map.h"
#include "umma_ptx.cuh"
#include "wgrad_first.h"

namespace seldq {
namespace first {

constexpr int kProducers = 256;
constexpr int kThreads = 64 + kProducers;
constexpr int kStages = 4;

__global__ void __launch_bounds__(kThreads, 1)
first_layer_bwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                       const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  __shared__ __align__(8) uint64_t full_a[kStages], full_y[kStages], full_b[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  const long long per = (p.ksteps + gridDim.x - 1) / gridDim.x;
  const long long k_begin = (long long)blockIdx.x * per;
  const long long k_end = min(p.ksteps, k_begin + per);
  const int nk = (int)max(0LL, k_end - k_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_a[i], 1);
      ptx::mbar_init(&full_y[i], 1);
      ptx::mbar_init(&full_b[i], kProducers);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_x);
    ptx::prefetch_tensormap(&tm_y);
  }
  // rows of the A tile that no tap box covers stay zero for the whole kernel
  for (int s = 0; s < kStages; ++s) {
    uint4* a = reinterpret_cast<uint4*>(smem + (size_t)s * p.stage_bytes);
    for (int i = p.rows * 8 + threadIdx.x; i < 128 * 8; i += kThreads) a[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async();
  if (warp == 1) ptx::tmem_alloc(&tmem_slot, 256u);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();

  if (nk > 0) {
    if (warp == 0) {
      // ===== x taps by TMA ===========================================================================
      if (ptx::elect_one()) {
        uint32_t slot = 0, parity = 0;
        for (int ks = 0; ks < nk; ++ks) {
          // K-step order: the `pool` rows under one pooled row are consecutive (the transform warps keep that pooled
          // row's flags and gradient in registers)
          long long u = k_begin + ks;
          const int kk = (int)(u % p.pool); u /= p.pool;
          const int wc = (int)(u % p.chunks_w); u /= p.chunks_w;
          const int h = (int)(u % p.HP) * p.pool + kk;
          const int n = (int)(u / p.HP);
          const int w0 = wc * 64;
          ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
          ptx::mbar_arrive_expect_tx(&full_a[slot], (uint32_t)p.ntaps * (uint32_t)p.R * 128u);
          uint8_t* st = smem + (size_t)slot * p.stage_bytes;
          // off_w already contains the tap's mirror shift: the inner coordinate is a multiple of 8
          for (int t = 0; t < p.ntaps; ++t)
            ptx::tma_load_5d(st + (size_t)t * p.R * 128, &tm_x, &full_a[slot], w0 + p.off_w[t], h + p.off_h[t], 0, n,
                             p.tap_sidx[t]);
          // the conv output tile [NB channels x 64 t] lands where the transform warps turn it into d in place
          ptx::mbar_arrive_expect_tx(&full_y[slot], p.b_bytes);
          ptx::tma_load_4d(st + p.a_bytes, &tm_y, &full_y[slot], w0, h, 0, n);
          if (++slot == kStages) { slot = 0; parity ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer ==============================================================================
      if (ptx::elect_one()) {
        // both operands K-major, 128B swizzle: 8-row groups 1024 B apart (SBO), K advances inside the swizzled row
        const uint64_t hi = ptx::make_smem_desc_hi(16, 1024, ptx::kSwizzle128B);
        const uint32_t idesc = ptx::make_idesc_bf16(128, (uint32_t)p.NB, 0, 0, 0, 0);
        const uint32_t base = ptx::smem_u32(smem);
        uint32_t slot = 0, parity = 0;
        for (int ks = 0; ks < nk; ++ks) {
          ptx::mbar_wait(&full_a[slot], parity);
          ptx::mbar_wait(&full_b[slot], parity);
          ptx::tc_fence_after();
          const uint32_t st = base + slot * p.stage_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_base, ptx::smem_desc(hi, st + k * 32u), ptx::smem_desc(hi, st + p.a_bytes + k * 32u), idesc,
                          (ks > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(&empty_bar[slot]);
          if (++slot == kStages) { slot = 0; parity ^= 1; }
        }
        ptx::umma_commit(&done_bar);
      }
    } else {
      // ===== transform producers: d(conv out) tile [NB channels x 64 t], K-major / 128B swizzle ======
      const int pt = threadIdx.x - 64;                     // 0..255
      constexpr int VPT = 8;                               // 16-byte vectors per thread and K-step (NB <= 256)
      const int vec_per_step = p.NB * 8;                   // a vector = 8 consecutive t of one channel
      uint32_t slot = 0, parity = 0;
      // per vector (channel c, 8 consecutive w), kept in registers while the `pool` rows of one pooled row go by:
      //   d = a (g - m1 - xhat m2) = ga - B y + Cc   with ga = a g (0 off the arg-max row), B = a rstd m2,
      //   Cc = a (mean rstd m2 - m1)
      uint2 idv[VPT];
      float ga[VPT][8];
      float Bc[VPT], Cc[VPT], Ac[VPT];
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (pt + kProducers * i) >> 3;
        const bool on = pt + kProducers * i < vec_per_step && c < p.C;
        const float4 cf = on ? __ldg(reinterpret_cast<const float4*>(p.coef) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 dm = on ? __ldg(p.dmean + c) : make_float2(0.f, 0.f);
        Ac[i] = cf.x * p.scale;
        Bc[i] = cf.x * cf.w * dm.y;
        Cc[i] = cf.x * (cf.z * cf.w * dm.y - dm.x);
      }
      for (int ks = 0; ks < nk; ++ks) {
        long long u = k_begin + ks;
        const int kk = (int)(u % p.pool); u /= p.pool;
        const int wc = (int)(u % p.chunks_w); u /= p.chunks_w;
        const int hp = (int)(u % p.HP);
        const int n = (int)(u / p.HP);
        const uint32_t want4 = (0x80u | (uint32_t)kk) * 0x01010101u;
        if (ks == 0 || kk == 0) {                          // a new pooled row: its flags and gradient
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            const int v = pt + kProducers * i;
            const int c = v >> 3, w = wc * 64 + (v & 7) * 8;
            const bool on = v < vec_per_step && c < p.C && w < p.W;
            const long long e = (((long long)n * p.C + c) * p.HP + hp) * p.W + w;
            idv[i] = on ? __ldg(reinterpret_cast<const uint2*>(p.idx + e)) : make_uint2(0u, 0u);
            const float4 g0 = on ? __ldg(reinterpret_cast<const float4*>(p.gz + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 g1 = on ? __ldg(reinterpret_cast<const float4*>(p.gz + e + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            ga[i][0] = g0.x * Ac[i]; ga[i][1] = g0.y * Ac[i]; ga[i][2] = g0.z * Ac[i]; ga[i][3] = g0.w * Ac[i];
            ga[i][4] = g1.x * Ac[i]; ga[i][5] = g1.y * Ac[i]; ga[i][6] = g1.z * Ac[i]; ga[i][7] = g1.w * Ac[i];
          }
        }
        ptx::mbar_wait(&full_y[slot], parity);             // y tile landed (TMA, 128B swizzle)
        uint8_t* bt = smem + (size_t)slot * p.stage_bytes + p.a_bytes;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int v = pt + kProducers * i;
          if (v >= vec_per_step) break;
          const int c = v >> 3, j = v & 7;
          const int w = wc * 64 + j * 8;
          uint4* cell = reinterpret_cast<uint4*>(bt + c * 128 + ((j ^ (c & 7)) << 4));
          uint32_t o[4] = {0u, 0u, 0u, 0u};
          if (c < p.C && w < p.W) {
            const uint4 yv = *cell;
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
            // byte-wise: 0xff where (flag & 0x87) == (0x80 | kk), i.e. kept and arg-max row == this row
            const uint32_t hit[2] = {__vcmpeq4(idv[i].x & 0x87878787u, want4), __vcmpeq4(idv[i].y & 0x87878787u, want4)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 yf = __half22float2(*reinterpret_cast<const __half2*>(&yw[q]));
              const float y0 = yf.x, y1 = yf.y;
              const float g0 = (hit[q >> 1] & (1u << (16 * (q & 1)))) ? ga[i][2 * q] : 0.f;
              const float g1 = (hit[q >> 1] & (1u << (16 * (q & 1) + 8))) ? ga[i][2 * q + 1] : 0.f;
              const __nv_bfloat162 b2 = __floats2bfloat162_rn(fmaf(-Bc[i], y0, Cc[i] + g0), fmaf(-Bc[i], y1, Cc[i] + g1));
              o[q] = *reinterpret_cast<const uint32_t*>(&b2);
            }
          }
          *cell = make_uint4(o[0], o[1], o[2], o[3]);
        }
        ptx::fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core
        ptx::mbar_arrive(&full_b[slot]);
        if (++slot == kStages) { slot = 0; parity ^= 1; }
      }

      // ===== epilogue: TMEM -> shared staging -> fold onto the compact gradients ======================
      ptx::mbar_wait(&done_bar, 0);
      ptx::tc_fence_after();
      float* stg = reinterpret_cast<float*>(smem);         // all MMAs have retired: the operand ring is free
      const int pitch = p.NB + 1;
      if (warp < 6) {                                      // warps 2..5 cover the four TMEM lane quarters
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.NB; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(t_row + (uint32_t)c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) stg[row * pitch + c0 + j] = __uint_as_float(v[j]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const ConvGeom& g = p.g;
      const int per_e = g.Oc * g.Ic * p.ntaps;
      for (int tg = pt; tg < g.tab.nw * per_e; tg += kProducers) {
        const int e = tg / per_e;
        int r = tg - e * per_e;
        const int tap = r % p.ntaps; r /= p.ntaps;
        const int i = r % g.Ic;
        const int oo = r / g.Ic;
        float acc = 0.f;
        for (int k = 0; k < p.pair_n[e]; ++k) {
          const float val = stg[(tap * p.R + p.pair_b[e][k] * g.Ic + i) * pitch + p.pair_a[e][k] * g.Oc + oo];
          acc += p.pair_neg[e][k] ? -val : val;
        }
        atomicAdd(p.gw[e] + (long long)oo * g.wsO + (long long)i * g.wsI + (long long)tap * g.wsT, acc);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256u);
}

}  // namespace first

bool first_layer_bwd_supported(const ConvGeom& g) {
  const int ntaps = g.KH * g.KW;
  return g.sh == 1 && g.sw == 1 && ntaps <= umma::kMaxTaps && g.R % 8 == 0 && g.R * ntaps <= 128 && g.P <= 256 &&
         g.OW % 8 == 0 && g.tab.nc > 1;   // (+ pooled height divides the height: checked against the tail)
}

int launch_first_layer_bwd(const ConvGeom& g, const MirrorSet& x, const epi::TailParams& t, const float2* dmean,
                           float* const* host_gw, cudaStream_t st) {
  using namespace first;
  if (!first_layer_bwd_supported(g)) return fail(SELDQ_ERR_UNSUPPORTED, "fused first-layer backward: unsupported geometry");
  Params p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  for (int i = 0; i < g.tab.nw; ++i) p.gw[i] = host_gw[i];
  p.y = t.y; p.coef = t.coef; p.idx = t.idx; p.gz = t.gz; p.dmean = dmean;
  p.pool = t.pool;
  p.scale = t.drop_p > 0.f ? 1.f / (1.f - t.drop_p) : 1.f;
  p.N = g.N; p.C = g.P; p.H = g.OH; p.W = g.OW; p.HP = g.OH / t.pool;
  p.R = g.R;
  p.ntaps = g.KH * g.KW;
  for (int k = 0; k < p.ntaps; ++k) {
    p.off_h[k] = (k / g.KW) * g.dh - g.ph;
    p.off_w[k] = (k % g.KW) * g.dw - g.pw;
    const int s = ((-p.off_w[k]) % 8 + 8) % 8;
    int idx = -1;
    for (int i = 0; i < x.nshifts; ++i)
      if (x.shifts[i] == s) idx = i;
    if (idx < 0) return fail(SELDQ_ERR_INVALID, "bf16 mirror set of x lacks shift %d needed by tap %d", s, k);
    p.tap_sidx[k] = idx;
    p.off_w[k] += s;
  }
  p.rows = p.ntaps * p.R;
  p.NB = (g.P + 15) / 16 * 16;
  p.chunks_w = (g.OW + 63) / 64;
  p.ksteps = (long long)g.N * p.HP * p.pool * p.chunks_w;
  p.a_bytes = 128u * 128u;
  p.b_bytes = (uint32_t)p.NB * 128u;
  p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
  for (int a = 0; a < g.tab.nc; ++a)
    for (int b = 0; b < g.tab.nc; ++b) {
      const int e = g.tab.widx[a][b];
      if (e < 0) continue;
      const int k = p.pair_n[e]++;
      p.pair_a[e][k] = (int8_t)a; p.pair_b[e][k] = (int8_t)b; p.pair_neg[e][k] = (int8_t)(g.tab.sign[a][b] < 0);
    }
  size_t smem = (size_t)kStages * p.stage_bytes;
  const size_t stg = (size_t)128 * (p.NB + 1) * 4;
  if (smem < stg) smem = stg;
  smem += 4096;   // slack: the tensor core may fetch past the logical end of the last operand tile
  alignas(64) CUtensorMap tm_x, tm_y;
  int rc = encode_mirror_map(&tm_x, x, g.IW, g.IH, g.R, g.N, p.R);
  if (rc) return rc;
  {
    // y: (W, H, C, N) view of the bf16 conv output, box {64 t, 1, NB channels, 1}, 128B swizzle
    const uint64_t dims[4] = {(uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.P, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)g.OW * 2, (uint64_t)g.OW * g.OH * 2, (uint64_t)g.OW * g.OH * g.P * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)p.NB, 1};
    if ((rc = encode_tensor_map(&tm_y, t.y, 2, 4, dims, strides, box, 3))) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(first_layer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "first-layer backward smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  long long grid = umma::num_sms();
  if (grid > p.ksteps) grid = p.ksteps;
  const cudaError_t le = launch_pdl(first_layer_bwd_kernel, dim3((unsigned)grid), dim3(kThreads), smem, st, tm_x, tm_y, p);
  if (le != cudaSuccess) return fail(SELDQ_ERR_CUDA, "first_layer_bwd_kernel: %s", cudaGetErrorString(le));
  return check_launch("first_layer_bwd_kernel");
}

}  // namespace seldq

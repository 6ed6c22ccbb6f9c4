// Bandwidth-bound kernels for the glue of the TCN residual blocks (model.py:109-132):
//
//   x   = tanh(BN1(r))                                  tcn_preact_fwd      -> x fp32 + x as the bf16 operand of conv1
//   y_f = conv1_filter(x), y_g = conv1_gate(x)          (tcgen05 convolutions, fp32 NCW output)
//   y   = dropout1d(tanh(BN_f(y_f)) * sigmoid(BN_g(y_g)))   tcn_gate_fwd    -> y only as the bf16 operand of conv2
//   r'  = x + conv2_residual(y), skips += conv2_skip(y) tcn_residual_fwd    (+ the BN1 statistics of r' for the next block)
//
// and their backward counterparts: two per-channel reductions (sum dz, sum dz * xhat: the BatchNorm backward
// terms, which are also d beta / d gamma) followed by an apply pass that writes the gradient straight into the
// bf16 operand layouts the gradient convolutions read (channels-last for dgrad, pitched NCW for wgrad).
//
// All tensors are fp32 [N][C][T] unless noted.  BatchNorm is train-mode: batch statistics arrive as per-channel
// (sum, sum of squares) in double; every kernel derives mean / rstd / affine coefficients from them itself, so
// there is no finalize launch.  Two kernel shapes:
//   row kernels   grid (C, N, splits): one block owns a slice of one (n, c) row; float4 loads, block reduction,
//                 one double atomicAdd per quantity and block
//   tile kernels  one block = 64 padded channels x 128 t; a thread owns 8 consecutive t of one channel and the
//                 results leave as channels-last rows through the shared tile of tile_cl.cuh
#include <cstdio>
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "launch.h"
#include "pdl.cuh"
#include "tcn_glue.h"
#include "tile_cl.cuh"

namespace seldq {
namespace tcn {

// 64 channels x 64 t per block: at batch 1 (T = 4800) that is 450 blocks, all resident at once
constexpr int kTileW = 64;
using epi::vec_tile_off;

// {a, b, mean, rstd}: BN(v) = a v + b
// (everything in double, where the mean / variance difference cancels safely, but without double division or square
// root: those cost hundreds of cycles at the GPU's fp64 rate and every thread of the row kernels evaluates this)
__device__ __forceinline__ float4 bn_coef(const BnRef& bn, int c, double inv_count, float eps) {
  const double mean = bn.sums[2 * c] * inv_count;
  double var = bn.sums[2 * c + 1] * inv_count - mean * mean;
  if (var < 0) var = 0;
  // 1 / sqrt(var + eps) to double accuracy from the fp32 estimate and two Newton steps (multiplies only)
  const double v = var + (double)eps;
  double r = (double)rsqrtf((float)v);
  r = r * (1.5 - 0.5 * v * r * r);
  r = r * (1.5 - 0.5 * v * r * r);
  const float rstd = (float)r;
  const float g = bn.gamma ? __ldg(bn.gamma + c) : 1.f, bt = bn.beta ? __ldg(bn.beta + c) : 0.f;
  const float a = g * rstd;
  return make_float4(a, bt - (float)mean * a, (float)mean, rstd);
}

// nn.BatchNorm's running-statistics update (momentum, unbiased variance); one thread per channel
__device__ __forceinline__ void bn_update_running(const BnRef& bn, int c, double count, float momentum) {
  if (!bn.running_mean && !bn.running_var) return;
  const double mean = bn.sums[2 * c] / count;
  double var = bn.sums[2 * c + 1] / count - mean * mean;
  if (var < 0) var = 0;
  if (bn.running_mean) bn.running_mean[c] = (1.f - momentum) * bn.running_mean[c] + momentum * (float)mean;
  if (bn.running_var) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    bn.running_var[c] = (1.f - momentum) * bn.running_var[c] + momentum * (float)unbiased;
  }
}

// tanh / sigmoid from the special-function unit's exp2 and reciprocal: tanh v = 1 - 2 / (e^{2v} + 1),
// sigmoid v = 1 / (1 + e^{-v}), five and four instructions instead of tanhf's ~20 and a full-precision division;
// absolute error ~2e-7 (both saturate correctly: e^{2v} = inf gives 1, 0 gives -1).  tanh.approx.f32 (one MUFU,
// relative error 2^-11) measured another 0.01 ms faster per step but moved a borderline gradient of the reduced
// 16-channel model across its parity gate in one run of five, so the near-exact form stays.
// -DSELDQ_GLUE_EXACT_TANH restores tanhf / 1 / (1 + exp(-v)).
#if defined(SELDQ_GLUE_EXACT_TANH)
__device__ __forceinline__ float tanh_(float v) { return tanhf(v); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }
#else
__device__ __forceinline__ float tanh_(float v) { return 1.f - __fdividef(2.f, __expf(2.f * v) + 1.f); }
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
#endif

// inverted channel dropout (nn.Dropout1d): one decision per (n, c)
__device__ __forceinline__ float drop_scale(const GlueParams& p, unsigned long long seed, int n, int c) {
  if (p.drop_p <= 0.f) return 1.f;
  const uint32_t thresh = (uint32_t)fminf(p.drop_p * 4294967296.f, 4294967295.f);
  return epi::dropout_keep(seed, p.salt, (long long)n * p.C + c, thresh) ? 1.f / (1.f - p.drop_p) : 0.f;
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    o[j] = *reinterpret_cast<const uint32_t*>(&b2);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ void tile_put8(uint8_t* tile, int wl, int cl, const uint4& q) {
  const uint32_t o[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint16_t*>(tile + vec_tile_off(wl + j, cl * 2)) = (uint16_t)(o[j >> 1] >> (16 * (j & 1)));
}

// tile decode shared by the tile kernels
struct TileIdx {
  int n, ct, wt;
};
__device__ __forceinline__ TileIdx tile_decode(const GlueParams& p, long long blk) {
  TileIdx t;
  t.wt = (int)(blk % p.tiles_t); blk /= p.tiles_t;
  t.ct = (int)(blk % p.tiles_c);
  t.n = (int)(blk / p.tiles_c);
  return t;
}
// item k of this thread: local channel, local t, real channel (or -1 for a pad channel / out of range)
__device__ __forceinline__ void item_decode(const GlueParams& p, const TileIdx& ti, int k, int* cl, int* wl, int* c,
                                            int* t) {
  const int item = threadIdx.x + 256 * k;
  constexpr int kVecPerRow = kTileW / 8;
  *cl = item / kVecPerRow;
  *wl = (item % kVecPerRow) * 8;
  *t = ti.wt * kTileW + *wl;
  const int cp = ti.ct * 64 + *cl;
  const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
  *c = (cp < p.Cp && ci < p.cc && *t < p.T) ? comp * p.cc + ci : -1;
}
// per-tile BN coefficients of the tile's 64 channels -> shared memory
__device__ __forceinline__ void tile_coefs(const GlueParams& p, const TileIdx& ti, int which, float4* s_coef) {
  if (threadIdx.x < 64) {
    const int cp = ti.ct * 64 + threadIdx.x;
    const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
    if (cp < p.Cp && ci < p.cc) s_coef[threadIdx.x] = bn_coef(p.bn[which], comp * p.cc + ci, p.inv_count, p.eps);
  }
}

// ---- forward -------------------------------------------------------------------------------------------------
// x = tanh(BN1(r)):  in[0] = r -> out32 = x, out_cl[0] = x operand
__global__ void __launch_bounds__(256) preact_fwd_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[64];
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < p.C; c += 256) bn_update_running(p.bn[0], c, p.count, p.momentum);
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    const TileIdx ti = tile_decode(p, blk);
    tile_coefs(p, ti, 0, s_coef);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileW / 32; ++k) {
      int cl, wl, c, t;
      item_decode(p, ti, k, &cl, &wl, &c, &t);
      float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c >= 0) {
        const long long e = ((long long)ti.n * p.C + c) * p.T + t;
        const float4 cf = s_coef[cl];
        float r[8];
        load8(p.in[0] + e, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = tanh_(cf.x * r[j] + cf.y);
        store8(p.out32 + e, x);
      }
      tile_put8(tile, wl, cl, pack8(x));
    }
    __syncthreads();
    epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
    __syncthreads();
  }
}

// y = dropout1d(tanh(BN_f(y_f)) * sigmoid(BN_g(y_g))):  in[0] = y_f, in[1] = y_g -> out_cl[0] = y operand
__global__ void __launch_bounds__(256) gate_fwd_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[2][64];
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < p.C; c += 256) {
      bn_update_running(p.bn[0], c, p.count, p.momentum);
      bn_update_running(p.bn[1], c, p.count, p.momentum);
    }
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    const TileIdx ti = tile_decode(p, blk);
    tile_coefs(p, ti, 0, s_coef[0]);
    tile_coefs(p, ti, 1, s_coef[1]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileW / 32; ++k) {
      int cl, wl, c, t;
      item_decode(p, ti, k, &cl, &wl, &c, &t);
      float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c >= 0) {
        const long long e = ((long long)ti.n * p.C + c) * p.T + t;
        const float m = drop_scale(p, seed, ti.n, c);
        if (m != 0.f) {
          const float4 cf = s_coef[0][cl], cg = s_coef[1][cl];
          float f[8], g[8];
          load8(p.in[0] + e, f);
          load8(p.in[1] + e, g);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = m * tanh_(cf.x * f[j] + cf.y) * sigmoidf_(cg.x * g[j] + cg.y);
        }
      }
      tile_put8(tile, wl, cl, pack8(y));
    }
    __syncthreads();
    epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
    __syncthreads();
  }
}

// row kernel: per-channel (sum, sum of squares) of up to two tensors: in[z] -> stats_out[z] (zeroed by the
// caller); blockIdx.z = z * splits + split
__global__ void __launch_bounds__(256) row_stats_kernel(const __grid_constant__ GlueParams p, int splits, int chunk) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y;
  const int z = blockIdx.z / splits, sp = blockIdx.z - z * splits;
  const int lo = sp * chunk, hi = min(p.T, lo + chunk);
  const float* row = p.in[z] + ((long long)n * p.C + c) * p.T;
  float s1 = 0.f, s2 = 0.f;
  for (int t = lo + threadIdx.x * 4; t < hi; t += 256 * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + t));
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  epi::block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    atomicAdd(p.stats_out[z] + 2 * c, (double)s1);
    atomicAdd(p.stats_out[z] + 2 * c + 1, (double)s2);
  }
}

// row kernel: r' = in[0] (x) + in[1] (residual conv) -> out32, BN statistics of r' -> dsums[c][2];
// skips: accum = (accum_init ? 0 : accum) + in[2].  in[1] == null (last block): only the skip accumulation.
__global__ void __launch_bounds__(256) residual_fwd_kernel(const __grid_constant__ GlueParams p, int splits, int chunk,
                                                          int accum_init) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y, sp = blockIdx.z;
  const int lo = sp * chunk, hi = min(p.T, lo + chunk);
  const bool res = p.in[1] != nullptr && c < p.C;
  const bool skip = p.in[2] != nullptr && c < p.C2;
  const long long base = ((long long)n * p.C + c) * p.T, base2 = ((long long)n * p.C2 + c) * p.T;
  float s1 = 0.f, s2 = 0.f;
  for (int t = lo + threadIdx.x * 4; t < hi; t += 256 * 4) {
    if (res) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.in[0] + base + t));
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.in[1] + base + t));
      const float4 v = make_float4(x.x + r.x, x.y + r.y, x.z + r.z, x.w + r.w);
      *reinterpret_cast<float4*>(p.out32 + base + t) = v;
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (skip) {
      float4 a = __ldg(reinterpret_cast<const float4*>(p.in[2] + base2 + t));
      if (!accum_init) {
        const float4 o = *reinterpret_cast<const float4*>(p.accum + base2 + t);
        a = make_float4(a.x + o.x, a.y + o.y, a.z + o.z, a.w + o.w);
      }
      *reinterpret_cast<float4*>(p.accum + base2 + t) = a;
    }
  }
  if (p.in[1] != nullptr && p.dsums != nullptr) {
    epi::block_sum2(s1, s2);
    if (threadIdx.x == 0 && res) {
      atomicAdd(p.dsums + 2 * c, (double)s1);
      atomicAdd(p.dsums + 2 * c + 1, (double)s2);
    }
  }
}

// ---- backward ------------------------------------------------------------------------------------------------
// gate, per element: z_f = BN_f(y_f), z_g = BN_g(y_g), h_f = tanh z_f, h_g = sigmoid z_g, gy = m (gy1 + gy2)
//   dz_f = gy h_g (1 - h_f^2)      dz_g = gy h_f h_g (1 - h_g)
__device__ __forceinline__ void gate_dz(float f, float g, float gy, const float4& cf, const float4& cg, float* dzf,
                                        float* dzg, float* xhf, float* xhg) {
  const float hf = tanh_(cf.x * f + cf.y), hg = sigmoidf_(cg.x * g + cg.y);
  *dzf = gy * hg * (1.f - hf * hf);
  *dzg = gy * hf * hg * (1.f - hg);
  *xhf = (f - cf.z) * cf.w;
  *xhg = (g - cg.z) * cg.w;
}

// row kernel: dsums[0..3][c] = sum dz_f, sum dz_f xhat_f, sum dz_g, sum dz_g xhat_g (planar);  in = {y_f, y_g, gy1, gy2 | null}
__global__ void __launch_bounds__(256) gate_bwd_reduce_kernel(const __grid_constant__ GlueParams p, int splits,
                                                             int chunk) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y, sp = blockIdx.z;
  const int lo = sp * chunk, hi = min(p.T, lo + chunk);
  const long long base = ((long long)n * p.C + c) * p.T;
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const float m = drop_scale(p, seed, n, c);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (m != 0.f) {
    const float4 cf = bn_coef(p.bn[0], c, p.inv_count, p.eps), cg = bn_coef(p.bn[1], c, p.inv_count, p.eps);
    for (int t = lo + threadIdx.x * 4; t < hi; t += 256 * 4) {
      const float4 f4 = __ldg(reinterpret_cast<const float4*>(p.in[0] + base + t));
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.in[1] + base + t));
      float4 y4 = __ldg(reinterpret_cast<const float4*>(p.in[2] + base + t));
      if (p.in[3]) {
        const float4 y2 = __ldg(reinterpret_cast<const float4*>(p.in[3] + base + t));
        y4 = make_float4(y4.x + y2.x, y4.y + y2.y, y4.z + y2.z, y4.w + y2.w);
      }
      const float f[4] = {f4.x, f4.y, f4.z, f4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float dzf, dzg, xhf, xhg;
        gate_dz(f[j], g[j], m * y[j], cf, cg, &dzf, &dzg, &xhf, &xhg);
        s[0] += dzf; s[1] += dzf * xhf; s[2] += dzg; s[3] += dzg * xhg;
      }
    }
  }
  epi::block_sum2(s[0], s[1]);
  __syncthreads();
  epi::block_sum2(s[2], s[3]);
  if (threadIdx.x == 0 && m != 0.f) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(p.dsums + (long long)j * p.C + c, (double)s[j]);
  }
}

// tile kernel: d y_f, d y_g -> channels-last operands out_cl[0|1] and pitched NCW operands out_t16[0|1]
__global__ void __launch_bounds__(256) gate_bwd_apply_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[2][kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[2][64];
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const float inv_count = (float)p.inv_count;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    const TileIdx ti = tile_decode(p, blk);
    tile_coefs(p, ti, 0, s_coef[0]);
    tile_coefs(p, ti, 1, s_coef[1]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileW / 32; ++k) {
      int cl, wl, c, t;
      item_decode(p, ti, k, &cl, &wl, &c, &t);
      float df[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c >= 0) {
        const long long e = ((long long)ti.n * p.C + c) * p.T + t;
        const float m = drop_scale(p, seed, ti.n, c);
        const float4 cf = s_coef[0][cl], cg = s_coef[1][cl];
        const float m1f = (float)p.dsums[c] * inv_count, m2f = (float)p.dsums[p.C + c] * inv_count;
        const float m1g = (float)p.dsums[2 * p.C + c] * inv_count, m2g = (float)p.dsums[3 * p.C + c] * inv_count;
        float f[8], g[8], y[8];
        load8(p.in[0] + e, f);
        load8(p.in[1] + e, g);
        load8(p.in[2] + e, y);
        if (p.in[3]) {
          float y2[8];
          load8(p.in[3] + e, y2);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] += y2[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dzf, dzg, xhf, xhg;
          gate_dz(f[j], g[j], m * y[j], cf, cg, &dzf, &dzg, &xhf, &xhg);
          df[j] = cf.x * (dzf - m1f - xhf * m2f);
          dg[j] = cg.x * (dzg - m1g - xhg * m2g);
        }
        const long long et = ((long long)ti.n * p.C + c) * p.pitch + t;
        *reinterpret_cast<uint4*>(p.out_t16[0] + et) = pack8(df);
        *reinterpret_cast<uint4*>(p.out_t16[1] + et) = pack8(dg);
      }
      tile_put8(tile[0], wl, cl, pack8(df));
      tile_put8(tile[1], wl, cl, pack8(dg));
    }
    __syncthreads();
    epi::vec_tile_store_cl<kTileW>(tile[0], p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
    epi::vec_tile_store_cl<kTileW>(tile[1], p.out_cl[1], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
    __syncthreads();
  }
}

// pre-activation: x = tanh(BN1(r)), dz = (g_r' + gx1 + gx2) (1 - x^2);  in = {g_r' | null, gx1, gx2, x, r}
// row kernel: dsums[0..1][c] = sum dz, sum dz xhat (planar)
__global__ void __launch_bounds__(256) preact_bwd_reduce_kernel(const __grid_constant__ GlueParams p, int splits,
                                                               int chunk) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y, sp = blockIdx.z;
  const int lo = sp * chunk, hi = min(p.T, lo + chunk);
  const long long base = ((long long)n * p.C + c) * p.T;
  const float4 cf = bn_coef(p.bn[0], c, p.inv_count, p.eps);
  float s1 = 0.f, s2 = 0.f;
  for (int t = lo + threadIdx.x * 4; t < hi; t += 256 * 4) {
    float4 g = __ldg(reinterpret_cast<const float4*>(p.in[1] + base + t));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(p.in[2] + base + t));
    g = make_float4(g.x + g2.x, g.y + g2.y, g.z + g2.z, g.w + g2.w);
    if (p.in[0]) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.in[0] + base + t));
      g = make_float4(g.x + g0.x, g.y + g0.y, g.z + g0.z, g.w + g0.w);
    }
    const float4 x = __ldg(reinterpret_cast<const float4*>(p.in[3] + base + t));
    const float4 r = __ldg(reinterpret_cast<const float4*>(p.in[4] + base + t));
    const float gg[4] = {g.x, g.y, g.z, g.w}, xx[4] = {x.x, x.y, x.z, x.w}, rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dz = gg[j] * (1.f - xx[j] * xx[j]);
      s1 += dz;
      s2 += dz * (rr[j] - cf.z) * cf.w;
    }
  }
  epi::block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    atomicAdd(p.dsums + c, (double)s1);
    atomicAdd(p.dsums + p.C + c, (double)s2);
  }
}

// tile kernel: g_r = a (dz - mean dz - xhat mean(dz xhat)) -> out32 (fp32) and, when requested, the operands
// out_cl[0] / out_t16[0] of the previous block's conv2 gradient kernels
__global__ void __launch_bounds__(256) preact_bwd_apply_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[64];
  const float inv_count = (float)p.inv_count;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    const TileIdx ti = tile_decode(p, blk);
    tile_coefs(p, ti, 0, s_coef);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTileW / 32; ++k) {
      int cl, wl, c, t;
      item_decode(p, ti, k, &cl, &wl, &c, &t);
      float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c >= 0) {
        const long long e = ((long long)ti.n * p.C + c) * p.T + t;
        const float4 cf = s_coef[cl];
        const float m1 = (float)p.dsums[c] * inv_count, m2 = (float)p.dsums[p.C + c] * inv_count;
        float g[8], g2[8], x[8], r[8];
        load8(p.in[1] + e, g);
        load8(p.in[2] + e, g2);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += g2[j];
        if (p.in[0]) {
          load8(p.in[0] + e, g2);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += g2[j];
        }
        load8(p.in[3] + e, x);
        load8(p.in[4] + e, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = g[j] * (1.f - x[j] * x[j]);
          d[j] = cf.x * (dz - m1 - (r[j] - cf.z) * cf.w * m2);
        }
        store8(p.out32 + e, d);
        if (p.out_t16[0]) *reinterpret_cast<uint4*>(p.out_t16[0] + ((long long)ti.n * p.C + c) * p.pitch + t) = pack8(d);
      }
      if (p.out_cl[0]) tile_put8(tile, wl, cl, pack8(d));
    }
    if (p.out_cl[0]) {
      __syncthreads();
      epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
    }
    __syncthreads();
  }
}


// ---- single-launch steps (reduce + apply behind a grid barrier) ---------------------------------------------------
// At batch 1 a reduce / apply pair costs two launches of 8-12 us for tensors that sit in L2, and the apply pass
// re-reads everything the reduction read.  Here ONE block owns one 64-channel x 64-t tile for the whole step: it
// reduces its tile (eight lanes share a channel: three shuffles, then one double atomicAdd per quantity), keeps what
// the apply pass needs in registers, waits until every block of the grid has added its partial sums, and applies.
// All blocks must be resident at once: the launcher checks the grid against the kernel's occupancy and the host falls
// back to the two-launch form otherwise (tcn_glue_fused_supported).  Nothing a resident block waits for can depend on
// a block that is not resident yet: work of other streams finishes on its own, and a programmatic dependent of this
// kernel is launched only after every block has started.  The spin is bounded (trap, as mbar_wait).
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    unsigned int seen = 0, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (seen >= nblocks) break;
      if ((++spins & 255u) == 0 && clock64() - t0 > 4000000000LL) {
        printf("seldq: grid barrier timed out (block %d: %u of %u blocks arrived)\n", (int)blockIdx.x, seen, nblocks);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// sum over the eight consecutive lanes that share a channel (item layout of the tile kernels)
__device__ __forceinline__ float sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ double ldcg_f64(const double* p) {       // partial sums live in L2 (atomics): bypass L1
  return __ldcg(p);
}
// bn_coef from statistics other blocks have just accumulated
__device__ __forceinline__ float4 bn_coef_fresh(const BnRef& bn, const double* sums, int c, double inv_count, float eps) {
  const double mean = ldcg_f64(sums + 2 * c) * inv_count;
  double var = ldcg_f64(sums + 2 * c + 1) * inv_count - mean * mean;
  if (var < 0) var = 0;
  const double v = var + (double)eps;
  double r = (double)rsqrtf((float)v);
  r = r * (1.5 - 0.5 * v * r * r);
  r = r * (1.5 - 0.5 * v * r * r);
  const float rstd = (float)r;
  const float g = bn.gamma ? __ldg(bn.gamma + c) : 1.f, bt = bn.beta ? __ldg(bn.beta + c) : 0.f;
  const float a = g * rstd;
  return make_float4(a, bt - (float)mean * a, (float)mean, rstd);
}
__device__ __forceinline__ void bn_update_running_fresh(const BnRef& bn, const double* sums, int c, double count,
                                                        float momentum) {
  if (!bn.running_mean && !bn.running_var) return;
  const double mean = ldcg_f64(sums + 2 * c) / count;
  double var = ldcg_f64(sums + 2 * c + 1) / count - mean * mean;
  if (var < 0) var = 0;
  if (bn.running_mean) bn.running_mean[c] = (1.f - momentum) * bn.running_mean[c] + momentum * (float)mean;
  if (bn.running_var) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    bn.running_var[c] = (1.f - momentum) * bn.running_var[c] + momentum * (float)unbiased;
  }
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// GATE_BWD = GATE_BWD_REDUCE + GATE_BWD_APPLY.  Kept across the barrier: dz_f, dz_g of the thread's 2 x 8 elements;
// y_f / y_g are read again for xhat (16 of the 44 MB the two-launch form moves twice).
__global__ void __launch_bounds__(256, 4) gate_bwd_fused_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[2][kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[2][64];
  __shared__ float4 s_mom[64];
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const float inv_count = (float)p.inv_count;
  const TileIdx ti = tile_decode(p, blockIdx.x);
  tile_coefs(p, ti, 0, s_coef[0]);
  tile_coefs(p, ti, 1, s_coef[1]);
  __syncthreads();
  float dzf[kTileW / 32][8], dzg[kTileW / 32][8];
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) { dzf[k][j] = 0.f; dzg[k][j] = 0.f; }
    if (c >= 0) {
      const float m = drop_scale(p, seed, ti.n, c);
      if (m != 0.f) {
        const long long e = ((long long)ti.n * p.C + c) * p.T + t;
        const float4 cf = s_coef[0][cl], cg = s_coef[1][cl];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float f[4], g[4], y[4];
          ld4(p.in[0] + e + 4 * h, f);
          ld4(p.in[1] + e + 4 * h, g);
          ld4(p.in[2] + e + 4 * h, y);
          if (p.in[3]) {
            float y2[4];
            ld4(p.in[3] + e + 4 * h, y2);
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] += y2[j];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float xhf, xhg;
            gate_dz(f[j], g[j], m * y[j], cf, cg, &dzf[k][4 * h + j], &dzg[k][4 * h + j], &xhf, &xhg);
            s[0] += dzf[k][4 * h + j]; s[1] += dzf[k][4 * h + j] * xhf;
            s[2] += dzg[k][4 * h + j]; s[3] += dzg[k][4 * h + j] * xhg;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = sum8(s[i]);
    if ((threadIdx.x & 7) == 0 && c >= 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (s[i] != 0.f) atomicAdd(p.dsums + (long long)i * p.C + c, (double)s[i]);
    }
  }
  grid_barrier(p.sync, gridDim.x);
  if (threadIdx.x < 64) {
    const int cp = ti.ct * 64 + threadIdx.x;
    const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
    if (cp < p.Cp && ci < p.cc) {
      const int c = comp * p.cc + ci;
      s_mom[threadIdx.x] = make_float4((float)ldcg_f64(p.dsums + c) * inv_count, (float)ldcg_f64(p.dsums + p.C + c) * inv_count,
                                       (float)ldcg_f64(p.dsums + 2 * p.C + c) * inv_count,
                                       (float)ldcg_f64(p.dsums + 3 * p.C + c) * inv_count);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float df[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dg[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      const float4 cf = s_coef[0][cl], cg = s_coef[1][cl], mo = s_mom[cl];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[4], g[4];
        ld4(p.in[0] + e + 4 * h, f);
        ld4(p.in[1] + e + 4 * h, g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          df[4 * h + j] = cf.x * (dzf[k][4 * h + j] - mo.x - (f[j] - cf.z) * cf.w * mo.y);
          dg[4 * h + j] = cg.x * (dzg[k][4 * h + j] - mo.z - (g[j] - cg.z) * cg.w * mo.w);
        }
      }
      const long long et = ((long long)ti.n * p.C + c) * p.pitch + t;
      *reinterpret_cast<uint4*>(p.out_t16[0] + et) = pack8(df);
      *reinterpret_cast<uint4*>(p.out_t16[1] + et) = pack8(dg);
    }
    tile_put8(tile[0], wl, cl, pack8(df));
    tile_put8(tile[1], wl, cl, pack8(dg));
  }
  __syncthreads();
  epi::vec_tile_store_cl<kTileW>(tile[0], p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
  epi::vec_tile_store_cl<kTileW>(tile[1], p.out_cl[1], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
}

// PREACT_BWD = PREACT_BWD_REDUCE + PREACT_BWD_APPLY.  Kept across the barrier: dz and xhat (nothing is read twice).
__global__ void __launch_bounds__(256, 4) preact_bwd_fused_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[64];
  __shared__ float2 s_mom[64];
  const float inv_count = (float)p.inv_count;
  const TileIdx ti = tile_decode(p, blockIdx.x);
  tile_coefs(p, ti, 0, s_coef);
  __syncthreads();
  float dz[kTileW / 32][8], xh[kTileW / 32][8];
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dz[k][j] = 0.f; xh[k][j] = 0.f; }
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      const float4 cf = s_coef[cl];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float g[4], g2[4], x[4], r[4];
        ld4(p.in[1] + e + 4 * h, g);
        ld4(p.in[2] + e + 4 * h, g2);
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] += g2[j];
        if (p.in[0]) {
          ld4(p.in[0] + e + 4 * h, g2);
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] += g2[j];
        }
        ld4(p.in[3] + e + 4 * h, x);
        ld4(p.in[4] + e + 4 * h, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d = g[j] * (1.f - x[j] * x[j]), xx = (r[j] - cf.z) * cf.w;
          dz[k][4 * h + j] = d;
          xh[k][4 * h + j] = xx;
          s1 += d;
          s2 += d * xx;
        }
      }
    }
    s1 = sum8(s1);
    s2 = sum8(s2);
    if ((threadIdx.x & 7) == 0 && c >= 0) {
      atomicAdd(p.dsums + c, (double)s1);
      atomicAdd(p.dsums + p.C + c, (double)s2);
    }
  }
  grid_barrier(p.sync, gridDim.x);
  if (threadIdx.x < 64) {
    const int cp = ti.ct * 64 + threadIdx.x;
    const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
    if (cp < p.Cp && ci < p.cc) {
      const int c = comp * p.cc + ci;
      s_mom[threadIdx.x] = make_float2((float)ldcg_f64(p.dsums + c) * inv_count, (float)ldcg_f64(p.dsums + p.C + c) * inv_count);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      const float a = s_coef[cl].x;
      const float2 mo = s_mom[cl];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = a * (dz[k][j] - mo.x - xh[k][j] * mo.y);
      store8(p.out32 + e, d);
      if (p.out_t16[0]) *reinterpret_cast<uint4*>(p.out_t16[0] + ((long long)ti.n * p.C + c) * p.pitch + t) = pack8(d);
    }
    if (p.out_cl[0]) tile_put8(tile, wl, cl, pack8(d));
  }
  if (p.out_cl[0]) {
    __syncthreads();
    epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
  }
}

// GATE_FWD_STATS = ROW_STATS(y_f, y_g) + GATE_FWD.  Kept across the barrier: y_f, y_g.
__global__ void __launch_bounds__(256, 4) gate_fwd_fused_kernel(const __grid_constant__ GlueParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[2][64];
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const TileIdx ti = tile_decode(p, blockIdx.x);
  float f[kTileW / 32][8], g[kTileW / 32][8];
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[k][j] = 0.f; g[k][j] = 0.f; }
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      load8(p.in[0] + e, f[k]);
      load8(p.in[1] + e, g[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[0] += f[k][j]; s[1] = fmaf(f[k][j], f[k][j], s[1]);
        s[2] += g[k][j]; s[3] = fmaf(g[k][j], g[k][j], s[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = sum8(s[i]);
    if ((threadIdx.x & 7) == 0 && c >= 0) {
      atomicAdd(p.stats_out[0] + 2 * c, (double)s[0]);
      atomicAdd(p.stats_out[0] + 2 * c + 1, (double)s[1]);
      atomicAdd(p.stats_out[1] + 2 * c, (double)s[2]);
      atomicAdd(p.stats_out[1] + 2 * c + 1, (double)s[3]);
    }
  }
  grid_barrier(p.sync, gridDim.x);
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < p.C; c += 256) {
      bn_update_running_fresh(p.bn[0], p.stats_out[0], c, p.count, p.momentum);
      bn_update_running_fresh(p.bn[1], p.stats_out[1], c, p.count, p.momentum);
    }
  if (threadIdx.x < 64) {
    const int cp = ti.ct * 64 + threadIdx.x;
    const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
    if (cp < p.Cp && ci < p.cc) {
      s_coef[0][threadIdx.x] = bn_coef_fresh(p.bn[0], p.stats_out[0], comp * p.cc + ci, p.inv_count, p.eps);
      s_coef[1][threadIdx.x] = bn_coef_fresh(p.bn[1], p.stats_out[1], comp * p.cc + ci, p.inv_count, p.eps);
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c >= 0) {
      const float m = drop_scale(p, seed, ti.n, c);
      if (m != 0.f) {
        const float4 cf = s_coef[0][cl], cg = s_coef[1][cl];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = m * tanh_(cf.x * f[k][j] + cf.y) * sigmoidf_(cg.x * g[k][j] + cg.y);
      }
    }
    tile_put8(tile, wl, cl, pack8(y));
  }
  __syncthreads();
  epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
}

// RESIDUAL_PREACT_FWD = RESIDUAL_FWD + the next block's PREACT_FWD (C2 == C):
//   r' = in[0] + in[1] -> out32, its statistics -> dsums; accum = (accum_init ? 0 : accum) + in[2];
//   behind the barrier x' = tanh(BN(r')) with bn[0] = the NEXT block's batch_filter1 -> out32b, out_cl[0].
// Kept across the barrier: r'.
__global__ void __launch_bounds__(256, 4) residual_preact_fused_kernel(const __grid_constant__ GlueParams p, int accum_init) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kTileW * epi::kVecPitch];
  __shared__ float4 s_coef[64];
  const TileIdx ti = tile_decode(p, blockIdx.x);
  float r[kTileW / 32][8];
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[k][j] = 0.f;
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      float x[8];
      load8(p.in[0] + e, x);
      load8(p.in[1] + e, r[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[k][j] += x[j];
        s1 += r[k][j];
        s2 = fmaf(r[k][j], r[k][j], s2);
      }
      store8(p.out32 + e, r[k]);
      load8(p.in[2] + e, x);
      if (!accum_init) {
        float o[8];
        load8(p.accum + e, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += o[j];
      }
      store8(p.accum + e, x);
    }
    s1 = sum8(s1);
    s2 = sum8(s2);
    if ((threadIdx.x & 7) == 0 && c >= 0) {
      atomicAdd(p.dsums + 2 * c, (double)s1);
      atomicAdd(p.dsums + 2 * c + 1, (double)s2);
    }
  }
  grid_barrier(p.sync, gridDim.x);
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < p.C; c += 256) bn_update_running_fresh(p.bn[0], p.dsums, c, p.count, p.momentum);
  if (threadIdx.x < 64) {
    const int cp = ti.ct * 64 + threadIdx.x;
    const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
    if (cp < p.Cp && ci < p.cc) s_coef[threadIdx.x] = bn_coef_fresh(p.bn[0], p.dsums, comp * p.cc + ci, p.inv_count, p.eps);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTileW / 32; ++k) {
    int cl, wl, c, t;
    item_decode(p, ti, k, &cl, &wl, &c, &t);
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c >= 0) {
      const long long e = ((long long)ti.n * p.C + c) * p.T + t;
      const float4 cf = s_coef[cl];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = tanh_(cf.x * r[k][j] + cf.y);
      store8(p.out32b + e, x);
    }
    tile_put8(tile, wl, cl, pack8(x));
  }
  __syncthreads();
  epi::vec_tile_store_cl<kTileW>(tile, p.out_cl[0], (long long)ti.n * p.T, ti.wt * kTileW, p.T, p.Cp, ti.ct);
}

}  // namespace tcn

// ---- launchers -------------------------------------------------------------------------------------------------
static int tile_grid(tcn::GlueParams& p) {
  p.tiles_t = (p.T + tcn::kTileW - 1) / tcn::kTileW;
  p.tiles_c = (p.Cp + 63) / 64;
  p.total_blocks = (long long)p.tiles_t * p.tiles_c * p.N;
  const long long cap = 148LL * 16;
  return (int)(p.total_blocks < cap ? (p.total_blocks < 1 ? 1 : p.total_blocks) : cap);
}
static void row_grid(const tcn::GlueParams& p, int ntensors, dim3* grid, int* splits, int* chunk) {
  // about 2 K elements per block
  int s = (p.T + 2047) / 2048;
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  int ch = (p.T + s - 1) / s;
  ch = (ch + 3) & ~3;
  s = (p.T + ch - 1) / ch;
  *splits = s;
  *chunk = ch;
  *grid = dim3((unsigned)p.C, (unsigned)p.N, (unsigned)(s * ntensors));
}

static int launch_tcn_glue_fused(int op, tcn::GlueParams& p, int flag, cudaStream_t st);

int launch_tcn_glue(int op, tcn::GlueParams& p, int flag, cudaStream_t st) {
  if (p.T % 8) return fail(SELDQ_ERR_UNSUPPORTED, "TCN glue kernels need a time extent that is a multiple of 8, got %d", p.T);
  if (p.N > 65535) return fail(SELDQ_ERR_UNSUPPORTED, "TCN glue: batch too large for the grid");
  dim3 grid;
  int splits = 1, chunk = 0;
  switch (op) {
    case tcn::OP_GATE_BWD:
    case tcn::OP_PREACT_BWD:
    case tcn::OP_GATE_FWD_STATS:
    case tcn::OP_RESIDUAL_PREACT_FWD:
      return launch_tcn_glue_fused(op, p, flag, st);
    case tcn::OP_PREACT_FWD:
      launch_pdl(tcn::preact_fwd_kernel, dim3(tile_grid(p)), dim3(256), 0, st, p);
      return check_launch("tcn::preact_fwd_kernel");
    case tcn::OP_GATE_FWD:
      launch_pdl(tcn::gate_fwd_kernel, dim3(tile_grid(p)), dim3(256), 0, st, p);
      return check_launch("tcn::gate_fwd_kernel");
    case tcn::OP_ROW_STATS:
      row_grid(p, flag, &grid, &splits, &chunk);
      launch_pdl(tcn::row_stats_kernel, grid, dim3(256), 0, st, p, splits, chunk);
      return check_launch("tcn::row_stats_kernel");
    case tcn::OP_RESIDUAL_FWD:
      row_grid(p, 1, &grid, &splits, &chunk);
      grid.x = (unsigned)(p.C2 > p.C ? p.C2 : p.C);
      launch_pdl(tcn::residual_fwd_kernel, grid, dim3(256), 0, st, p, splits, chunk, flag);
      return check_launch("tcn::residual_fwd_kernel");
    case tcn::OP_GATE_BWD_REDUCE:
      row_grid(p, 1, &grid, &splits, &chunk);
      launch_pdl(tcn::gate_bwd_reduce_kernel, grid, dim3(256), 0, st, p, splits, chunk);
      return check_launch("tcn::gate_bwd_reduce_kernel");
    case tcn::OP_GATE_BWD_APPLY:
      launch_pdl(tcn::gate_bwd_apply_kernel, dim3(tile_grid(p)), dim3(256), 0, st, p);
      return check_launch("tcn::gate_bwd_apply_kernel");
    case tcn::OP_PREACT_BWD_REDUCE:
      row_grid(p, 1, &grid, &splits, &chunk);
      launch_pdl(tcn::preact_bwd_reduce_kernel, grid, dim3(256), 0, st, p, splits, chunk);
      return check_launch("tcn::preact_bwd_reduce_kernel");
    case tcn::OP_PREACT_BWD_APPLY:
      launch_pdl(tcn::preact_bwd_apply_kernel, dim3(tile_grid(p)), dim3(256), 0, st, p);
      return check_launch("tcn::preact_bwd_apply_kernel");
  }
  return fail(SELDQ_ERR_INVALID, "unknown TCN glue op %d", op);
}

// ---- single-launch steps: every tile's block must be resident at once (grid barrier) --------------------------
static int fused_capacity() {
  static int cap = -1;
  if (cap < 0) {
    int c = 1 << 30;
    const void* kernels[4] = {(const void*)tcn::gate_bwd_fused_kernel, (const void*)tcn::preact_bwd_fused_kernel,
                              (const void*)tcn::gate_fwd_fused_kernel, (const void*)tcn::residual_preact_fused_kernel};
    for (int i = 0; i < 4; ++i) {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernels[i], 256, 0) != cudaSuccess) per_sm = 0;
      if (per_sm * cl::num_sms() < c) c = per_sm * cl::num_sms();
    }
    cap = c;
  }
  return cap;
}
int tcn_glue_fused_supported(int n, int c, int t) {
  if (n < 1 || c < 1 || t < 1 || (t % 8)) return 0;
  if (const char* e = getenv("SELDQ_TCN_FUSED_GLUE")) if (e[0] == '0') return 0;
  const long long blocks = (long long)((t + tcn::kTileW - 1) / tcn::kTileW) * ((c + 63) / 64) * n;
  return blocks <= fused_capacity() ? 1 : 0;
}
static int launch_tcn_glue_fused(int op, tcn::GlueParams& p, int flag, cudaStream_t st) {
  const int grid = tile_grid(p);
  if (p.total_blocks > fused_capacity())
    return fail(SELDQ_ERR_UNSUPPORTED, "single-launch TCN glue step: %lld tiles cannot be resident at once (capacity %d); "
                "use the reduce / apply pair", p.total_blocks, fused_capacity());
  if ((long long)grid != p.total_blocks) return fail(SELDQ_ERR_INVALID, "single-launch TCN glue step: grid mismatch");
  switch (op) {
    case tcn::OP_GATE_BWD:
      launch_pdl(tcn::gate_bwd_fused_kernel, dim3(grid), dim3(256), 0, st, p);
      return check_launch("tcn::gate_bwd_fused_kernel");
    case tcn::OP_PREACT_BWD:
      launch_pdl(tcn::preact_bwd_fused_kernel, dim3(grid), dim3(256), 0, st, p);
      return check_launch("tcn::preact_bwd_fused_kernel");
    case tcn::OP_GATE_FWD_STATS:
      launch_pdl(tcn::gate_fwd_fused_kernel, dim3(grid), dim3(256), 0, st, p);
      return check_launch("tcn::gate_fwd_fused_kernel");
    case tcn::OP_RESIDUAL_PREACT_FWD:
      launch_pdl(tcn::residual_preact_fused_kernel, dim3(grid), dim3(256), 0, st, p, flag);
      return check_launch("tcn::residual_preact_fused_kernel");
  }
  return fail(SELDQ_ERR_INVALID, "unknown TCN glue op %d", op);
}

}  // namespace seldq

// Bandwidth-bound kernels for the glue between the convolutions:
//
//   CNN block   model.py:276-283   conv -> BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d([p,1]) -> Dropout
//
// The convolution writes its output once, in 16 bits, in the tensor's own NCHW order.  The format is IEEE fp16
// (10-bit mantissa), not bf16: this tensor is only stored, never fed to the tensor cores, its values are O(1)
// pre-BatchNorm activations (the store saturates at +-65504), and the 8x finer rounding keeps the fused path's
// distance to the fp32 reference at that of the layer-by-layer bf16 path instead of using up the 2e-2 budget.  Then
//   bn_stats_kernel        per-channel sum / sum of squares (warp-shuffle + block reduction, one double
//                          atomicAdd pair per block)
//   bn_finalize_kernel     mean / rstd / affine coefficients per channel, running-statistics update
//   cnn_tail_fwd_kernel    BN-apply + ReLU + max over p frequency rows + dropout in one pass; emits the pooled
//                          activation directly as the NEXT convolution's channels-last bf16 operand (transposed
//                          through shared memory, conv_cl.h), optionally as fp32 NCHW, and one byte per pooled
//                          element (argmax row | keep flag) for the backward pass
//   cnn_tail_bwd_reduce    sum(dy), sum(dy * xhat) per channel (the two BatchNorm backward reductions), reading
//                          only the pooled gradient and the arg-max elements of the conv output
//   cnn_tail_bwd_apply     d(conv out) for every element, written straight into the two bf16 operand layouts the
//                          gradient kernels read (pitched NCHW for wgrad, channels-last for dgrad)
// so the full-resolution fp32 tensors of the reference (944 MB per sample after the first convolution, read
// and written by five separate PyTorch kernels) never exist.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "epilogue.h"
#include "launch.h"
#include "pdl.cuh"
#include "tile_cl.cuh"
#include "wgrad_first.h"

namespace seldq {
namespace epi {

__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __half* p) { return __half2float(*p); }

// grid (C, N, splits): block (c, n, z) reduces elements [z*chunk, (z+1)*chunk) of plane (n, c)
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ src, long long plane, int C,
                                                       long long chunk, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y;
  // chunks are visited from the END of the plane: the convolution that has just written this tensor finished with its
  // last rows, which are what the 126 MB L2 still holds of a 472 MB tensor (and the pooling kernel behind this one starts
  // at the first rows, which this pass then read last)
  const long long lo = (long long)(gridDim.z - 1 - blockIdx.z) * chunk;
  const long long hi = lo + chunk < plane ? lo + chunk : plane;
  const T* base = src + ((long long)n * C + c) * plane;
  float s1 = 0.f, s2 = 0.f;
  if (sizeof(T) == 2 && (plane & 7) == 0 && (chunk & 7) == 0) {
    const uint4* v = reinterpret_cast<const uint4*>(base + lo);
    const long long nv = (hi - lo) >> 3;
    for (long long i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 q = __ldg(v + i);
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        s1 += f.x + f.y;
        s2 += f.x * f.x + f.y * f.y;
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const float f = load_as_float(base + i);
      s1 += f;
      s2 += f * f;
    }
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, (double)s1);
    atomicAdd(sums + 2 * c + 1, (double)s2);
  }
}

// coef[c] = {a, b, mean, rstd} with BN(v) = a*v + b;  running statistics as nn.BatchNorm (momentum update,
// unbiased variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, double count, float eps, float momentum,
                                   float* running_mean, float* running_var, float* __restrict__ coef) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[2 * c] / count;
  double var = sums[2 * c + 1] / count - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float a = g * rstd;
  coef[4 * c + 0] = a;
  coef[4 * c + 1] = bt - (float)mean * a;
  coef[4 * c + 2] = (float)mean;
  coef[4 * c + 3] = rstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// One block: 64 padded channels x 32 w positions of one pooled row (n, h').
__global__ void __launch_bounds__(256) cnn_tail_fwd_kernel(const __grid_constant__ TailParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[64][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HP = p.H / p.pool;
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const uint32_t thresh = p.drop_p > 0.f ? (uint32_t)fminf(p.drop_p * 4294967296.f, 4294967295.f) : 0u;
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    long long r = blk;
    const int wt = (int)(r % p.tiles_w); r /= p.tiles_w;
    const int ct = (int)(r % p.tiles_c); r /= p.tiles_c;
    const int hp = (int)(r % HP);
    const int n = (int)(r / HP);
    const int w = wt * 32 + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cl = warp + 8 * k;
      const int cp = ct * 64 + cl;
      const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
      float z = 0.f;
      if (cp < p.Cp && ci < p.cc && w < p.W) {
        const int c = comp * p.cc + ci;
        const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
        const __half* src = p.y + (((long long)n * p.C + c) * p.H + (long long)hp * p.pool) * p.W + w;
        float best = -INFINITY;
        int arg = 0;
        __half yb = src[0];
        for (int j = 0; j < p.pool; ++j) {
          const __half yj = src[(long long)j * p.W];
          const float v = cf.x * __half2float(yj) + cf.y;
          if (v > best || v != v) { best = v; arg = j; yb = yj; }
        }
        const long long e = (((long long)n * p.C + c) * HP + hp) * p.W + w;
        if (p.ymax) p.ymax[e] = yb;
        bool keep = best > 0.f;
        if (keep && thresh) keep = dropout_keep_fast(dropout_key(seed, p.salt), e, thresh);
        z = keep ? best * scale : 0.f;
        p.idx[e] = (uint8_t)(arg | (keep ? 0x80 : 0));
        if (p.z32) p.z32[e] = z;
      }
      tile[cl][lane] = z;
    }
    __syncthreads();
    if (p.z_cl) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int wl = warp + 8 * k;
        const int ww = wt * 32 + wl;
        const int cp = ct * 64 + 2 * lane;
        if (ww < p.W && cp < p.Cp) {
          const __nv_bfloat162 o = __floats2bfloat162_rn(tile[2 * lane][wl], tile[2 * lane + 1][wl]);
          *reinterpret_cast<__nv_bfloat162*>(p.z_cl + (((long long)n * HP + hp) * p.W + ww) * p.Cp + cp) = o;
        }
      }
    }
    __syncthreads();
  }
}

// grid (C, N, splits) over the pooled plane H' x W of (n, c)
__global__ void __launch_bounds__(256) cnn_tail_bwd_reduce_kernel(const __grid_constant__ TailParams p, long long chunk,
                                                                 double* __restrict__ dsums) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y;
  const int HP = p.H / p.pool;
  const long long plane = (long long)HP * p.W;
  const long long lo = (long long)blockIdx.z * chunk;
  const long long hi = lo + chunk < plane ? lo + chunk : plane;
  const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const long long base = ((long long)n * p.C + c) * plane;
  const __half* y = p.y + ((long long)n * p.C + c) * (long long)p.H * p.W;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint8_t id = p.idx[base + i];
    if (id & 0x80) {
      const int hp = (int)(i / p.W), w = (int)(i - (long long)hp * p.W);
      const float g = __ldg(p.gz + base + i) * scale;
      const float yv = __half2float(p.ymax ? p.ymax[base + i] : y[((long long)hp * p.pool + (id & 7)) * p.W + w]);
      s1 += g;
      s2 += g * (yv - cf.z) * cf.w;
    }
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    atomicAdd(dsums + 2 * c, (double)s1);
    atomicAdd(dsums + 2 * c + 1, (double)s2);
  }
}

// same reduction from the arg-max values the forward pass kept (ymax): 8 pooled elements per thread, every
// stream (1-byte flags, fp32 gradient, bf16 values) read contiguously
__global__ void __launch_bounds__(256) cnn_tail_bwd_reduce_vec_kernel(const __grid_constant__ TailParams p,
                                                                     long long chunk, double* __restrict__ dsums) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x, n = blockIdx.y;
  const int HP = p.H / p.pool;
  const long long plane = (long long)HP * p.W;
  const long long lo = (long long)blockIdx.z * chunk;
  const long long hi = lo + chunk < plane ? lo + chunk : plane;
  const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const long long base = ((long long)n * p.C + c) * plane;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lo + (long long)threadIdx.x * 8; i < hi; i += 256 * 8) {
    const uint2 idv = __ldg(reinterpret_cast<const uint2*>(p.idx + base + i));
    if (((idv.x | idv.y) & 0x80808080u) == 0u) continue;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gz + base + i));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gz + base + i + 4));
    const uint4 yv = __ldg(reinterpret_cast<const uint4*>(p.ymax + base + i));
    const __half2* y2 = reinterpret_cast<const __half2*>(&yv);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t b = ((j < 4 ? idv.x : idv.y) >> (8 * (j & 3))) & 0xffu;
      if (b & 0x80u) {
        const float2 f = __half22float2(y2[j >> 1]);
        const float g = gg[j] * scale;
        s1 += g;
        s2 += g * (((j & 1) ? f.y : f.x) - cf.z) * cf.w;
      }
    }
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    atomicAdd(dsums + 2 * c, (double)s1);
    atomicAdd(dsums + 2 * c + 1, (double)s2);
  }
}

// dmean[c] = {sum(dy) / M, sum(dy * xhat) / M} in fp32 (keeps the FP64 pipe out of the per-element kernel)
__global__ void cnn_tail_bwd_finalize_kernel(const double* __restrict__ dsums, int C, double count,
                                             float2* __restrict__ dmean) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dmean[c] = make_float2((float)(dsums[2 * c] / count), (float)(dsums[2 * c + 1] / count));
}

// One block: 64 padded channels x 32 w positions of one full-resolution row (n, h).
__global__ void __launch_bounds__(256) cnn_tail_bwd_apply_kernel(const __grid_constant__ TailParams p,
                                                                const float2* __restrict__ dmean) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[64][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HP = p.H / p.pool;
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    long long r = blk;
    const int wt = (int)(r % p.tiles_w); r /= p.tiles_w;
    const int ct = (int)(r % p.tiles_c); r /= p.tiles_c;
    const int h = (int)(r % p.H);
    const int n = (int)(r / p.H);
    const int hp = h / p.pool, kk = h - hp * p.pool;
    const int w = wt * 32 + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cl = warp + 8 * k;
      const int cp = ct * 64 + cl;
      const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
      float d = 0.f;
      if (cp < p.Cp && ci < p.cc) {
        const int c = comp * p.cc + ci;
        const long long row = ((long long)n * p.C + c) * p.H + h;
        if (w < p.W) {
          const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
          const float xhat = (__half2float(p.y[row * p.W + w]) - cf.z) * cf.w;
          float g = 0.f;
          if (hp < HP) {
            const long long e = (((long long)n * p.C + c) * HP + hp) * p.W + w;
            const uint8_t id = p.idx[e];
            if ((id & 0x80) && (id & 7) == kk) g = __ldg(p.gz + e) * scale;
          }
          const float2 dm = __ldg(dmean + c);
          d = cf.x * (g - dm.x - xhat * dm.y);
        }
        if (p.d_t16 && w < p.pitch) p.d_t16[row * p.pitch + w] = __float2bfloat16_rn(d);
      }
      tile[cl][lane] = d;
    }
    __syncthreads();
    if (p.d_cl) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int wl = warp + 8 * k;
        const int ww = wt * 32 + wl;
        const int cp = ct * 64 + 2 * lane;
        if (ww < p.W && cp < p.Cp) {
          const __nv_bfloat162 o = __floats2bfloat162_rn(tile[2 * lane][wl], tile[2 * lane + 1][wl]);
          *reinterpret_cast<__nv_bfloat162*>(p.d_cl + (((long long)n * p.H + h) * p.W + ww) * p.Cp + cp) = o;
        }
      }
    }
    __syncthreads();
  }
}


// ---- vectorised variants (W a multiple of 8, which covers every shipped configuration) -----------------
// One block: 64 padded channels x 128 w.  Phase 1: a thread owns 8 consecutive w of one channel (16-byte
// loads / stores in the tensor's own NCHW order) and drops its results, as bf16, into a [w][64 ch] shared
// tile; phase 2 writes the tile as channels-last rows, 16 bytes per thread.
template <int POOL>   // pooled rows as a compile-time constant: all POOL loads of an item are in flight together
__global__ void __launch_bounds__(256, 3) cnn_tail_fwd_vec_kernel(const __grid_constant__ TailParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kVecTileW * kVecPitch];
  const int HP = p.H / p.pool;
  const unsigned long long seed = p.seed_ptr ? (unsigned long long)*p.seed_ptr : 0ULL;
  const uint32_t dkey = dropout_key(seed, p.salt);
  const uint32_t thresh = p.drop_p > 0.f ? (uint32_t)fminf(p.drop_p * 4294967296.f, 4294967295.f) : 0u;
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    long long r = blk;
    const int wt = (int)(r % p.tiles_w); r /= p.tiles_w;
    const int ct = (int)(r % p.tiles_c); r /= p.tiles_c;
    const int hp = (int)(r % HP);
    const int n = (int)(r / HP);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int item = threadIdx.x + 256 * k;
      const int cl = item >> 4, wl = (item & 15) * 8;
      const int w = wt * kVecTileW + wl;
      const int cp = ct * 64 + cl;
      const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
      float z[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = 0.f;
      if (cp < p.Cp && ci < p.cc && w < p.W) {
        const int c = comp * p.cc + ci;
        const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
        const __half* src = p.y + (((long long)n * p.C + c) * p.H + (long long)hp * p.pool) * p.W + w;
        // BN is affine, v = a y + b, so the pooled maximum of v is attained where t = sign(a) y is largest: the row
        // scan runs on packed fp16 pairs (sign flip, compare, max and arg-select are one instruction per PAIR; the
        // fp32 compare / select chain per element kept the ALU pipe 60 % busy, ncu) and BN is applied once per
        // pooled element -- the same a*y + b FMA as before, bit for bit.  First maximum wins (strict >); NaN propagates.
        const uint32_t sflip = cf.x < 0.f ? 0x80008000u : 0u;
        uint32_t best2[4], arg2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { best2[j] = 0xfc00fc00u; arg2[j] = 0u; }       // -inf, row 0
        uint4 rows[POOL];
#pragma unroll
        for (int q = 0; q < POOL; ++q) rows[q] = __ldg(reinterpret_cast<const uint4*>(src + (long long)q * p.W));
#pragma unroll
        for (int q = 0; q < POOL; ++q) {
          const uint32_t rw[4] = {rows[q].x, rows[q].y, rows[q].z, rows[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t tb = rw[j] ^ sflip;
            const __half2 t = *reinterpret_cast<const __half2*>(&tb), bo = *reinterpret_cast<const __half2*>(&best2[j]);
            const uint32_t gt = __hgt2_mask(t, bo);                                  // 0xffff per half where t > best
            const __half2 bn = __hmax2_nan(bo, t);
            best2[j] = *reinterpret_cast<const uint32_t*>(&bn);
            arg2[j] = (arg2[j] & ~gt) | ((uint32_t)(q * 0x00010001) & gt);
          }
        }
        float best[8], ybest[8];
        int arg[8];
        const float aa = fabsf(cf.x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&best2[j]));
          const uint32_t yb = best2[j] ^ sflip;
          const float2 yf = __half22float2(*reinterpret_cast<const __half2*>(&yb));
          best[2 * j] = fmaf(aa, tf.x, cf.y); best[2 * j + 1] = fmaf(aa, tf.y, cf.y);
          ybest[2 * j] = yf.x; ybest[2 * j + 1] = yf.y;
          arg[2 * j] = (int)(arg2[j] & 0xffffu); arg[2 * j + 1] = (int)(arg2[j] >> 16);
        }
        const long long e = (((long long)n * p.C + c) * HP + hp) * p.W + w;
        if (p.ymax) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2 b2 = __floats2half2_rn(ybest[2 * j], ybest[2 * j + 1]);   // exact: fp16 values
            o[j] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(p.ymax + e) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        uint32_t id[2] = {0u, 0u};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bool keep = best[j] > 0.f;
          if (keep && thresh) keep = dropout_keep_fast(dkey, e + j, thresh);
          z[j] = keep ? best[j] * scale : 0.f;
          id[j >> 2] |= (uint32_t)(arg[j] | (keep ? 0x80 : 0)) << (8 * (j & 3));
        }
        *reinterpret_cast<uint2*>(p.idx + e) = make_uint2(id[0], id[1]);
        if (p.z32) {
          *reinterpret_cast<float4*>(p.z32 + e) = make_float4(z[0], z[1], z[2], z[3]);
          *reinterpret_cast<float4*>(p.z32 + e + 4) = make_float4(z[4], z[5], z[6], z[7]);
        }
      }
      if (p.z_cl) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<__nv_bfloat16*>(tile + vec_tile_off(wl + j, cl * 2)) = __float2bfloat16_rn(z[j]);
      }
    }
    if (p.z_cl) {
      __syncthreads();
      vec_tile_store_cl(tile, p.z_cl, ((long long)n * HP + hp) * p.W, wt * kVecTileW, p.W, p.Cp, ct);
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) cnn_tail_bwd_apply_vec_kernel(const __grid_constant__ TailParams p,
                                                                    const float2* __restrict__ dmean) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[kVecTileW * kVecPitch];
  const int HP = p.H / p.pool;
  const float scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  for (long long blk = blockIdx.x; blk < p.total_blocks; blk += gridDim.x) {
    long long r = blk;
    const int wt = (int)(r % p.tiles_w); r /= p.tiles_w;
    const int ct = (int)(r % p.tiles_c); r /= p.tiles_c;
    const int h = (int)(r % p.H);
    const int n = (int)(r / p.H);
    const int hp = h / p.pool, kk = h - hp * p.pool;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int item = threadIdx.x + 256 * k;
      const int cl = item >> 4, wl = (item & 15) * 8;
      const int w = wt * kVecTileW + wl;
      const int cp = ct * 64 + cl;
      const int comp = cp / p.cpad, ci = cp - comp * p.cpad;
      uint32_t o[4] = {0u, 0u, 0u, 0u};
      if (cp < p.Cp && ci < p.cc && w < p.W) {
        const int c = comp * p.cc + ci;
        const long long row = ((long long)n * p.C + c) * p.H + h;
        const float4 cf = __ldg(reinterpret_cast<const float4*>(p.coef) + c);
        const float2 dm = __ldg(dmean + c);
        const uint4 yv = __ldg(reinterpret_cast<const uint4*>(p.y + row * p.W + w));
        const __half2* y2 = reinterpret_cast<const __half2*>(&yv);
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        if (hp < HP) {
          const long long e = (((long long)n * p.C + c) * HP + hp) * p.W + w;
          const uint2 idv = __ldg(reinterpret_cast<const uint2*>(p.idx + e));
          const uint32_t want = 0x80u | (uint32_t)kk;
          // only load the pooled gradient when at least one of the 8 elements selects this row
          uint32_t hit = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t b = ((j < 4 ? idv.x : idv.y) >> (8 * (j & 3))) & 0xffu;
            hit |= (uint32_t)((b & 0x87u) == want) << j;
          }
          if (hit) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gz + e));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gz + e + 4));
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = ((hit >> j) & 1u) ? gg[j] * scale : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(y2[j]);
          const float d0 = cf.x * (g[2 * j] - dm.x - (f.x - cf.z) * cf.w * dm.y);
          const float d1 = cf.x * (g[2 * j + 1] - dm.x - (f.y - cf.z) * cf.w * dm.y);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
          o[j] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        if (p.d_t16) *reinterpret_cast<uint4*>(p.d_t16 + row * p.pitch + w) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (p.d_cl) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint16_t*>(tile + vec_tile_off(wl + j, cl * 2)) = (uint16_t)(o[j >> 1] >> (16 * (j & 1)));
      }
    }
    if (p.d_cl) {
      __syncthreads();
      vec_tile_store_cl(tile, p.d_cl, ((long long)n * p.H + h) * p.W, wt * kVecTileW, p.W, p.Cp, ct);
      __syncthreads();
    }
  }
}

}  // namespace epi

// ---- host launchers ------------------------------------------------------------------------------------
static int grid_cap() { return 148 * 16; }

int launch_bn_stats(const void* src, int is_f16, int n, int c, long long plane, double* sums, cudaStream_t st) {
  if (n > 65535) return fail(SELDQ_ERR_UNSUPPORTED, "bn_stats: batch too large for the grid");
  // about 32 K elements per block, but never more than 64 splits of a plane
  long long splits = (plane + 32767) / 32768;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  long long chunk = (plane + splits - 1) / splits;
  chunk = (chunk + 7) & ~7LL;
  splits = (plane + chunk - 1) / chunk;
  dim3 grid((unsigned)c, (unsigned)n, (unsigned)splits);
  if (is_f16)
    launch_pdl(epi::bn_stats_kernel<__half>, dim3(grid), dim3(256), 0, st, reinterpret_cast<const __half*>(src), plane, c, chunk, sums);
  else
    launch_pdl(epi::bn_stats_kernel<float>, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float*>(src), plane, c, chunk, sums);
  return check_launch("bn_stats_kernel");
}

int launch_bn_finalize(const double* sums, const float* gamma, const float* beta, int c, double count, float eps,
                       float momentum, float* running_mean, float* running_var, float* coef, cudaStream_t st) {
  launch_pdl(epi::bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, st, sums, gamma, beta, c, count, eps, momentum,
             running_mean, running_var, coef);
  return check_launch("bn_finalize_kernel");
}

static bool tail_vec_ok(const epi::TailParams& p) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p.W % 8 == 0 && al16(p.y) && al16(p.idx) && al16(p.z32) && al16(p.z_cl) && al16(p.gz) && al16(p.d_t16) &&
         al16(p.d_cl) && (p.Cp % 8 == 0);
}

int launch_cnn_tail_fwd(epi::TailParams& p, cudaStream_t st) {
  if (tail_vec_ok(p)) {
    p.tiles_w = (p.W + epi::kVecTileW - 1) / epi::kVecTileW;
    p.tiles_c = (p.Cp + 63) / 64;
    p.total_blocks = (long long)p.tiles_w * p.tiles_c * (p.H / p.pool) * p.N;
    const long long blocks = p.total_blocks < grid_cap() ? p.total_blocks : grid_cap();
    const dim3 grid((unsigned)(blocks < 1 ? 1 : blocks));
    switch (p.pool) {
      case 1: launch_pdl(epi::cnn_tail_fwd_vec_kernel<1>, grid, dim3(256), 0, st, p); break;
      case 2: launch_pdl(epi::cnn_tail_fwd_vec_kernel<2>, grid, dim3(256), 0, st, p); break;
      case 3: launch_pdl(epi::cnn_tail_fwd_vec_kernel<3>, grid, dim3(256), 0, st, p); break;
      case 4: launch_pdl(epi::cnn_tail_fwd_vec_kernel<4>, grid, dim3(256), 0, st, p); break;
      case 5: launch_pdl(epi::cnn_tail_fwd_vec_kernel<5>, grid, dim3(256), 0, st, p); break;
      case 6: launch_pdl(epi::cnn_tail_fwd_vec_kernel<6>, grid, dim3(256), 0, st, p); break;
      case 7: launch_pdl(epi::cnn_tail_fwd_vec_kernel<7>, grid, dim3(256), 0, st, p); break;
      default: launch_pdl(epi::cnn_tail_fwd_vec_kernel<8>, grid, dim3(256), 0, st, p); break;
    }
    return check_launch("cnn_tail_fwd_vec_kernel");
  }
  p.tiles_w = (p.W + 31) / 32;
  p.tiles_c = (p.Cp + 63) / 64;
  p.total_blocks = (long long)p.tiles_w * p.tiles_c * (p.H / p.pool) * p.N;
  const long long blocks = p.total_blocks < grid_cap() ? p.total_blocks : grid_cap();
  launch_pdl(epi::cnn_tail_fwd_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks)), dim3(256), 0, st, p);
  return check_launch("cnn_tail_fwd_kernel");
}

int launch_cnn_tail_bwd(epi::TailParams& p, double* dsums, cudaStream_t st) {
  int rc = launch_cnn_tail_bwd_reduce(p, dsums, st);
  if (rc) return rc;
  float2* dmean = reinterpret_cast<float2*>(dsums + 2 * (size_t)p.C);      // third C doubles of the caller's buffer
  p.pitch = nchw16_pitch(p.W);
  if (tail_vec_ok(p)) {      // W % 8 == 0  =>  pitch == W
    p.tiles_w = (p.W + epi::kVecTileW - 1) / epi::kVecTileW;
    p.tiles_c = (p.Cp + 63) / 64;
    p.total_blocks = (long long)p.tiles_w * p.tiles_c * p.H * p.N;
    const long long blocks = p.total_blocks < grid_cap() ? p.total_blocks : grid_cap();
    launch_pdl(epi::cnn_tail_bwd_apply_vec_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks)), dim3(256), 0, st, p, dmean);
    return check_launch("cnn_tail_bwd_apply_vec_kernel");
  }
  p.tiles_w = (p.pitch + 31) / 32;
  p.tiles_c = (p.Cp + 63) / 64;
  p.total_blocks = (long long)p.tiles_w * p.tiles_c * p.H * p.N;
  const long long blocks = p.total_blocks < grid_cap() ? p.total_blocks : grid_cap();
  launch_pdl(epi::cnn_tail_bwd_apply_kernel, dim3((unsigned)(blocks < 1 ? 1 : blocks)), dim3(256), 0, st, p, dmean);
  return check_launch("cnn_tail_bwd_apply_kernel");
}

int launch_cnn_tail_bwd_reduce(epi::TailParams& p, double* dsums, cudaStream_t st) {
  const int HP = p.H / p.pool;
  const long long plane = (long long)HP * p.W;
  long long splits = (plane + 16383) / 16384;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  long long chunk = (plane + splits - 1) / splits;
  chunk = (chunk + 7) & ~7LL;
  splits = (plane + chunk - 1) / chunk;
  if (p.N > 65535) return fail(SELDQ_ERR_UNSUPPORTED, "cnn tail: batch too large for the grid");
  dim3 grid((unsigned)p.C, (unsigned)p.N, (unsigned)splits);
  if (p.ymax && tail_vec_ok(p) && chunk % 8 == 0 && (reinterpret_cast<uintptr_t>(p.ymax) & 15) == 0)
    launch_pdl(epi::cnn_tail_bwd_reduce_vec_kernel, dim3(grid), dim3(256), 0, st, p, chunk, dsums);
  else
    launch_pdl(epi::cnn_tail_bwd_reduce_kernel, dim3(grid), dim3(256), 0, st, p, chunk, dsums);
  int rc = check_launch("cnn_tail_bwd_reduce_kernel");
  if (rc) return rc;
  const double count = (double)p.N * p.H * p.W;
  float2* dmean = reinterpret_cast<float2*>(dsums + 2 * (size_t)p.C);      // third C doubles of the caller's buffer
  launch_pdl(epi::cnn_tail_bwd_finalize_kernel, dim3((p.C + 127) / 128), dim3(128), 0, st, dsums, p.C, count, dmean);
  return check_launch("cnn_tail_bwd_finalize_kernel");
}

}  // namespace seldq

// __global__ wrappers of the FP32 SIMT kernels (phase functions in conv_simt.cuh) plus the small
// bandwidth-bound helpers (bias gradient, fp32 -> bf16 cast).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "conv_simt.cuh"
#include "launch.h"

namespace seldq {
namespace simt {

__global__ void __launch_bounds__(NT) conv_simt_kernel(const __grid_constant__ ConvParams p) {
  __shared__ ConvShared s;
  ConvThread t;
  conv_init(t);
  const int ntap = p.g.KH * p.g.KW;
  for (int tap = 0; tap < ntap; ++tap)
    for (int r0 = 0; r0 < p.g.R; r0 += BK) {
      if (conv_chunk_is_zero(p.g, blockIdx.y * BM, r0)) continue;  // block-uniform
      conv_load(p, s, threadIdx.x, blockIdx.x, blockIdx.y, blockIdx.z, tap, r0);
      __syncthreads();
      conv_mac(s, t, threadIdx.x, conv_swap(p.g));
      __syncthreads();
    }
  conv_store(p, t, threadIdx.x, blockIdx.x, blockIdx.y, blockIdx.z);
}

__global__ void __launch_bounds__(NT) wgrad_simt_kernel(const __grid_constant__ WgradParams p) {
  __shared__ WgradShared s;
  WgradThread t;
  wgrad_init(t);
  const int ntap = p.g.KH * p.g.KW;
  const long long units = wgrad_units(p.g, blockIdx.y / ntap);
  const long long per = (units + p.splits - 1) / p.splits;
  const long long u0 = blockIdx.z * per;
  const long long u1 = u0 + per < units ? u0 + per : units;
  for (long long u = u0; u < u1; ++u) {
    wgrad_load(p, s, threadIdx.x, blockIdx.x, blockIdx.y, u);
    __syncthreads();
    wgrad_mac(s, t, threadIdx.x);
    __syncthreads();
  }
  if (u0 < u1) wgrad_store(p, t, threadIdx.x, blockIdx.x, blockIdx.y, [](float* a, float v) { atomicAdd(a, v); });
}

// gb[c] = sum over (n, h, w) of gy[n, c, h, w]; one block per channel, warp-shuffle reduction
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ gy, float* __restrict__ gb, int N,
                                                        int H, int W, long long sN, long long sC, long long sH,
                                                        long long sW, int accumulate) {
  const int c = blockIdx.x;
  const long long total = (long long)N * H * W;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    acc += gy[n * sN + c * sC + h * sH + w * sW];
  }
  __shared__ float part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) gb[c] = accumulate ? gb[c] + v : v;
  }
}

// fp32 -> bf16 (round to nearest even), 8 elements per thread per step
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        size_t n) {
  const size_t nvec = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    __nv_bfloat162 o[4];
    o[0] = __floats2bfloat162_rn(a.x, a.y);
    o[1] = __floats2bfloat162_rn(a.z, a.w);
    o[2] = __floats2bfloat162_rn(b.x, b.y);
    o[3] = __floats2bfloat162_rn(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<const uint4*>(o);
  }
  for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// fp32 (rows, w) -> bf16 mirror set [nshifts][rows][pitch]: copy k holds the row shifted right by
// shifts[k] elements with zeros shifted in and zero padding up to the pitch (conv_umma.h: MirrorSet).
// One thread produces one 16-byte chunk (8 elements) of one copy.
struct MirrorShifts { int v[8]; };
__global__ void __launch_bounds__(256) cast_bf16_mirror_kernel(const float* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dst, long long rows, int w,
                                                               int pitch, MirrorShifts shifts, int nshifts) {
  const int cpr = pitch >> 3;                                   // chunks per row
  const long long per_copy = rows * cpr;
  const long long total = per_copy * nshifts;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i / per_copy);
    const long long r = (i - k * per_copy) / cpr;
    const int chunk = (int)(i - k * per_copy - r * cpr);
    const int c0 = chunk * 8 - shifts.v[k];
    const float* s = src + r * w;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      v[j] = __float2bfloat16_rn((c >= 0 && c < w) ? __ldg(s + c) : 0.f);
    }
    reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<const uint4*>(v);
  }
}


// fp32 NCHW -> channels-last bf16 operand [n][h][w][Cp] (conv_cl.h: each of the nc components padded to cpad
// channels, pad channels zero) and, optionally, the pitched NCHW bf16 copy [n][c][h][pitch] that the
// weight-gradient kernel reads gy from (pad columns zero).  One block transposes a tile of 64 padded
// channels x 32 positions through shared memory: reads are coalesced along w, writes along the channels.
__global__ void __launch_bounds__(256) stage_operand_kernel(const float* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst_cl,
                                                            __nv_bfloat16* __restrict__ dst_t, int C, int H, int W,
                                                            int nc, int cc, int cpad, int Cp, int pitch, int tiles_w,
                                                            int tiles_c, long long total_blocks) {
  __shared__ float tile[64][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {
    long long r = blk;
    const int wt = (int)(r % tiles_w); r /= tiles_w;
    const int ct = (int)(r % tiles_c); r /= tiles_c;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const int w = wt * 32 + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cl = warp + 8 * k;                 // padded channel within the tile
      const int cp = ct * 64 + cl;
      const int comp = cp / cpad, ci = cp - comp * cpad;
      float v = 0.f;
      if (cp < Cp && ci < cc) {
        const int c = comp * cc + ci;
        const long long row = ((long long)n * C + c) * H + h;
        if (w < W) v = __ldg(src + row * W + w);
        if (dst_t && w < pitch) dst_t[row * pitch + w] = __float2bfloat16_rn(v);
      }
      tile[cl][lane] = v;
    }
    __syncthreads();
    if (dst_cl) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int wl = warp + 8 * k;
        const int ww = wt * 32 + wl;
        const int cp = ct * 64 + 2 * lane;
        if (ww < W && cp < Cp) {
          const __nv_bfloat162 o = __floats2bfloat162_rn(tile[2 * lane][wl], tile[2 * lane + 1][wl]);
          *reinterpret_cast<__nv_bfloat162*>(dst_cl + (((long long)n * H + h) * W + ww) * Cp + cp) = o;
        }
      }
    }
    __syncthreads();
  }
}

// Narrow single-component operands (the model input: C <= 16 channels padded to 16): a thread owns 4 consecutive w
// of every channel -- float4 loads, coalesced per channel row -- and writes the 4 channels-last rows (32 B each) as
// one contiguous 128-byte run; no shared memory.  W % 4 == 0.
template <int C>
__global__ void __launch_bounds__(256) stage_narrow_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst_cl,
                                                           long long planes /* N */, int H, int W) {
  const long long quads_per_plane = (long long)H * W / 4;
  const long long total = planes * quads_per_plane;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long n = q / quads_per_plane, pos = (q - n * quads_per_plane) * 4;     // pos = h * W + w
    const float* base = src + n * C * (long long)H * W + pos;
    float4 v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(base + (long long)c * H * W));
    uint4* out = reinterpret_cast<uint4*>(dst_cl + (n * (long long)H * W + pos) * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t wv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c0 = 2 * k, c1 = 2 * k + 1;
        const float a = c0 < C ? (j == 0 ? v[c0 < C ? c0 : 0].x : j == 1 ? v[c0 < C ? c0 : 0].y : j == 2 ? v[c0 < C ? c0 : 0].z : v[c0 < C ? c0 : 0].w) : 0.f;
        const float b = c1 < C ? (j == 0 ? v[c1 < C ? c1 : 0].x : j == 1 ? v[c1 < C ? c1 : 0].y : j == 2 ? v[c1 < C ? c1 : 0].z : v[c1 < C ? c1 : 0].w) : 0.f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
        wv[k] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      out[2 * j] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      out[2 * j + 1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
    }
  }
}

}  // namespace simt

// ---- host launchers ------------------------------------------------------------------------------
int launch_conv_simt(const simt::ConvParams& p, cudaStream_t st) {
  const ConvGeom& g = p.g;
  dim3 grid((g.OW + simt::BN - 1) / simt::BN, (g.P + simt::BM - 1) / simt::BM, (unsigned)(g.N * g.OH));
  if (grid.y > 65535 || grid.z > 65535) return fail(SELDQ_ERR_UNSUPPORTED, "problem too large for the fp32 grid");
  simt::conv_simt_kernel<<<grid, simt::NT, 0, st>>>(p);
  return check_launch("conv_simt_kernel");
}

int launch_wgrad_simt(simt::WgradParams& p, cudaStream_t st) {
  const ConvGeom& g = p.g;
  const int ntile = ((g.Oc + simt::WT - 1) / simt::WT) * ((g.Ic + simt::WT - 1) / simt::WT);
  const int ny = g.tab.nw * g.KH * g.KW;
  // enough splits to cover the chip about four times, never more than the shortest reduction
  long long min_units = simt::wgrad_units(g, 0);
  for (int e = 1; e < g.tab.nw; ++e) {
    const long long u = simt::wgrad_units(g, e);
    if (u < min_units) min_units = u;
  }
  long long splits = (4LL * 148 + (long long)ntile * ny - 1) / ((long long)ntile * ny);
  if (splits > min_units) splits = min_units;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.splits = (int)splits;
  dim3 grid(ntile, ny, (unsigned)splits);
  simt::wgrad_simt_kernel<<<grid, simt::NT, 0, st>>>(p);
  return check_launch("wgrad_simt_kernel");
}

int launch_bias_grad(const float* gy, float* gb, int C, int N, int H, int W, long long sN, long long sC, long long sH,
                     long long sW, int accumulate, cudaStream_t st) {
  simt::bias_grad_kernel<<<C, 256, 0, st>>>(gy, gb, N, H, W, sN, sC, sH, sW, accumulate);
  return check_launch("bias_grad_kernel");
}

int launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t st) {
  if (n == 0) return SELDQ_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16 needs 16-byte aligned buffers");
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  simt::cast_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return check_launch("cast_bf16_kernel");
}

int launch_cast_bf16_mirror(const float* src, void* dst, long long rows, int w, int pitch, const int* shifts,
                            int nshifts, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(dst) & 15) return fail(SELDQ_ERR_INVALID, "bf16 mirror buffer must be 16-byte aligned");
  simt::MirrorShifts s{};
  for (int i = 0; i < nshifts; ++i) s.v[i] = shifts[i];
  long long blocks = (rows * (pitch / 8) * nshifts + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  simt::cast_bf16_mirror_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), rows, w,
                                                                 pitch, s, nshifts);
  return check_launch("cast_bf16_mirror_kernel");
}


int launch_stage_operand(const float* src, void* dst_cl, void* dst_nchw16, const cl::OperandLayout& l, int n, int c,
                         int h, int w, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(dst_cl) & 15) || (reinterpret_cast<uintptr_t>(dst_nchw16) & 15))
    return fail(SELDQ_ERR_INVALID, "bf16 operand buffers must be 16-byte aligned");
  if (l.nc == 1 && l.Cp == 16 && !dst_nchw16 && dst_cl && (w & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (c == 8 || c == 16)) {
    const long long total = (long long)n * h * w / 4;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (c == 8)
      simt::stage_narrow_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_cl), n, h, w);
    else
      simt::stage_narrow_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_cl), n, h, w);
    return check_launch("stage_narrow_kernel");
  }
  const int pitch = nchw16_pitch(w);
  const int tiles_w = (pitch + 31) / 32, tiles_c = (l.Cp + 63) / 64;
  const long long total = (long long)tiles_w * tiles_c * h * n;
  long long blocks = total < 148LL * 32 ? total : 148LL * 32;
  if (blocks < 1) blocks = 1;
  simt::stage_operand_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_cl),
                                                              reinterpret_cast<__nv_bfloat16*>(dst_nchw16), c, h, w,
                                                              l.nc, l.cc, l.cpad, l.Cp, pitch, tiles_w, tiles_c, total);
  return check_launch("stage_operand_kernel");
}

}  // namespace seldq

// FP32 (FFMA) implicit-GEMM kernels for the quaternion / dual-quaternion convolution and linear
// layers: forward, dgrad, wgrad and the bias gradient.  This is the SELDQ_PREC_FP32 path (parity
// gate rel 1e-4); the tensor-core path lives in conv_umma.cu.
//
// The expanded 4x/8x weight (quaternion_ops.py:131-135, dual_quaternion_ops.py:122-140) is never
// built: tiles of it are gathered on the fly from the compact weights through BlockTable, and
// tiles that fall entirely into the dual-quaternion zero block are skipped.
//
// Every kernel body is written as __host__ __device__ "phases" over (Shared, Thread) state so
// that tests/host_emul can run the very same index arithmetic on the CPU.
#pragma once
#include "common.cuh"

namespace seldq {
namespace simt {

constexpr int BM = 64;   // out channels per block
constexpr int BN = 64;   // out positions (along W) per block
constexpr int BK = 16;   // reduce channels per step
constexpr int NT = 256;  // threads per block

struct ConvParams {
  ConvGeom g;
  const float* in;
  float* out;
  const float* w[8];
  const float* bias;  // per out channel, fwd only (may be null)
};

struct ConvShared {
  float Ws[BK][BM + 1];
  float Xs[BK][BN + 1];
};
struct ConvThread {
  float acc[4][4];
};

SELDQ_HD void conv_init(ConvThread& t) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) t.acc[i][j] = 0.f;
}

// true when every (out comp, in comp) pair touched by this (p tile, r chunk) is a zero block
SELDQ_HD bool conv_chunk_is_zero(const ConvGeom& g, int p0, int r0) {
  if (g.tab.nc == 1) return false;
  const int pc = g.transposed ? g.Ic : g.Oc, rc = g.transposed ? g.Oc : g.Ic;
  const int pl = p0 / pc, ph = (imin(p0 + BM, g.P) - 1) / pc;
  const int rl = r0 / rc, rh = (imin(r0 + BK, g.R) - 1) / rc;
  for (int x = pl; x <= ph; ++x)
    for (int y = rl; y <= rh; ++y) {
      const int e = g.transposed ? g.tab.widx[y][x] : g.tab.widx[x][y];
      if (e >= 0) return false;
    }
  return true;
}

SELDQ_HD void conv_load(const ConvParams& p, ConvShared& s, int tid, int bx, int by, int bz, int tap, int r0) {
  const ConvGeom& g = p.g;
  const int p0 = by * BM, ow0 = bx * BN;
  const int n = bz / g.OH, oh = bz - n * g.OH;
  const int kh = tap / g.KW, kw = tap - kh * g.KW;
#pragma unroll
  for (int j = 0; j < (BK * BM) / NT; ++j) {
    const int idx = tid + j * NT;
    const int k = idx / BM, m = idx - k * BM;
    float v = 0.f;
    if (p0 + m < g.P && r0 + k < g.R) v = expanded_weight(g, p.w, p0 + m, r0 + k, tap);
    s.Ws[k][m] = v;
  }
  int ih;
  const bool hok = map_pos(g.transposed, oh, kh, g.sh, g.ph, g.dh, g.IH, &ih);
  // consecutive threads walk along the contiguous axis of `in`: positions for NCW / NCHW tensors, channels for
  // the (rows, features) matrices of the linear layers
  const bool chan_fast = g.in_sC == 1 && g.in_sW != 1;
#pragma unroll
  for (int j = 0; j < (BK * BN) / NT; ++j) {
    const int idx = tid + j * NT;
    const int k = chan_fast ? idx % BK : idx / BN, q = chan_fast ? idx / BK : idx - (idx / BN) * BN;
    float v = 0.f;
    int iw;
    if (hok && r0 + k < g.R && ow0 + q < g.OW &&
        map_pos(g.transposed, ow0 + q, kw, g.sw, g.pw, g.dw, g.IW, &iw))
      v = p.in[n * g.in_sN + (long long)(r0 + k) * g.in_sC + ih * g.in_sH + iw * g.in_sW];
    s.Xs[k][q] = v;
  }
}

SELDQ_HD bool conv_swap(const ConvGeom& g) { return g.out_sC == 1 && g.out_sW != 1; }

// acc[i][j]: out channel ty + 16 i, position tx + 16 j.  swap: the 16 consecutive lanes index channels instead of
// positions (output with unit channel stride, i.e. the linear layers), so that stores stay coalesced
SELDQ_HD void conv_mac(const ConvShared& s, ConvThread& t, int tid, bool swap = false) {
  const int ty = swap ? (tid & 15) : (tid >> 4), tx = swap ? (tid >> 4) : (tid & 15);
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = s.Ws[k][ty + 16 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = s.Xs[k][tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) t.acc[i][j] = fmaf(a[i], b[j], t.acc[i][j]);
  }
}

SELDQ_HD void conv_store(const ConvParams& p, const ConvThread& t, int tid, int bx, int by, int bz) {
  const ConvGeom& g = p.g;
  const bool swap = conv_swap(g);
  const int ty = swap ? (tid & 15) : (tid >> 4), tx = swap ? (tid >> 4) : (tid & 15);
  const int n = bz / g.OH, oh = bz - n * g.OH;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = by * BM + ty + 16 * i;
    if (c >= g.P) continue;
    const float bv = p.bias ? p.bias[c] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ow = bx * BN + tx + 16 * j;
      if (ow < g.OW)
        p.out[n * g.out_sN + (long long)c * g.out_sC + oh * g.out_sH + ow * g.out_sW] = t.acc[i][j] + bv;
    }
  }
}

// ---- wgrad ----------------------------------------------------------------------------------------
// gW_e[o, i, tap] = sum_{(a,b): widx[a][b]=e} sign[a][b] * sum_{n, oh, ow} gy[n, a*Oc+o, oh, ow] * x[n, b*Ic+i, ih, iw]
// (Appendix B of SURVEY.md).  Block = 32 x 32 tile of one compact tensor for one tap; the reduction
// over (pair, n, oh, ow-chunk) is split across gridDim.z and combined with atomicAdd.
constexpr int WT = 32;   // tile edge (o and i)
constexpr int WK = 32;   // positions per step

struct WgradParams {
  ConvGeom g;          // forward orientation (transposed = 0); "in" strides = x, "out" strides = gy
  const float* x;
  const float* gy;
  float* gw[8];
  int splits;
};
struct WgradShared {
  float Gs[WK][WT + 1];
  float Xs[WK][WT + 1];
};
struct WgradThread {
  float acc[2][2];
};

SELDQ_HD void wgrad_init(WgradThread& t) { t.acc[0][0] = t.acc[0][1] = t.acc[1][0] = t.acc[1][1] = 0.f; }

// number of (a,b) pairs feeding compact weight e, and the idx-th of them
SELDQ_HD int wgrad_pair(const BlockTable& tab, int e, int idx, int* a_out, int* b_out) {
  int cnt = 0;
  for (int a = 0; a < tab.nc; ++a)
    for (int b = 0; b < tab.nc; ++b)
      if (tab.widx[a][b] == e) {
        if (cnt == idx) { *a_out = a; *b_out = b; }
        ++cnt;
      }
  return cnt;
}

// one reduction step: unit u enumerates (pair, n, oh, ow-chunk)
SELDQ_HD void wgrad_load(const WgradParams& p, WgradShared& s, int tid, int bx, int by, long long u) {
  const ConvGeom& g = p.g;
  const int ntile_i = (g.Ic + WT - 1) / WT;
  const int o0 = (bx / ntile_i) * WT, i0 = (bx % ntile_i) * WT;
  const int ntap = g.KH * g.KW;
  const int e = by / ntap, tap = by - e * ntap;
  const int kh = tap / g.KW, kw = tap - kh * g.KW;
  const int nchunk = (g.OW + WK - 1) / WK;
  const int chunk = (int)(u % nchunk); u /= nchunk;
  const int oh = (int)(u % g.OH); u /= g.OH;
  const int n = (int)(u % g.N); u /= g.N;
  int a = 0, b = 0;
  wgrad_pair(g.tab, e, (int)u, &a, &b);
  const float sg = g.tab.sign[a][b] > 0 ? 1.f : -1.f;
  int ih;
  const bool hok = map_pos(0, oh, kh, g.sh, g.ph, g.dh, g.IH, &ih);
#pragma unroll
  for (int j = 0; j < (WK * WT) / NT; ++j) {
    const int idx = tid + j * NT;
    // consecutive threads walk along the contiguous axis: W for NCW / NCHW, channels for the linear layers
    const bool chan_fast = g.out_sC == 1 && g.out_sW != 1;
    const int c = chan_fast ? idx % WT : idx / WK, q = chan_fast ? idx / WT : idx - (idx / WK) * WK;
    const int ow = chunk * WK + q;
    float gv = 0.f, xv = 0.f;
    if (ow < g.OW) {
      if (o0 + c < g.Oc)
        gv = sg * p.gy[n * g.out_sN + (long long)(a * g.Oc + o0 + c) * g.out_sC + oh * g.out_sH + ow * g.out_sW];
      int iw;
      if (hok && i0 + c < g.Ic && map_pos(0, ow, kw, g.sw, g.pw, g.dw, g.IW, &iw))
        xv = p.x[n * g.in_sN + (long long)(b * g.Ic + i0 + c) * g.in_sC + ih * g.in_sH + iw * g.in_sW];
    }
    s.Gs[q][c] = gv;
    s.Xs[q][c] = xv;
  }
}

SELDQ_HD void wgrad_mac(const WgradShared& s, WgradThread& t, int tid) {
  const int to = tid >> 4, ti = tid & 15;
#pragma unroll
  for (int k = 0; k < WK; ++k) {
    const float g0 = s.Gs[k][to], g1 = s.Gs[k][to + 16];
    const float x0 = s.Xs[k][ti], x1 = s.Xs[k][ti + 16];
    t.acc[0][0] = fmaf(g0, x0, t.acc[0][0]);
    t.acc[0][1] = fmaf(g0, x1, t.acc[0][1]);
    t.acc[1][0] = fmaf(g1, x0, t.acc[1][0]);
    t.acc[1][1] = fmaf(g1, x1, t.acc[1][1]);
  }
}

// total reduction units for compact weight e
SELDQ_HD long long wgrad_units(const ConvGeom& g, int e) {
  int a, b;
  const int npair = wgrad_pair(g.tab, e, -1, &a, &b);
  return (long long)npair * g.N * g.OH * ((g.OW + WK - 1) / WK);
}

template <class AtomicAdd>
SELDQ_HD void wgrad_store(const WgradParams& p, const WgradThread& t, int tid, int bx, int by, AtomicAdd add) {
  const ConvGeom& g = p.g;
  const int ntile_i = (g.Ic + WT - 1) / WT;
  const int o0 = (bx / ntile_i) * WT, i0 = (bx % ntile_i) * WT;
  const int ntap = g.KH * g.KW;
  const int e = by / ntap, tap = by - e * ntap;
  const int to = tid >> 4, ti = tid & 15;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int o = o0 + to + 16 * a, i = i0 + ti + 16 * b;
      if (o < g.Oc && i < g.Ic)
        add(&p.gw[e][(long long)o * g.wsO + (long long)i * g.wsI + (long long)tap * g.wsT], t.acc[a][b]);
    }
}

}  // namespace simt
}  // namespace seldq

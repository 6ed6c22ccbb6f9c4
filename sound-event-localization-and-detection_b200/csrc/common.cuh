// Shared host/device definitions of libseldq.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/seldq.h"

#if defined(__CUDACC__)
#define SELDQ_HD __host__ __device__ __forceinline__
#else
#define SELDQ_HD inline
#endif

namespace seldq {

SELDQ_HD int imin(int a, int b) { return a < b ? a : b; }
SELDQ_HD int imax(int a, int b) { return a > b ? a : b; }

// ---- error string (thread local; seldq.h: "re-entrant") -------------------------------------
inline char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// ---- block structure of the expanded weight ------------------------------------------------------
// widx[a][b] = which compact weight (0..7, reference order r,i,j,k,r_2,i_2,j_2,k_2) maps input
// component b to output component a, sign[a][b] = +-1; widx = -1 is a structural zero block.
// Reference: quaternion_ops.py:131-135 (Q), dual_quaternion_ops.py:122-140 (DQ conv),
// dual_quaternion_ops.py:170-188 (DQ linear: same matrix used from the right => transposed table).
struct BlockTable {
  int nc;
  int nw;
  int8_t widx[8][8];
  int8_t sign[8][8];
};

inline bool make_block_table(int algebra, bool dq_linear, BlockTable* t) {
  static const int8_t S[4][4] = {{+1, -1, -1, -1}, {+1, +1, -1, +1}, {+1, +1, +1, -1}, {+1, -1, +1, +1}};
  memset(t, 0, sizeof(*t));
  for (int a = 0; a < 8; ++a)
    for (int b = 0; b < 8; ++b) t->widx[a][b] = -1;
  if (algebra == SELDQ_ALG_Q_LINEAR_IO) algebra = SELDQ_ALG_Q;      // quaternion_linear's table is the convolution's
  if (algebra == SELDQ_ALG_REAL) {
    t->nc = 1; t->nw = 1; t->widx[0][0] = 0; t->sign[0][0] = 1;
    return true;
  }
  if (algebra == SELDQ_ALG_Q) {
    t->nc = 4; t->nw = 4;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) { t->widx[a][b] = (int8_t)(a ^ b); t->sign[a][b] = S[a][b]; }
    return true;
  }
  if (algebra == SELDQ_ALG_DQ_LINEAR || algebra == SELDQ_ALG_DQ_LINEAR_IO) { algebra = SELDQ_ALG_DQ; dq_linear = true; }
  if (algebra == SELDQ_ALG_DQ) {
    t->nc = 8; t->nw = 8;
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b) {
        const int ha = a >> 2, ca = a & 3, hb = b >> 2, cb = b & 3;
        if (!dq_linear) {
          if (ha == hb) { t->widx[a][b] = (int8_t)(ca ^ cb); t->sign[a][b] = S[ca][cb]; }
          else if (ha == 1 && hb == 0) { t->widx[a][b] = (int8_t)(4 + (ca ^ cb)); t->sign[a][b] = S[ca][cb]; }
        } else {
          if (ha == hb) { t->widx[a][b] = (int8_t)(ca ^ cb); t->sign[a][b] = S[cb][ca]; }
          else if (ha == 0 && hb == 1) { t->widx[a][b] = (int8_t)(4 + (ca ^ cb)); t->sign[a][b] = S[cb][ca]; }
        }
      }
    return true;
  }
  return false;
}

// ---- geometry shared by the SIMT kernels ---------------------------------------------------------
// One struct describes forward, dgrad (transposed = 1) and the linear layers (generic strides).
//   "out"  = tensor being produced: y (fwd) or gx (dgrad);  P channels, OH x OW positions
//   "in"   = tensor being read:     x (fwd) or gy (dgrad);  R channels, IH x IW positions
struct ConvGeom {
  int N;
  int P, R;        // expanded channel counts of out / in
  int Oc, Ic;      // per-component out/in channels in the FORWARD sense (compact weight is Oc x Ic x taps)
  int OH, OW, IH, IW;
  int KH, KW;
  int sh, sw, ph, pw, dh, dw;
  int transposed;
  long long in_sN, in_sC, in_sH, in_sW;
  long long out_sN, out_sC, out_sH, out_sW;
  int wsO, wsI, wsT;  // compact weight element (o, i, tap) lives at o*wsO + i*wsI + tap*wsT
  BlockTable tab;
};

// position of the `in` sample feeding output position o along one axis for tap k; false = padding
SELDQ_HD bool map_pos(int transposed, int o, int k, int s, int p, int d, int in_extent, int* i) {
  if (!transposed) {
    const int v = o * s - p + k * d;
    *i = v;
    return v >= 0 && v < in_extent;
  }
  const int v = o + p - k * d;          // = y * s
  if (v < 0) return false;
  const int q = v / s;
  *i = q;
  return (q * s == v) && q < in_extent;
}

// expanded weight element between out channel p and in channel r (pass orientation), tap t
SELDQ_HD float expanded_weight(const ConvGeom& g, const float* const* w, int p, int r, int t) {
  int a, o, b, i;
  if (!g.transposed) { a = p / g.Oc; o = p - a * g.Oc; b = r / g.Ic; i = r - b * g.Ic; }
  else               { b = p / g.Ic; i = p - b * g.Ic; a = r / g.Oc; o = r - a * g.Oc; }
  const int e = g.tab.widx[a][b];
  if (e < 0) return 0.f;
  const float v = w[e][(long long)o * g.wsO + (long long)i * g.wsI + (long long)t * g.wsT];
  return g.tab.sign[a][b] > 0 ? v : -v;
}

}  // namespace seldq

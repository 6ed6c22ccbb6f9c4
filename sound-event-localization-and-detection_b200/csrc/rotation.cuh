// Rotation kernels of the quaternion layers -- element arithmetic shared by the device kernels
// (rotation.cu) and the host emulation (tests/host_emul/emul.cpp).
//
// quaternion_conv_rotation, quaternion_transpose_conv_rotation and quaternion_linear_rotation
// (quaternion/quaternion_ops.py:174-232, :235-295, :330-388) turn the four compact tensors (r, i, j, k) -- shape
// (d0, d1, taps...) -- into ONE real weight of nc x nc blocks, nc = 3 (quaternion_format False) or 4 (True: block row 0
// and block column 0 are zero), and hand it to a plain real convolution / matrix product.  With n = |q| and f = 2 n
// (the reference multiplies by the norm where the textbook formula divides by its square; restated as written) the
// 3 x 3 core, block row a / block column b, is
//        [ 1 - f (jj + kk)    f ij + f rk        f ik - f rj      ]
//        [ f ij - f rk        1 - f (ii + kk)    f jk + f ri      ]
//        [ f ik + f rj        f jk - f ri        1 - f (ii + jj)  ]
// Block (a, b) of element (x0, x1, tap) lands at row a d0 + x0, column b d1 + x1 of the (nc d0, nc d1, taps) weight;
// `transpose_out` writes the (nc d1, nc d0, taps) transpose instead (the linear variant runs as a 1 x 1 convolution,
// whose weight is (out, in)).  The forward follows the reference's fp32 order of operations with IEEE-rounded products,
// sums and square root and no fused multiply-adds: bit-identical to the oracle's float32 restatement
// (oracle/algebra.py rotation_weight); torch's vectorised CPU square root is 1 ulp off on some elements, so against
// the reference's own float32 weight the agreement is to the last bit or the one before it.
#pragma once
#include <cmath>

#include "common.cuh"

namespace seldq {
namespace rot {

struct RotGeom {
  long long d0, d1, taps;
  int nc;              // 3 or 4
  int transpose_out;
};

#if defined(__CUDA_ARCH__)
SELDQ_HD float r_mul(float a, float b) { return __fmul_rn(a, b); }
SELDQ_HD float r_add(float a, float b) { return __fadd_rn(a, b); }
SELDQ_HD float r_sub(float a, float b) { return __fsub_rn(a, b); }
SELDQ_HD float r_sqrt(float a) { return __fsqrt_rn(a); }
#else
SELDQ_HD float r_mul(float a, float b) { volatile float v = a * b; return v; }
SELDQ_HD float r_add(float a, float b) { volatile float v = a + b; return v; }
SELDQ_HD float r_sub(float a, float b) { volatile float v = a - b; return v; }
SELDQ_HD float r_sqrt(float a) { return sqrtf(a); }
#endif

// e[a][b]: the 3 x 3 core above (quaternion_ops.py:188-220), every product and sum rounded as torch rounds them
SELDQ_HD void rotation_entries(float r, float i, float j, float k, float e[3][3]) {
  const float n = r_sqrt(r_add(r_add(r_add(r_mul(r, r), r_mul(i, i)), r_mul(j, j)), r_mul(k, k)));
  const float f = r_mul(2.0f, n);
  const float si = r_mul(f, r_mul(i, i)), sj = r_mul(f, r_mul(j, j)), sk = r_mul(f, r_mul(k, k));
  const float fr = r_mul(f, r), fi = r_mul(f, i), fj = r_mul(f, j);
  const float ri = r_mul(fr, i), rj = r_mul(fr, j), rk = r_mul(fr, k);
  const float ij = r_mul(fi, j), ik = r_mul(fi, k), jk = r_mul(fj, k);
  e[0][0] = r_sub(1.0f, r_add(sj, sk)); e[1][0] = r_sub(ij, rk);               e[2][0] = r_add(ik, rj);
  e[0][1] = r_add(ij, rk);              e[1][1] = r_sub(1.0f, r_add(si, sk)); e[2][1] = r_sub(jk, ri);
  e[0][2] = r_sub(ik, rj);              e[1][2] = r_add(jk, ri);              e[2][2] = r_sub(1.0f, r_add(si, sj));
}

// gradient of the four compact values given the gradient G[a][b] of the nine entries:
//   entry = delta_ab + f P_ab(q),  d entry / d q_c = (2 q_c / n) P_ab + f dP_ab / dq_c
// (n = 0 gives 0 / 0 = NaN, as torch.sqrt's backward does for the reference)
SELDQ_HD void rotation_entries_bwd(float r, float i, float j, float k, const float G[3][3], float g[4]) {
  const float n = sqrtf(r * r + i * i + j * j + k * k), f = 2.0f * n;
  const float P00 = -(j * j + k * k), P11 = -(i * i + k * k), P22 = -(i * i + j * j);
  const float S = G[0][0] * P00 + G[1][1] * P11 + G[2][2] * P22 + G[1][0] * (i * j - r * k) + G[2][0] * (i * k + r * j) +
                  G[0][1] * (i * j + r * k) + G[2][1] * (j * k - r * i) + G[0][2] * (i * k - r * j) +
                  G[1][2] * (j * k + r * i);
  const float s01 = G[1][0] + G[0][1], s02 = G[2][0] + G[0][2], s12 = G[2][1] + G[1][2];
  const float a01 = G[0][1] - G[1][0], a20 = G[2][0] - G[0][2], a12 = G[1][2] - G[2][1];
  const float Dr = k * a01 + j * a20 + i * a12;
  const float Di = j * s01 + k * s02 + r * a12 - 2.0f * i * (G[1][1] + G[2][2]);
  const float Dj = i * s01 + r * a20 + k * s12 - 2.0f * j * (G[0][0] + G[2][2]);
  const float Dk = r * a01 + i * s02 + j * s12 - 2.0f * k * (G[0][0] + G[1][1]);
  const float c = 2.0f * S / n;
  g[0] = c * r + f * Dr;
  g[1] = c * i + f * Di;
  g[2] = c * j + f * Dj;
  g[3] = c * k + f * Dk;
}

// element `idx` of the launch -> (x0, x1, tap).  The fastest index is the one that is contiguous in the EXPANDED
// weight (tap, then x1; with transpose_out: tap, then x0), so the nc x nc stores of a warp are coalesced.
SELDQ_HD void rot_element(const RotGeom& g, long long idx, long long* x0, long long* x1, long long* t) {
  *t = idx % g.taps; idx /= g.taps;
  if (g.transpose_out) { *x0 = idx % g.d0; *x1 = idx / g.d0; }
  else                 { *x1 = idx % g.d1; *x0 = idx / g.d1; }
}
SELDQ_HD long long rot_compact_offset(const RotGeom& g, long long x0, long long x1, long long t) {
  return (x0 * g.d1 + x1) * g.taps + t;
}
// offset of block (a, b) -- 0-based in the nc x nc block grid -- of element (x0, x1, tap) in the expanded weight
SELDQ_HD long long rot_expanded_offset(const RotGeom& g, int a, int b, long long x0, long long x1, long long t) {
  const long long R = g.nc * g.d0, C = g.nc * g.d1, row = a * g.d0 + x0, col = b * g.d1 + x1;
  return g.transpose_out ? (col * R + row) * g.taps + t : (row * C + col) * g.taps + t;
}

SELDQ_HD void rot_fwd_element(const RotGeom& g, long long idx, const float* r, const float* i, const float* j, const float* k,
                              float* out) {
  long long x0, x1, t;
  rot_element(g, idx, &x0, &x1, &t);
  const long long c = rot_compact_offset(g, x0, x1, t);
  float e[3][3];
  rotation_entries(r[c], i[c], j[c], k[c], e);
  const int z = g.nc - 3;                 // quaternion_format: the zero block row / column come first
  for (int a = 0; a < g.nc; ++a)
    for (int b = 0; b < g.nc; ++b)
      out[rot_expanded_offset(g, a, b, x0, x1, t)] = (a < z || b < z) ? 0.0f : e[a - z][b - z];
}

SELDQ_HD void rot_bwd_element(const RotGeom& g, long long idx, const float* r, const float* i, const float* j, const float* k,
                              const float* gout, float* gr, float* gi, float* gj, float* gk) {
  long long x0, x1, t;
  rot_element(g, idx, &x0, &x1, &t);
  const long long c = rot_compact_offset(g, x0, x1, t);
  const int z = g.nc - 3;
  float G[3][3], gq[4];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) G[a][b] = gout[rot_expanded_offset(g, a + z, b + z, x0, x1, t)];
  rotation_entries_bwd(r[c], i[c], j[c], k[c], G, gq);
  gr[c] = gq[0]; gi[c] = gq[1]; gj[c] = gq[2]; gk[c] = gq[3];
}

// ---- Hamilton product, normalisation and exponential of quaternion-valued tensors -------------------------------
// (quaternion_ops.py:467-507 = dual_quaternion_ops.py:374-414; dual_quaternion_ops.py:206-246.)  A tensor is read as
// (outer, 4, m): component c of quaternion (o, x) at o * 4 m + c * m + x -- the reference's get_r / get_i / get_j /
// get_k slices of dimension 1.
SELDQ_HD void hamilton(const float a[4], const float b[4], float o[4]) {
  // sums in the reference's order: ((t0 +- t1) +- t2) +- t3 of the four products of a row
  o[0] = r_sub(r_sub(r_sub(r_mul(a[0], b[0]), r_mul(a[1], b[1])), r_mul(a[2], b[2])), r_mul(a[3], b[3]));
  o[1] = r_sub(r_add(r_add(r_mul(a[0], b[1]), r_mul(a[1], b[0])), r_mul(a[2], b[3])), r_mul(a[3], b[2]));
  o[2] = r_add(r_add(r_sub(r_mul(a[0], b[2]), r_mul(a[1], b[3])), r_mul(a[2], b[0])), r_mul(a[3], b[1]));
  o[3] = r_add(r_sub(r_add(r_mul(a[0], b[3]), r_mul(a[1], b[2])), r_mul(a[2], b[1])), r_mul(a[3], b[0]));
}

// q_normalize (dual_quaternion_ops.py:206-223): q / sqrt(|q|^2 + 1e-4)
SELDQ_HD void qnormalize(const float q[4], float o[4]) {
  const float n = r_sqrt(r_add(r_add(r_add(r_add(r_mul(q[0], q[0]), r_mul(q[1], q[1])), r_mul(q[2], q[2])), r_mul(q[3], q[3])),
                               0.0001f));
  for (int c = 0; c < 4; ++c) o[c] = q[c] / n;
}
SELDQ_HD void qnormalize_bwd(const float q[4], const float g[4], float gq[4]) {
  const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + 0.0001f);
  const float dot = (g[0] * q[0] + g[1] * q[1] + g[2] * q[2] + g[3] * q[3]) / (n * n * n);
  for (int c = 0; c < 4; ++c) gq[c] = g[c] / n - q[c] * dot;
}

// quaternion_exp (dual_quaternion_ops.py:227-246): e^r (cos |v|', v / |v|' sin |v|'), |v|' = |v| + 1e-4
SELDQ_HD void qexp(const float q[4], float o[4]) {
  const float nv = r_add(r_sqrt(r_add(r_add(r_mul(q[1], q[1]), r_mul(q[2], q[2])), r_mul(q[3], q[3]))), 0.0001f);
  const float e = expf(q[0]), sn = sinf(nv);
  o[0] = r_mul(e, cosf(nv));
  for (int c = 1; c < 4; ++c) o[c] = r_mul(e, r_mul(q[c] / nv, sn));
}
SELDQ_HD void qexp_bwd(const float q[4], const float g[4], float gq[4]) {
  const float m = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), nv = m + 0.0001f;
  const float e = expf(q[0]), sn = sinf(nv), cs = cosf(nv), s = sn / nv, ds = (cs * nv - sn) / (nv * nv);
  const float gv = g[1] * q[1] + g[2] * q[2] + g[3] * q[3];
  gq[0] = e * (g[0] * cs + s * gv);
  const float radial = (ds * gv - g[0] * sn) / m;          // m = 0: 0 / 0, as torch.sqrt's backward gives the reference
  for (int c = 1; c < 4; ++c) gq[c] = e * (s * g[c] + radial * q[c]);
}

// one quaternion of the (outer, 4, m) tensors; op: seldq_qpointwise_op_t (seldq.h); b may be null for the one-operand ops
SELDQ_HD void qpointwise_element(int op, const float* a, const float* b, float* out, long long idx, long long m) {
  const long long o = idx / m, x = idx - o * m, base = o * 4 * m + x;
  float qa[4], qb[4] = {0.f, 0.f, 0.f, 0.f}, res[4];
  for (int c = 0; c < 4; ++c) qa[c] = a[base + c * m];
  if (b != nullptr)
    for (int c = 0; c < 4; ++c) qb[c] = b[base + c * m];
  switch (op) {
    case SELDQ_QOP_HAMILTON: hamilton(qa, qb, res); break;
    case SELDQ_QOP_HAMILTON_CONJ_B: qb[1] = -qb[1]; qb[2] = -qb[2]; qb[3] = -qb[3]; hamilton(qa, qb, res); break;
    case SELDQ_QOP_HAMILTON_CONJ_A: qa[1] = -qa[1]; qa[2] = -qa[2]; qa[3] = -qa[3]; hamilton(qa, qb, res); break;
    case SELDQ_QOP_NORMALIZE: qnormalize(qa, res); break;
    case SELDQ_QOP_NORMALIZE_BWD: qnormalize_bwd(qa, qb, res); break;
    case SELDQ_QOP_EXP: qexp(qa, res); break;
    default: qexp_bwd(qa, qb, res); break;
  }
  for (int c = 0; c < 4; ++c) out[base + c * m] = res[c];
}

}  // namespace rot
}  // namespace seldq

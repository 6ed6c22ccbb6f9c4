// Device side of the quaternion operators the SELD models never reach, so that the drop-in covers the
// whole of quaternion_ops.py / dual_quaternion_ops.py:
//   * rotation_weight_kernel / rotation_weight_bwd_kernel: the real (nc d0, nc d1, taps) weight of the three rotation
//     variants (quaternion_ops.py:174-232, :235-295, :330-388) and the gradient of its compact tensors.  The
//     contraction itself runs on the convolution kernels with the real algebra (SELDQ_ALG_REAL).
//   * quaternion_pointwise_kernel: hamilton_product (quaternion_ops.py:467-507, dual_quaternion_ops.py:374-414) with
//     optional conjugation of either factor (its two gradients are Hamilton products with a conjugate),
//     q_normalize and quaternion_exp (dual_quaternion_ops.py:206-246) and their gradients.
// All of it is bandwidth-bound element work (a few hundred KB per layer): one thread per compact element /
// quaternion, grid-stride, stores of a warp contiguous in the output.  Arithmetic: rotation.cuh.
#include <cuda_runtime.h>

#include "launch.h"
#include "rotation.cuh"
#include "rotation.h"

namespace seldq {
namespace rot {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) rotation_weight_kernel(RotGeom g, const float* __restrict__ r,
                                                                  const float* __restrict__ i, const float* __restrict__ j,
                                                                  const float* __restrict__ k, float* __restrict__ out) {
  const long long total = g.d0 * g.d1 * g.taps;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads)
    rot_fwd_element(g, idx, r, i, j, k, out);
}

__global__ void __launch_bounds__(kThreads) rotation_weight_bwd_kernel(RotGeom g, const float* __restrict__ r,
                                                                      const float* __restrict__ i,
                                                                      const float* __restrict__ j,
                                                                      const float* __restrict__ k,
                                                                      const float* __restrict__ gout, float* __restrict__ gr,
                                                                      float* __restrict__ gi, float* __restrict__ gj,
                                                                      float* __restrict__ gk) {
  const long long total = g.d0 * g.d1 * g.taps;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads)
    rot_bwd_element(g, idx, r, i, j, k, gout, gr, gi, gj, gk);
}

// op: seldq_qpointwise_op_t (seldq.h).  One thread per quaternion of the (outer, 4, m) tensors.
__global__ void __launch_bounds__(kThreads) quaternion_pointwise_kernel(int op, const float* __restrict__ a,
                                                                       const float* __restrict__ b, float* __restrict__ out,
                                                                       long long outer, long long m) {
  const long long total = outer * m;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
    qpointwise_element(op, a, b, out, idx, m);
  }
}

static int grid_for(long long total) {
  long long blocks = (total + kThreads - 1) / kThreads;
  const long long cap = 148LL * 8;                     // 148 SMs x 8 resident blocks of 256 threads
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static int make_geom(long long d0, long long d1, long long taps, int quaternion_format, int transpose_out, RotGeom* g) {
  if (d0 < 1 || d1 < 1 || taps < 1) return fail(SELDQ_ERR_INVALID, "rotation weight: non-positive size");
  if ((double)d0 * (double)d1 * (double)taps * 16.0 > 9.0e15) return fail(SELDQ_ERR_INVALID, "rotation weight: too large");
  g->d0 = d0; g->d1 = d1; g->taps = taps;
  g->nc = quaternion_format ? 4 : 3;
  g->transpose_out = transpose_out ? 1 : 0;
  return SELDQ_OK;
}

}  // namespace rot

int launch_rotation_weight(const float* const* w, long long d0, long long d1, long long taps, int quaternion_format,
                           int transpose_out, float* out, cudaStream_t st) {
  rot::RotGeom g;
  if (int rc = rot::make_geom(d0, d1, taps, quaternion_format, transpose_out, &g)) return rc;
  rot::rotation_weight_kernel<<<rot::grid_for(d0 * d1 * taps), rot::kThreads, 0, st>>>(g, w[0], w[1], w[2], w[3], out);
  return check_launch("rotation_weight_kernel");
}

int launch_rotation_weight_bwd(const float* const* w, const float* g_out, long long d0, long long d1, long long taps,
                               int quaternion_format, int transpose_out, float* const* gw, cudaStream_t st) {
  rot::RotGeom g;
  if (int rc = rot::make_geom(d0, d1, taps, quaternion_format, transpose_out, &g)) return rc;
  rot::rotation_weight_bwd_kernel<<<rot::grid_for(d0 * d1 * taps), rot::kThreads, 0, st>>>(g, w[0], w[1], w[2], w[3], g_out,
                                                                                         gw[0], gw[1], gw[2], gw[3]);
  return check_launch("rotation_weight_bwd_kernel");
}

int launch_quaternion_pointwise(int op, const float* a, const float* b, float* out, long long outer, long long m,
                                cudaStream_t st) {
  if (outer < 1 || m < 1) return fail(SELDQ_ERR_INVALID, "quaternion point-wise operator: non-positive size");
  if (op < SELDQ_QOP_HAMILTON || op > SELDQ_QOP_EXP_BWD) return fail(SELDQ_ERR_INVALID, "quaternion point-wise operator %d", op);
  const bool two = op != SELDQ_QOP_NORMALIZE && op != SELDQ_QOP_EXP;
  if (two && b == nullptr) return fail(SELDQ_ERR_INVALID, "quaternion point-wise operator %d needs two operands", op);
  rot::quaternion_pointwise_kernel<<<rot::grid_for(outer * m), rot::kThreads, 0, st>>>(op, a, two ? b : nullptr, out, outer, m);
  return check_launch("quaternion_pointwise_kernel");
}

}  // namespace seldq

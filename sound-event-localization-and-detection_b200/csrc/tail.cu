// Glue of the TC_Block tail (model.py:210-231):
//     sum of skips -> ReLU -> MaxPool1d(p) -> conv1 -> attention -> ReLU -> MaxPool1d(p) -> conv2 -> tanh -> MaxPool1d(p)
// Each activation + pooling pair is ONE bandwidth-bound kernel per direction instead of two PyTorch kernels each
// (activation, pooling; pooling backward, activation backward).  Both activations are monotone, so
// max_k act(x[p t + k]) = act(max_k x[p t + k]) -- one activation per pooled element instead of p -- and the first
// maximum wins, as in nn.MaxPool1d.  x: fp32 (rows = N * C, T); y: fp32 (rows, T / p) (floor mode: a tail of T % p
// samples is dropped and gets a zero gradient).
#include <cuda_runtime.h>

#include "launch.h"
#include "pdl.cuh"
#include "tail.h"

namespace seldq {
namespace tail {

// act: 0 = ReLU, 1 = tanh
template <int P>
__global__ void __launch_bounds__(256) act_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows,
                                                          int T, int To, int pool, int act) {
  const long long total = rows * To;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / To;
    const int t = (int)(i - r * To);
    const float* src = x + r * T + (long long)t * pool;
    float m;
    if (P == 2) {                                          // the shipped configurations: one 8-byte load
      const float2 v = __ldg(reinterpret_cast<const float2*>(src));
      m = fmaxf(v.x, v.y);
    } else {
      m = __ldg(src);
      for (int k = 1; k < pool; ++k) m = fmaxf(m, __ldg(src + k));
    }
    y[i] = act == 0 ? fmaxf(m, 0.f) : tanhf(m);
  }
}

template <int P>
__global__ void __launch_bounds__(256) act_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          const float* __restrict__ gy, float* __restrict__ gx,
                                                          long long rows, int T, int To, int pool, int act) {
  // one thread per POOLED element writes the p input gradients of its window (+ the dropped tail of the row)
  const long long total = rows * To;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / To;
    const int t = (int)(i - r * To);
    const float* src = x + r * T + (long long)t * pool;
    float* dst = gx + r * T + (long long)t * pool;
    const float yo = __ldg(y + i), g = __ldg(gy + i);
    // d act / d (its argument) at the maximum: ReLU: 1 where the maximum is positive; tanh: 1 - y^2
    const float d = act == 0 ? (yo > 0.f ? g : 0.f) : g * (1.f - yo * yo);
    if (P == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src));
      const bool first = v.x >= v.y;                       // ties: the first index, as nn.MaxPool1d
      *reinterpret_cast<float2*>(dst) = make_float2(first ? d : 0.f, first ? 0.f : d);
    } else {
      int arg = 0;
      float m = __ldg(src);
      for (int k = 1; k < pool; ++k) {
        const float v = __ldg(src + k);
        if (v > m) { m = v; arg = k; }
      }
      for (int k = 0; k < pool; ++k) dst[k] = k == arg ? d : 0.f;
    }
    if (t == To - 1)
      for (int k = To * pool; k < T; ++k) gx[r * T + k] = 0.f;
  }
}

}  // namespace tail

static int grid_for(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int launch_act_pool_fwd(const float* x, float* y, long long rows, int T, int pool, int act, cudaStream_t st) {
  if (pool < 1 || T / pool < 1) return fail(SELDQ_ERR_INVALID, "act_pool: pool %d does not fit a row of %d", pool, T);
  const int To = T / pool;
  const bool p2 = pool == 2 && (T % 2) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
  if (p2) tail::act_pool_fwd_kernel<2><<<grid_for(rows * To), 256, 0, st>>>(x, y, rows, T, To, pool, act);
  else tail::act_pool_fwd_kernel<0><<<grid_for(rows * To), 256, 0, st>>>(x, y, rows, T, To, pool, act);
  return check_launch("act_pool_fwd_kernel");
}

int launch_act_pool_bwd(const float* x, const float* y, const float* gy, float* gx, long long rows, int T, int pool, int act,
                        cudaStream_t st) {
  if (pool < 1 || T / pool < 1) return fail(SELDQ_ERR_INVALID, "act_pool: pool %d does not fit a row of %d", pool, T);
  const int To = T / pool;
  const bool p2 = pool == 2 && (T % 2) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gx)) & 7) == 0;
  if (p2) tail::act_pool_bwd_kernel<2><<<grid_for(rows * To), 256, 0, st>>>(x, y, gy, gx, rows, T, To, pool, act);
  else tail::act_pool_bwd_kernel<0><<<grid_for(rows * To), 256, 0, st>>>(x, y, gy, gx, rows, T, To, pool, act);
  return check_launch("act_pool_bwd_kernel");
}

}  // namespace seldq

// ---- optimiser step ------------------------------------------------------------------------------------------------
// Adam as torch.optim.Adam(lr, betas, eps; no weight decay, no amsgrad) applies it (train.py:502-504), over the
// trainer's ONE flat parameter / gradient bucket: a single bandwidth-bound pass (7 x 4 bytes per parameter) instead
// of the multi-tensor kernel's chunked walk.  The step counter lives on the device (float, as PyTorch's capturable
// Adam keeps it), so the launch can be captured in a CUDA graph; the kernel increments it.
namespace seldq {
namespace tail {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, long long n, double lr, double b1d, double b2d,
                                                  float eps, float* __restrict__ step) {
  // scalars as torch/optim/adam.py forms them: Python doubles (1 - beta, beta^t, lr / (1 - beta1^t), sqrt(1 - beta2^t)),
  // rounded to float only where they meet the tensors
  const double t = (double)*step + 1.0;
  const float b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
  const float step_size = (float)(lr / (1.0 - pow(b1d, t))), bc2s = (float)sqrt(1.0 - pow(b2d, t));
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = ma[k] + omb1 * (ga[k] - ma[k]);                       // exp_avg.lerp_(grad, 1 - beta1)
      va[k] = __fmaf_rn(omb2 * ga[k], ga[k], b2 * va[k]);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      pa[k] -= step_size * (ma[k] / (sqrtf(va[k]) / bc2s + eps));
    }
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) {
      m[i] = m[i] + omb1 * (g[i] - m[i]);
      v[i] = __fmaf_rn(omb2 * g[i], g[i], b2 * v[i]);
      p[i] -= step_size * (m[i] / (sqrtf(v[i]) / bc2s + eps));
    }
}
// the counter moves in its own one-thread launch BEHIND the update (every block of adam_kernel reads the old value)
__global__ void adam_count_kernel(float* step) { *step += 1.f; }

}  // namespace tail

int launch_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2, double eps, float* step,
                int advance, cudaStream_t st) {
  if (n < 1) return SELDQ_OK;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(SELDQ_ERR_INVALID, "seldq_adam_step: buffers must be 16-byte aligned");
  tail::adam_kernel<<<grid_for(n >> 2), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, (float)eps, step);
  int rc = check_launch("adam_kernel");
  if (rc) return rc;
  if (!advance) return SELDQ_OK;
  tail::adam_count_kernel<<<1, 1, 0, st>>>(step);
  return check_launch("adam_count_kernel");
}

}  // namespace seldq

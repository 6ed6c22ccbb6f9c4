// extern "C" surface of libseldq.so (see include/seldq.h for the contract of every entry point).
#include <cuda_runtime.h>

#include "conv_simt.cuh"
#include "conv_umma.h"
#include "geom.h"
#include "launch.h"
#include "stft.cuh"

using namespace seldq;

namespace {

inline int pitch8(int w) { return (w + 7) & ~7; }

size_t bf16_mirror_bytes(int n, int c, int h, int w) { return (size_t)n * c * h * pitch8(w) * 2; }

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int cuda_ready() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SELDQ_ERR_CUDA, "no CUDA device: libseldq has no CPU fallback");
  }
  return SELDQ_OK;
}

// fp32 (rows, w) -> bf16 (rows, pitch8(w)); the pad columns are zeroed
int cast_mirror(const float* src, void* dst, long long rows, int w, cudaStream_t st) {
  if (w % 8 == 0) return launch_cast_bf16(src, dst, (size_t)rows * w, st);
  return launch_cast_bf16_rows(src, dst, rows, w, pitch8(w), st);
}

}  // namespace

extern "C" int seldq_abi_version(void) { return SELDQ_ABI_VERSION; }
extern "C" const char* seldq_last_error(void) { return error_buffer(); }
extern "C" int seldq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int seldq_bf16_pitch(int32_t w) { return pitch8(w); }

extern "C" int seldq_conv_out_shape(const seldq_conv_desc_t* d, int32_t* out_h, int32_t* out_w) {
  BlockTable t;
  int oh, ow;
  const int rc = validate_conv(d, &t, &oh, &ow);
  if (rc) return rc;
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return SELDQ_OK;
}

extern "C" size_t seldq_conv_workspace_bytes(const seldq_conv_desc_t* d, int32_t pass) {
  BlockTable t;
  int oh, ow;
  if (validate_conv(d, &t, &oh, &ow)) return 0;
  if (d->precision != SELDQ_PREC_BF16) return 0;
  const size_t xb = align256(bf16_mirror_bytes(d->batch, d->cin, d->in_h, d->in_w));
  const size_t yb = align256(bf16_mirror_bytes(d->batch, d->cout, oh, ow));
  switch (pass) {
    case SELDQ_PASS_FWD: return xb;
    case SELDQ_PASS_DGRAD: return yb;
    case SELDQ_PASS_WGRAD: return xb + yb;
    default: return 0;
  }
}

extern "C" int seldq_cast_bf16(const float* src, void* dst_bf16, size_t n, void* stream) {
  if (!src || !dst_bf16) return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_cast_bf16(src, dst_bf16, n, (cudaStream_t)stream);
}

extern "C" int seldq_cast_bf16_mirror(const float* src, void* dst_bf16, int64_t rows, int32_t w, void* stream) {
  if (!src || !dst_bf16 || rows <= 0 || w <= 0) return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16_mirror: bad arguments");
  int rc = cuda_ready();
  if (rc) return rc;
  return cast_mirror(src, dst_bf16, rows, w, (cudaStream_t)stream);
}

extern "C" int seldq_conv_fwd(const seldq_conv_desc_t* d, const float* x, const void* x_bf16,
                              const float* const* host_w, const float* bias, float* y, void* y_bf16, void* workspace,
                              size_t workspace_bytes, void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_FWD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  if (!host_w || !y || (!x && !(bf16 && x_bf16))) return fail(SELDQ_ERR_INVALID, "seldq_conv_fwd: null pointer");
  for (int i = 0; i < g.tab.nw; ++i)
    if (!host_w[i]) return fail(SELDQ_ERR_INVALID, "seldq_conv_fwd: weight %d is null", i);
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->precision == SELDQ_PREC_FP32) {
    simt::ConvParams p{};
    p.g = g; p.in = x; p.out = y; p.bias = bias;
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
    return launch_conv_simt(p, st);
  }
  const void* xb = x_bf16;
  if (!xb) {
    if (workspace_bytes < seldq_conv_workspace_bytes(d, SELDQ_PASS_FWD) || !workspace)
      return fail(SELDQ_ERR_WORKSPACE, "seldq_conv_fwd: workspace too small (%zu < %zu)", workspace_bytes,
                  seldq_conv_workspace_bytes(d, SELDQ_PASS_FWD));
    if ((rc = cast_mirror(x, workspace, (long long)d->batch * d->cin * d->in_h, d->in_w, st))) return rc;
    xb = workspace;
  }
  return launch_umma_fprop(g, xb, pitch8(d->in_w), host_w, bias, y, y_bf16, pitch8(g.OW), st);
}

extern "C" int seldq_conv_dgrad(const seldq_conv_desc_t* d, const float* gy, const void* gy_bf16,
                                const float* const* host_w, float* gx, void* workspace, size_t workspace_bytes,
                                void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_DGRAD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  if (!host_w || !gx || (!gy && !(bf16 && gy_bf16))) return fail(SELDQ_ERR_INVALID, "seldq_conv_dgrad: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->precision == SELDQ_PREC_FP32) {
    simt::ConvParams p{};
    p.g = g; p.in = gy; p.out = gx; p.bias = nullptr;
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
    return launch_conv_simt(p, st);
  }
  const void* gb = gy_bf16;
  if (!gb) {
    if (workspace_bytes < seldq_conv_workspace_bytes(d, SELDQ_PASS_DGRAD) || !workspace)
      return fail(SELDQ_ERR_WORKSPACE, "seldq_conv_dgrad: workspace too small");
    if ((rc = cast_mirror(gy, workspace, (long long)d->batch * d->cout * g.IH, g.IW, st))) return rc;
    gb = workspace;
  }
  return launch_umma_fprop(g, gb, pitch8(g.IW), host_w, nullptr, gx, nullptr, 0, st);
}

extern "C" int seldq_conv_wgrad(const seldq_conv_desc_t* d, const float* x, const void* x_bf16, const float* gy,
                                const void* gy_bf16, float* const* host_gw, float* gbias, void* workspace,
                                size_t workspace_bytes, void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_WGRAD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  if (!host_gw || (!x && !(bf16 && x_bf16)) || (!gy && (gbias || !(bf16 && gy_bf16))))
    return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wbytes = (size_t)g.Oc * g.Ic * g.KH * g.KW * sizeof(float);
  for (int i = 0; i < g.tab.nw; ++i) {
    if (!host_gw[i]) return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad: gradient %d is null", i);
    const cudaError_t e = cudaMemsetAsync(host_gw[i], 0, wbytes, st);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  if (gbias && (rc = launch_bias_grad(gy, gbias, g.P, g.N, g.OH, g.OW, g.out_sN, g.out_sC, g.out_sH, g.out_sW, st)))
    return rc;
  if (d->precision == SELDQ_PREC_FP32) {
    simt::WgradParams p{};
    p.g = g; p.x = x; p.gy = gy;
    for (int i = 0; i < g.tab.nw; ++i) p.gw[i] = host_gw[i];
    return launch_wgrad_simt(p, st);
  }
  const size_t xbytes = align256(bf16_mirror_bytes(d->batch, d->cin, d->in_h, d->in_w));
  const void* xb = x_bf16;
  const void* gb = gy_bf16;
  if (!xb || !gb) {
    if (workspace_bytes < seldq_conv_workspace_bytes(d, SELDQ_PASS_WGRAD) || !workspace)
      return fail(SELDQ_ERR_WORKSPACE, "seldq_conv_wgrad: workspace too small");
    if (!xb) {
      if ((rc = cast_mirror(x, workspace, (long long)d->batch * d->cin * d->in_h, d->in_w, st))) return rc;
      xb = workspace;
    }
    if (!gb) {
      void* dst = (char*)workspace + xbytes;
      if ((rc = cast_mirror(gy, dst, (long long)d->batch * d->cout * g.OH, g.OW, st))) return rc;
      gb = dst;
    }
  }
  return launch_umma_wgrad(g, xb, pitch8(d->in_w), gb, pitch8(g.OW), host_gw, st);
}

// ---- linear: 0.18 GFLOP per call in the reference configs -> always the fp32 FFMA kernels -----------
extern "C" size_t seldq_linear_workspace_bytes(const seldq_linear_desc_t*, int32_t) { return 0; }

extern "C" int seldq_linear_fwd(const seldq_linear_desc_t* d, const float* x, const float* const* host_w,
                                const float* bias, float* y, void*, size_t, void* stream) {
  simt::ConvParams p{};
  int rc = make_linear_geom(d, SELDQ_PASS_FWD, &p.g);
  if (rc) return rc;
  if (!x || !host_w || !y) return fail(SELDQ_ERR_INVALID, "seldq_linear_fwd: null pointer");
  if ((rc = cuda_ready())) return rc;
  p.in = x; p.out = y; p.bias = bias;
  for (int i = 0; i < p.g.tab.nw; ++i) p.w[i] = host_w[i];
  return launch_conv_simt(p, (cudaStream_t)stream);
}

extern "C" int seldq_linear_dgrad(const seldq_linear_desc_t* d, const float* gy, const float* const* host_w, float* gx,
                                  void*, size_t, void* stream) {
  simt::ConvParams p{};
  int rc = make_linear_geom(d, SELDQ_PASS_DGRAD, &p.g);
  if (rc) return rc;
  if (!gy || !host_w || !gx) return fail(SELDQ_ERR_INVALID, "seldq_linear_dgrad: null pointer");
  if ((rc = cuda_ready())) return rc;
  p.in = gy; p.out = gx; p.bias = nullptr;
  for (int i = 0; i < p.g.tab.nw; ++i) p.w[i] = host_w[i];
  return launch_conv_simt(p, (cudaStream_t)stream);
}

extern "C" int seldq_linear_wgrad(const seldq_linear_desc_t* d, const float* x, const float* gy, float* const* host_gw,
                                  float* gbias, void*, size_t, void* stream) {
  simt::WgradParams p{};
  int rc = make_linear_geom(d, SELDQ_PASS_WGRAD, &p.g);
  if (rc) return rc;
  if (!x || !gy || !host_gw) return fail(SELDQ_ERR_INVALID, "seldq_linear_wgrad: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const ConvGeom& g = p.g;
  for (int i = 0; i < g.tab.nw; ++i) {
    if (!host_gw[i]) return fail(SELDQ_ERR_INVALID, "seldq_linear_wgrad: gradient %d is null", i);
    const cudaError_t e = cudaMemsetAsync(host_gw[i], 0, (size_t)g.Oc * g.Ic * sizeof(float), st);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    p.gw[i] = host_gw[i];
  }
  p.x = x; p.gy = gy;
  if (gbias && (rc = launch_bias_grad(gy, gbias, g.P, 1, 1, g.OW, 0, g.out_sC, 0, g.out_sW, st))) return rc;
  return launch_wgrad_simt(p, st);
}

// ---- STFT ------------------------------------------------------------------------------------------
extern "C" int seldq_stft_shape(int64_t n_samples, int32_t nperseg, int32_t noverlap, int32_t cut_dc, int32_t cut_last,
                                int32_t* n_bins, int32_t* n_frames) {
  int b, f;
  const int rc = stft_shape(n_samples, nperseg, noverlap, cut_dc, cut_last, &b, &f);
  if (rc) return rc;
  if (n_bins) *n_bins = b;
  if (n_frames) *n_frames = f;
  return SELDQ_OK;
}

extern "C" int seldq_stft_magphase(const float* x, int32_t n_batch, int32_t n_ch, int64_t n_samples, int32_t nperseg,
                                   int32_t noverlap, int32_t cut_dc, int32_t output_phase, int32_t cut_last, float* out,
                                   void* stream) {
  if (!x || !out || n_batch <= 0 || n_ch <= 0) return fail(SELDQ_ERR_INVALID, "seldq_stft_magphase: bad arguments");
  stft::Params p{};
  int rc = stft_shape(n_samples, nperseg, noverlap, cut_dc, cut_last, &p.n_bins, &p.n_frames);
  if (rc) return rc;
  if ((rc = cuda_ready())) return rc;
  p.x = x; p.out = out; p.n_samples = n_samples; p.n_ch = n_ch; p.hop = nperseg - noverlap;
  p.bin0 = cut_dc ? 1 : 0; p.output_phase = output_phase ? 1 : 0;
  return launch_stft(p, n_batch * n_ch, (cudaStream_t)stream);
}

// extern "C" surface of libseldq.so (see include/seldq.h for the contract of every entry point).
#include <cuda_runtime.h>

#include "conv_simt.cuh"
#include "conv_umma.h"
#include "geom.h"
#include "launch.h"
#include "stft.cuh"

using namespace seldq;

namespace {

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t mirror_bytes(long long rows, int w, int nshifts) {
  return align256((size_t)nshifts * (size_t)rows * mirror_pitch(w) * 2);
}

int cuda_ready() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SELDQ_ERR_CUDA, "no CUDA device: libseldq has no CPU fallback");
  }
  return SELDQ_OK;
}

// forward-orientation geometry (tap offsets are defined on it) + the shift list of one operand
int shifts_for(const seldq_conv_desc_t* d, int which, MirrorSet* m) {
  ConvGeom g;
  const int rc = make_conv_geom(d, SELDQ_PASS_FWD, &g);
  if (rc) return rc;
  mirror_shifts(g, which, m->shifts, &m->nshifts);
  return SELDQ_OK;
}

// uses the caller's mirror set if given, otherwise builds it in the workspace at *offset
int obtain_mirror(const seldq_conv_desc_t* d, int which, const float* src, const void* given, long long rows, int w,
                  void* workspace, size_t workspace_bytes, size_t* offset, cudaStream_t st, MirrorSet* m) {
  int rc = shifts_for(d, which, m);
  if (rc) return rc;
  if (given) {
    m->data = given;
    return SELDQ_OK;
  }
  const size_t need = mirror_bytes(rows, w, m->nshifts);
  if (!workspace || *offset + need > workspace_bytes)
    return fail(SELDQ_ERR_WORKSPACE, "workspace too small for the bf16 mirror set (%zu + %zu > %zu)", *offset, need,
                workspace_bytes);
  void* dst = (char*)workspace + *offset;
  if ((rc = launch_cast_bf16_mirror(src, dst, rows, w, mirror_pitch(w), m->shifts, m->nshifts, st))) return rc;
  m->data = dst;
  *offset += need;
  return SELDQ_OK;
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

extern "C" int seldq_conv_out_shape(const seldq_conv_desc_t* d, int32_t* out_h, int32_t* out_w) {
  BlockTable t;
  int oh, ow;
  const int rc = validate_conv(d, &t, &oh, &ow);
  if (rc) return rc;
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return SELDQ_OK;
}

// ---- bf16 mirror sets ---------------------------------------------------------------------------------
extern "C" int seldq_bf16_pitch(int32_t w) { return mirror_pitch(w); }

extern "C" int seldq_conv_mirror_shifts(const seldq_conv_desc_t* d, int32_t which, int32_t* shifts8, int32_t* nshifts) {
  if (!shifts8 || !nshifts || (which != 0 && which != 1)) return fail(SELDQ_ERR_INVALID, "bad mirror-shift query");
  MirrorSet m;
  const int rc = shifts_for(d, which, &m);
  if (rc) return rc;
  for (int i = 0; i < m.nshifts; ++i) shifts8[i] = m.shifts[i];
  *nshifts = m.nshifts;
  return SELDQ_OK;
}

extern "C" size_t seldq_bf16_mirror_bytes(int64_t rows, int32_t w, int32_t nshifts) {
  if (rows <= 0 || w <= 0 || nshifts <= 0) return 0;
  return mirror_bytes(rows, w, nshifts);
}

extern "C" int seldq_cast_bf16_mirror(const float* src, void* dst_bf16, int64_t rows, int32_t w, const int32_t* shifts,
                                      int32_t nshifts, void* stream) {
  if (!src || !dst_bf16 || rows <= 0 || w <= 0 || !shifts || nshifts < 1 || nshifts > 8 || shifts[0] != 0)
    return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16_mirror: bad arguments (shifts[0] must be 0, at most 8 shifts)");
  for (int i = 0; i < nshifts; ++i)
    if (shifts[i] < 0 || shifts[i] > 7) return fail(SELDQ_ERR_INVALID, "mirror shifts must be in [0, 8)");
  int rc = cuda_ready();
  if (rc) return rc;
  int s[8];
  for (int i = 0; i < nshifts; ++i) s[i] = shifts[i];
  return launch_cast_bf16_mirror(src, dst_bf16, rows, w, mirror_pitch(w), s, nshifts, (cudaStream_t)stream);
}

extern "C" int seldq_cast_bf16(const float* src, void* dst_bf16, size_t n, void* stream) {
  if (!src || !dst_bf16) return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_cast_bf16(src, dst_bf16, n, (cudaStream_t)stream);
}

extern "C" size_t seldq_conv_workspace_bytes(const seldq_conv_desc_t* d, int32_t pass) {
  BlockTable t;
  int oh, ow;
  if (validate_conv(d, &t, &oh, &ow)) return 0;
  if (d->precision != SELDQ_PREC_BF16) return 0;
  MirrorSet mx, mg;
  if (shifts_for(d, 0, &mx) || shifts_for(d, 1, &mg)) return 0;
  const size_t xb = mirror_bytes((long long)d->batch * d->cin * d->in_h, d->in_w, mx.nshifts);
  const size_t gb = mirror_bytes((long long)d->batch * d->cout * oh, ow, mg.nshifts);
  switch (pass) {
    case SELDQ_PASS_FWD: return xb;
    case SELDQ_PASS_DGRAD: return gb;
    case SELDQ_PASS_WGRAD: return xb + gb;
    default: return 0;
  }
}

// ---- convolution ------------------------------------------------------------------------------------
extern "C" int seldq_conv_fwd(const seldq_conv_desc_t* d, const float* x, const void* x_bf16,
                              const float* const* host_w, const float* bias, float* y, void* workspace,
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
  if (!bf16) {
    simt::ConvParams p{};
    p.g = g; p.in = x; p.out = y; p.bias = bias;
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
    return launch_conv_simt(p, st);
  }
  MirrorSet mx;
  size_t off = 0;
  if ((rc = obtain_mirror(d, 0, x, x_bf16, (long long)d->batch * d->cin * d->in_h, d->in_w, workspace, workspace_bytes,
                          &off, st, &mx)))
    return rc;
  return launch_umma_fprop(g, mx, host_w, bias, y, st);
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
  if (!bf16) {
    simt::ConvParams p{};
    p.g = g; p.in = gy; p.out = gx; p.bias = nullptr;
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
    return launch_conv_simt(p, st);
  }
  MirrorSet mg;
  size_t off = 0;
  if ((rc = obtain_mirror(d, 1, gy, gy_bf16, (long long)d->batch * d->cout * g.IH, g.IW, workspace, workspace_bytes,
                          &off, st, &mg)))
    return rc;
  return launch_umma_fprop(g, mg, host_w, nullptr, gx, st);
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
  if (!bf16) {
    simt::WgradParams p{};
    p.g = g; p.x = x; p.gy = gy;
    for (int i = 0; i < g.tab.nw; ++i) p.gw[i] = host_gw[i];
    return launch_wgrad_simt(p, st);
  }
  MirrorSet mx, mg;
  size_t off = 0;
  if ((rc = obtain_mirror(d, 0, x, x_bf16, (long long)d->batch * d->cin * d->in_h, d->in_w, workspace, workspace_bytes,
                          &off, st, &mx)))
    return rc;
  if ((rc = obtain_mirror(d, 1, gy, gy_bf16, (long long)d->batch * d->cout * g.OH, g.OW, workspace, workspace_bytes,
                          &off, st, &mg)))
    return rc;
  return launch_umma_wgrad(g, mx, mg, host_gw, st);
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

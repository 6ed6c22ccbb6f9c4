// extern "C" surface of libseldq.so (see include/seldq.h for the contract of every entry point).
#include <cuda_runtime.h>

#include "attention.h"
#include "conv_cl.h"
#include "conv_simt.cuh"
#include "conv_umma.h"
#include "epilogue.h"
#include "eval.h"
#include "rotation.h"
#include "geom.h"
#include "launch.h"
#include "stft.cuh"
#include "tail.h"
#include "tcn_glue.h"
#include "wgrad_first.h"

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

// bf16 operands of a convolution (forward-orientation geometry): which = 0 -> x, 1 -> gy
struct OperandInfo {
  cl::OperandLayout lay;
  int n, c, h, w;
  size_t cl_bytes, t16_bytes;
};
OperandInfo operand_info(const ConvGeom& fwd, int which) {
  OperandInfo o;
  o.lay = which == 0 ? x_operand_layout(fwd) : gy_operand_layout(fwd);
  o.n = fwd.N;
  o.c = which == 0 ? fwd.R : fwd.P;
  o.h = which == 0 ? fwd.IH : fwd.OH;
  o.w = which == 0 ? fwd.IW : fwd.OW;
  o.cl_bytes = align256((size_t)o.n * o.h * o.w * o.lay.Cp * 2);
  o.t16_bytes = align256((size_t)o.n * o.c * o.h * nchw16_pitch(o.w) * 2);
  return o;
}

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t size, used;
  void* take(size_t bytes) {
    bytes = align256(bytes);
    if (!base || used + bytes > size) return nullptr;
    void* p = base + used;
    used += bytes;
    return p;
  }
};
int workspace_short(size_t need, const Workspace& ws) {
  return fail(SELDQ_ERR_WORKSPACE, "workspace too small: %zu more bytes needed after %zu of %zu (see seldq_conv_workspace_bytes)",
              need, ws.used, ws.size);
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

// ---- bf16 operand staging ---------------------------------------------------------------------------
extern "C" int seldq_conv_operand_info(const seldq_conv_desc_t* d, int32_t which, int32_t* padded_channels,
                                       int32_t* dense, size_t* cl_bytes, size_t* t16_bytes) {
  if (which != 0 && which != 1) return fail(SELDQ_ERR_INVALID, "operand selector must be 0 (x) or 1 (gy)");
  ConvGeom g;
  const int rc = make_conv_geom(d, SELDQ_PASS_FWD, &g);
  if (rc) return rc;
  const OperandInfo o = operand_info(g, which);
  if (padded_channels) *padded_channels = o.lay.Cp;
  if (dense) *dense = (o.lay.nc != g.tab.nc || g.tab.nc == 1) ? 1 : 0;
  if (cl_bytes) *cl_bytes = o.cl_bytes;
  if (t16_bytes) *t16_bytes = o.t16_bytes;
  return SELDQ_OK;
}

extern "C" int seldq_stage_operand(const seldq_conv_desc_t* d, int32_t which, const float* src, void* dst_cl,
                                   void* dst_t16, void* stream) {
  if (which != 0 && which != 1) return fail(SELDQ_ERR_INVALID, "operand selector must be 0 (x) or 1 (gy)");
  if (!src || (!dst_cl && !dst_t16)) return fail(SELDQ_ERR_INVALID, "seldq_stage_operand: null pointer");
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_FWD, &g);
  if (rc) return rc;
  if ((rc = cuda_ready())) return rc;
  const OperandInfo o = operand_info(g, which);
  return launch_stage_operand(src, dst_cl, dst_t16, o.lay, o.n, o.c, o.h, o.w, (cudaStream_t)stream);
}

// many layers in one launch: the caller builds a table on the host (one entry per (layer, pass)), keeps a copy of
// it in device memory for as long as the weight / packed buffers stay where they are, and replays it every step
extern "C" size_t seldq_conv_pack_table_entry_bytes(void) { return pack_table_entry_bytes(); }

extern "C" int seldq_conv_pack_table_fill(const seldq_conv_desc_t* d, int32_t pass, const float* const* host_w,
                                          void* packed, void* host_entry, int32_t* items) {
  ConvGeom g;
  if (pass != SELDQ_PASS_FWD && pass != SELDQ_PASS_DGRAD) return fail(SELDQ_ERR_INVALID, "weights are packed for FWD and DGRAD");
  if (!host_w || !packed || !host_entry || !items) return fail(SELDQ_ERR_INVALID, "seldq_conv_pack_table_fill: null pointer");
  int rc = make_conv_geom(d, pass, &g);
  if (rc) return rc;
  int n = 0;
  if ((rc = fill_pack_table_entry(g, host_w, packed, host_entry, &n))) return rc;
  *items = n;
  return SELDQ_OK;
}

extern "C" int seldq_conv_pack_table_run(const void* dev_table, int32_t count, int32_t max_items, void* stream) {
  if (!dev_table || count < 0) return fail(SELDQ_ERR_INVALID, "seldq_conv_pack_table_run: bad arguments");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_pack_weights_multi(dev_table, count, max_items, (cudaStream_t)stream);
}

extern "C" size_t seldq_conv_packed_bytes(const seldq_conv_desc_t* d, int32_t pass) {
  ConvGeom g;
  if ((pass != SELDQ_PASS_FWD && pass != SELDQ_PASS_DGRAD) || make_conv_geom(d, pass, &g)) return 0;
  return align256(packed_weight_bytes(g));
}

extern "C" int seldq_conv_pack_weights(const seldq_conv_desc_t* d, int32_t pass, const float* const* host_w,
                                       void* packed, void* stream) {
  if (pass != SELDQ_PASS_FWD && pass != SELDQ_PASS_DGRAD)
    return fail(SELDQ_ERR_INVALID, "weights are packed for the forward or the dgrad pass");
  ConvGeom g;
  int rc = make_conv_geom(d, pass, &g);
  if (rc) return rc;
  if (!host_w) return fail(SELDQ_ERR_INVALID, "seldq_conv_pack_weights: null pointer");
  for (int i = 0; i < g.tab.nw; ++i)
    if (!host_w[i]) return fail(SELDQ_ERR_INVALID, "seldq_conv_pack_weights: weight %d is null", i);
  if (packed_weight_bytes(g) == 0) return SELDQ_OK;
  if (!packed) return fail(SELDQ_ERR_INVALID, "seldq_conv_pack_weights: null destination");
  if ((rc = cuda_ready())) return rc;
  return launch_pack_weights(g, host_w, packed, (cudaStream_t)stream);
}

extern "C" int seldq_cast_bf16(const float* src, void* dst_bf16, size_t n, void* stream) {
  if (!src || !dst_bf16) return fail(SELDQ_ERR_INVALID, "seldq_cast_bf16: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_cast_bf16(src, dst_bf16, n, (cudaStream_t)stream);
}

extern "C" size_t seldq_conv_workspace_bytes(const seldq_conv_desc_t* d, int32_t pass) {
  ConvGeom g;
  if (make_conv_geom(d, SELDQ_PASS_FWD, &g)) return 0;
  if (d->precision != SELDQ_PREC_BF16) return 0;
  const OperandInfo ox = operand_info(g, 0), og = operand_info(g, 1);
  switch (pass) {
    case SELDQ_PASS_FWD: return ox.cl_bytes + seldq_conv_packed_bytes(d, SELDQ_PASS_FWD);
    case SELDQ_PASS_DGRAD: return og.cl_bytes + seldq_conv_packed_bytes(d, SELDQ_PASS_DGRAD);
    case SELDQ_PASS_WGRAD: {
      size_t xb = ox.cl_bytes;
      if (ox.lay.nc != g.tab.nc || g.tab.nc == 1) {   // narrow first layer: shifted mirror set of x instead
        MirrorSet m;
        mirror_shifts(g, 0, m.shifts, &m.nshifts);
        xb = mirror_bytes((long long)g.N * g.R * g.IH, g.IW, m.nshifts);
      }
      return xb + og.t16_bytes;
    }
    default: return 0;
  }
}

// ---- convolution ------------------------------------------------------------------------------------
static int conv_fwd_impl(const seldq_conv_desc_t* d, const float* x, const void* x_cl, const float* const* host_w,
                         const void* packed_w, const float* bias, float* y, void* y_f16, void* workspace,
                         size_t workspace_bytes, void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_FWD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  if (!host_w || (!y && !y_f16) || (!x && !(bf16 && x_cl))) return fail(SELDQ_ERR_INVALID, "seldq_conv_fwd: null pointer");
  if (y_f16 && !bf16) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 output exists on the SELDQ_PREC_BF16 path only");
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
  Workspace ws{(char*)workspace, workspace_bytes, 0};
  if (!x_cl) {
    const OperandInfo o = operand_info(g, 0);
    void* buf = ws.take(o.cl_bytes);
    if (!buf) return workspace_short(o.cl_bytes, ws);
    if ((rc = launch_stage_operand(x, buf, nullptr, o.lay, o.n, o.c, o.h, o.w, st))) return rc;
    x_cl = buf;
  }
  const size_t pb = packed_weight_bytes(g);
  if (!packed_w && pb) {
    void* buf = ws.take(pb);
    if (!buf) return workspace_short(pb, ws);
    if ((rc = launch_pack_weights(g, host_w, buf, st))) return rc;
    packed_w = buf;
  }
  return launch_cl_fprop(g, x_cl, host_w, packed_w, bias, y, y_f16, st);
}

extern "C" int seldq_conv_fwd(const seldq_conv_desc_t* d, const float* x, const void* x_cl,
                              const float* const* host_w, const void* packed_w, const float* bias, float* y,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return conv_fwd_impl(d, x, x_cl, host_w, packed_w, bias, y, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int seldq_conv_fwd_bf16(const seldq_conv_desc_t* d, const float* x, const void* x_cl,
                                   const float* const* host_w, const void* packed_w, const float* bias, void* y_f16,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (!y_f16) return fail(SELDQ_ERR_INVALID, "seldq_conv_fwd_bf16: null output");
  return conv_fwd_impl(d, x, x_cl, host_w, packed_w, bias, nullptr, y_f16, workspace, workspace_bytes, stream);
}

extern "C" int seldq_conv_dgrad(const seldq_conv_desc_t* d, const float* gy, const void* gy_cl,
                                const float* const* host_w, const void* packed_w, float* gx, void* workspace,
                                size_t workspace_bytes, void* stream) {
  ConvGeom g, fwd;
  int rc = make_conv_geom(d, SELDQ_PASS_DGRAD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  if (!host_w || !gx || (!gy && !(bf16 && gy_cl))) return fail(SELDQ_ERR_INVALID, "seldq_conv_dgrad: null pointer");
  for (int i = 0; i < g.tab.nw; ++i)
    if (!host_w[i]) return fail(SELDQ_ERR_INVALID, "seldq_conv_dgrad: weight %d is null", i);
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!bf16) {
    simt::ConvParams p{};
    p.g = g; p.in = gy; p.out = gx; p.bias = nullptr;
    for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
    return launch_conv_simt(p, st);
  }
  if ((rc = make_conv_geom(d, SELDQ_PASS_FWD, &fwd))) return rc;
  Workspace ws{(char*)workspace, workspace_bytes, 0};
  if (!gy_cl) {
    const OperandInfo o = operand_info(fwd, 1);
    void* buf = ws.take(o.cl_bytes);
    if (!buf) return workspace_short(o.cl_bytes, ws);
    if ((rc = launch_stage_operand(gy, buf, nullptr, o.lay, o.n, o.c, o.h, o.w, st))) return rc;
    gy_cl = buf;
  }
  const size_t pb = packed_weight_bytes(g);
  if (!packed_w && pb) {
    void* buf = ws.take(pb);
    if (!buf) return workspace_short(pb, ws);
    if ((rc = launch_pack_weights(g, host_w, buf, st))) return rc;
    packed_w = buf;
  }
  return launch_cl_fprop(g, gy_cl, host_w, packed_w, nullptr, gx, nullptr, st);
}

// debug (tools/fprop_trace.py): 64 x uint64 device buffer that CTA 0 of every later convolution launch stamps with
// globaltimer values (conv_cl.cu trace_stamp); NULL switches the trace off.  Not part of the reference surface.
extern "C" int seldq_debug_fprop_trace(void* dev_buf) {
  cl::set_fprop_trace(dev_buf);
  return SELDQ_OK;
}

static cl::FpropEpilogue to_epilogue(const seldq_conv_epilogue_t* e) {
  cl::FpropEpilogue o;
  if (e) { o.mode = e->mode; o.addend = e->addend; o.stats = e->stats; }
  return o;
}

static int pair_geom(const seldq_conv_desc_t* d, int32_t pass, ConvGeom* g) {
  if (!d) return fail(SELDQ_ERR_INVALID, "null descriptor");
  if (pass != SELDQ_PASS_FWD && pass != SELDQ_PASS_DGRAD) return fail(SELDQ_ERR_INVALID, "pass must be FWD or DGRAD");
  if (d->precision != SELDQ_PREC_BF16) return fail(SELDQ_ERR_UNSUPPORTED, "pre-staged convolution entry points exist on the SELDQ_PREC_BF16 path only");
  return make_conv_geom(d, pass, g);
}

extern "C" int seldq_conv_epi(const seldq_conv_desc_t* d, int32_t pass, const void* in_cl, const void* packed_w,
                              float* out, const seldq_conv_epilogue_t* epi, void* stream) {
  ConvGeom g;
  int rc = pair_geom(d, pass, &g);
  if (rc) return rc;
  if (!in_cl || !packed_w || !out) return fail(SELDQ_ERR_INVALID, "seldq_conv_epi: null pointer");
  if (cl::is_dense(g)) return fail(SELDQ_ERR_UNSUPPORTED, "seldq_conv_epi serves layers with >= 8 channels per component");
  if ((rc = cuda_ready())) return rc;
  const cl::FpropEpilogue e = to_epilogue(epi);
  return launch_cl_fprop(g, in_cl, nullptr, packed_w, nullptr, out, nullptr, (cudaStream_t)stream, &e);
}

extern "C" int seldq_conv_pair_supported(const seldq_conv_desc_t* d, int32_t pass) {
  ConvGeom g;
  if (pair_geom(d, pass, &g) || cl::is_dense(g)) return 0;
  cl::FpropParams p;
  size_t smem = 0;
  return plan_cl_fprop_pair(g, &p, &smem) == SELDQ_OK ? 1 : 0;
}

extern "C" int seldq_conv_pair(const seldq_conv_desc_t* d, int32_t pass, const void* in_cl_a, const void* in_cl_b,
                               const void* packed_a, const void* packed_b, float* out_a, float* out_b,
                               const seldq_conv_epilogue_t* epi_a, const seldq_conv_epilogue_t* epi_b, void* stream) {
  ConvGeom g;
  int rc = pair_geom(d, pass, &g);
  if (rc) return rc;
  if ((rc = cuda_ready())) return rc;
  const void* in[2] = {in_cl_a, in_cl_b};
  const void* pk[2] = {packed_a, packed_b};
  float* out[2] = {out_a, out_b};
  const cl::FpropEpilogue e[2] = {to_epilogue(epi_a), to_epilogue(epi_b)};
  return launch_cl_fprop_pair(g, in, pk, out, e, (cudaStream_t)stream);
}

extern "C" int seldq_conv_wgrad(const seldq_conv_desc_t* d, const float* x, const void* x_cl, const float* gy,
                                const void* gy_t16, float* const* host_gw, float* gbias, int32_t accumulate,
                                void* workspace, size_t workspace_bytes, void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_WGRAD, &g);
  if (rc) return rc;
  const bool bf16 = d->precision == SELDQ_PREC_BF16;
  const OperandInfo ox = operand_info(g, 0), og = operand_info(g, 1);
  const bool narrow = bf16 && (ox.lay.nc != g.tab.nc || g.tab.nc == 1);   // first layer: mirror-set kernel, needs fp32 x
  if (!host_gw || (!x && (!bf16 || narrow || !x_cl)) || (!gy && (gbias || !(bf16 && gy_t16))))
    return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad: null pointer%s",
                narrow && !x ? " (this layer's weight gradient reads the float32 input)" : "");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wbytes = (size_t)g.Oc * g.Ic * g.KH * g.KW * sizeof(float);
  for (int i = 0; i < g.tab.nw; ++i) {
    if (!host_gw[i]) return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad: gradient %d is null", i);
    if (accumulate) continue;
    const cudaError_t e = cudaMemsetAsync(host_gw[i], 0, wbytes, st);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  if (gbias && (rc = launch_bias_grad(gy, gbias, g.P, g.N, g.OH, g.OW, g.out_sN, g.out_sC, g.out_sH, g.out_sW,
                                      accumulate, st)))
    return rc;
  if (!bf16) {
    simt::WgradParams p{};
    p.g = g; p.x = x; p.gy = gy;
    for (int i = 0; i < g.tab.nw; ++i) p.gw[i] = host_gw[i];
    return launch_wgrad_simt(p, st);
  }
  Workspace ws{(char*)workspace, workspace_bytes, 0};
  if (!gy_t16) {
    void* buf = ws.take(og.t16_bytes);
    if (!buf) return workspace_short(og.t16_bytes, ws);
    if ((rc = launch_stage_operand(gy, nullptr, buf, og.lay, og.n, og.c, og.h, og.w, st))) return rc;
    gy_t16 = buf;
  }
  if (narrow) {
    MirrorSet mx, mg;
    mirror_shifts(g, 0, mx.shifts, &mx.nshifts);
    mx.pitch = 0;
    const size_t need = mirror_bytes((long long)g.N * g.R * g.IH, g.IW, mx.nshifts);
    void* buf = ws.take(need);
    if (!buf) return workspace_short(need, ws);
    if ((rc = launch_cast_bf16_mirror(x, buf, (long long)g.N * g.R * g.IH, g.IW, mirror_pitch(g.IW), mx.shifts,
                                      mx.nshifts, st)))
      return rc;
    mx.data = buf;
    mg.data = gy_t16; mg.nshifts = 1; mg.shifts[0] = 0; mg.pitch = nchw16_pitch(g.OW);
    return launch_umma_wgrad(g, mx, mg, host_gw, st);
  }
  if (!x_cl) {
    void* buf = ws.take(ox.cl_bytes);
    if (!buf) return workspace_short(ox.cl_bytes, ws);
    if ((rc = launch_stage_operand(x, buf, nullptr, ox.lay, ox.n, ox.c, ox.h, ox.w, st))) return rc;
    x_cl = buf;
  }
  return launch_cl_wgrad(g, x_cl, gy_t16, host_gw, st);
}

// ---- glue between the convolutions (E1) ------------------------------------------------------------------
extern "C" int seldq_bn_stats(const void* src, int32_t is_f16, int32_t n, int32_t c, int64_t plane, double* sums,
                              void* stream) {
  if (!src || !sums || n <= 0 || c <= 0 || plane <= 0) return fail(SELDQ_ERR_INVALID, "seldq_bn_stats: bad arguments");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_bn_stats(src, is_f16, n, c, plane, sums, (cudaStream_t)stream);
}

extern "C" int seldq_bn_finalize(const double* sums, const float* gamma, const float* beta, int32_t c, double count,
                                 float eps, float momentum, float* running_mean, float* running_var, float* coef,
                                 void* stream) {
  if (!sums || !coef || c <= 0 || count <= 0) return fail(SELDQ_ERR_INVALID, "seldq_bn_finalize: bad arguments");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_bn_finalize(sums, gamma, beta, c, count, eps, momentum, running_mean, running_var, coef,
                            (cudaStream_t)stream);
}

static int tail_params(const seldq_cnn_tail_desc_t* t, const cl::OperandLayout* lay, epi::TailParams* p) {
  if (!t || t->n <= 0 || t->c <= 0 || t->h <= 0 || t->w <= 0 || t->pool <= 0 || t->pool > 8 || t->h / t->pool <= 0 ||
      t->drop_p < 0.f || t->drop_p >= 1.f)
    return fail(SELDQ_ERR_INVALID, "bad CNN tail descriptor (pool must be in [1, 8], 0 <= drop_p < 1)");
  memset(p, 0, sizeof(*p));
  p->N = t->n; p->C = t->c; p->H = t->h; p->W = t->w; p->pool = t->pool;
  p->drop_p = t->drop_p; p->salt = t->salt;
  if (lay) { p->nc = lay->nc; p->cc = lay->cc; p->cpad = lay->cpad; p->Cp = lay->Cp; }
  else { p->nc = 1; p->cc = t->c; p->cpad = (t->c + 63) / 64 * 64; p->Cp = p->cpad; }
  return SELDQ_OK;
}

extern "C" int seldq_cnn_tail_fwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* consumer, const void* y_f16,
                                  const float* coef, const int64_t* seed, void* z_cl, float* z_f32, uint8_t* idx,
                                  void* ymax, void* stream) {
  epi::TailParams p;
  cl::OperandLayout lay;
  int rc;
  if (z_cl) {
    if (!consumer) return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_fwd: z_cl needs the consuming convolution's descriptor");
    ConvGeom g;
    if ((rc = make_conv_geom(consumer, SELDQ_PASS_FWD, &g))) return rc;
    lay = x_operand_layout(g);
    if (g.R != t->c || g.N != t->n || g.IH != t->h / t->pool || g.IW != t->w)
      return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_fwd: pooled output does not match the consumer's input");
  }
  if ((rc = tail_params(t, z_cl ? &lay : nullptr, &p))) return rc;
  if (!y_f16 || !coef || !idx || (!z_cl && !z_f32)) return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_fwd: null pointer");
  if (t->drop_p > 0.f && !seed) return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_fwd: dropout needs a seed pointer");
  if ((rc = cuda_ready())) return rc;
  p.y = reinterpret_cast<const __half*>(y_f16); p.coef = coef;
  p.z_cl = reinterpret_cast<__nv_bfloat16*>(z_cl); p.z32 = z_f32; p.idx = idx;
  p.ymax = reinterpret_cast<__half*>(ymax);
  p.seed_ptr = reinterpret_cast<const long long*>(seed);
  return launch_cnn_tail_fwd(p, (cudaStream_t)stream);
}

extern "C" int seldq_cnn_tail_bwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* producer, const void* y_f16,
                                  const float* coef, const uint8_t* idx, const void* ymax, const float* gz, double* dsums,
                                  void* d_t16, void* d_cl, void* stream) {
  epi::TailParams p;
  cl::OperandLayout lay;
  int rc;
  if (d_cl) {
    if (!producer) return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_bwd: d_cl needs the producing convolution's descriptor");
    ConvGeom g;
    if ((rc = make_conv_geom(producer, SELDQ_PASS_FWD, &g))) return rc;
    lay = gy_operand_layout(g);
    if (g.P != t->c || g.N != t->n || g.OH != t->h || g.OW != t->w)
      return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_bwd: descriptor does not match the producer's output");
  }
  if ((rc = tail_params(t, d_cl ? &lay : nullptr, &p))) return rc;
  if (!y_f16 || !coef || !idx || !gz || !dsums || (!d_t16 && !d_cl))
    return fail(SELDQ_ERR_INVALID, "seldq_cnn_tail_bwd: null pointer");
  if ((rc = cuda_ready())) return rc;
  p.y = reinterpret_cast<const __half*>(y_f16); p.coef = coef;
  p.idx = const_cast<uint8_t*>(idx); p.gz = gz;
  p.ymax = reinterpret_cast<__half*>(const_cast<void*>(ymax));
  p.d_t16 = reinterpret_cast<__nv_bfloat16*>(d_t16); p.d_cl = reinterpret_cast<__nv_bfloat16*>(d_cl);
  return launch_cnn_tail_bwd(p, dsums, (cudaStream_t)stream);
}

// two convolutions of equal geometry reading the same x (filter / gate, skip / residual of a residual block):
// both weight gradients in one launch (twice the K extent per CTA for the same fixed cost)
extern "C" int seldq_conv_wgrad_pair(const seldq_conv_desc_t* d, const void* x_cl, const void* gy_t16_a,
                                     const void* gy_t16_b, float* const* host_gw_a, float* const* host_gw_b,
                                     int32_t accumulate, void* stream) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_WGRAD, &g);
  if (rc) return rc;
  if (d->precision != SELDQ_PREC_BF16) return fail(SELDQ_ERR_UNSUPPORTED, "seldq_conv_wgrad_pair exists on the SELDQ_PREC_BF16 path only");
  const OperandInfo ox = operand_info(g, 0);
  if (ox.lay.nc != g.tab.nc || g.tab.nc == 1)
    return fail(SELDQ_ERR_UNSUPPORTED, "seldq_conv_wgrad_pair: narrow / real layers take seldq_conv_wgrad");
  if (!x_cl || !gy_t16_a || !gy_t16_b || !host_gw_a || !host_gw_b) return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad_pair: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wbytes = (size_t)g.Oc * g.Ic * g.KH * g.KW * sizeof(float);
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < g.tab.nw; ++i) {
      float* gw = (k ? host_gw_b : host_gw_a)[i];
      if (!gw) return fail(SELDQ_ERR_INVALID, "seldq_conv_wgrad_pair: gradient %d is null", i);
      if (accumulate) continue;
      const cudaError_t e = cudaMemsetAsync(gw, 0, wbytes, st);
      if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    }
  return launch_cl_wgrad(g, x_cl, gy_t16_a, host_gw_a, st, gy_t16_b, host_gw_b);
}

// ---- first CNN block, backward: BatchNorm-backward apply fused into the weight-gradient kernel ----------------
static int first_bwd_geom(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* conv, ConvGeom* g) {
  int rc = make_conv_geom(conv, SELDQ_PASS_WGRAD, g);
  if (rc) return rc;
  if (!t || g->P != t->c || g->N != t->n || g->OH != t->h || g->OW != t->w)
    return fail(SELDQ_ERR_INVALID, "seldq_cnn_first_bwd: tail descriptor does not match the convolution's output");
  return SELDQ_OK;
}

extern "C" int seldq_cnn_first_bwd_supported(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* conv) {
  ConvGeom g;
  if (!conv || conv->precision != SELDQ_PREC_BF16 || first_bwd_geom(t, conv, &g)) return 0;
  return (first_layer_bwd_supported(g) && t->pool >= 1 && t->pool <= 8 && t->h / t->pool > 0 && t->h % t->pool == 0) ? 1 : 0;
}

extern "C" size_t seldq_cnn_first_bwd_workspace_bytes(const seldq_conv_desc_t* conv) {
  ConvGeom g;
  if (make_conv_geom(conv, SELDQ_PASS_WGRAD, &g)) return 0;
  int shifts[8], ns = 0;
  mirror_shifts(g, 0, shifts, &ns);
  return mirror_bytes((long long)g.N * g.R * g.IH, g.IW, ns) + 256;
}

extern "C" int seldq_cnn_first_bwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* conv, const float* x,
                                   const void* y_f16, const float* coef, const uint8_t* idx, const void* ymax,
                                   const float* gz, double* dsums, float* const* host_gw, int32_t accumulate,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  ConvGeom g;
  int rc = first_bwd_geom(t, conv, &g);
  if (rc) return rc;
  if (!seldq_cnn_first_bwd_supported(t, conv))
    return fail(SELDQ_ERR_UNSUPPORTED, "seldq_cnn_first_bwd: geometry outside the fused kernel (see seldq_cnn_first_bwd_supported)");
  if (!x || !y_f16 || !coef || !idx || !gz || !dsums || !host_gw)
    return fail(SELDQ_ERR_INVALID, "seldq_cnn_first_bwd: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  epi::TailParams p;
  if ((rc = tail_params(t, nullptr, &p))) return rc;
  p.y = reinterpret_cast<const __half*>(y_f16); p.coef = coef;
  p.idx = const_cast<uint8_t*>(idx); p.gz = gz;
  p.ymax = reinterpret_cast<__half*>(const_cast<void*>(ymax));
  for (int i = 0; i < g.tab.nw; ++i)
    if (!host_gw[i]) return fail(SELDQ_ERR_INVALID, "seldq_cnn_first_bwd: gradient %d is null", i);
  Workspace ws{(char*)workspace, workspace_bytes, 0};
  MirrorSet mx;
  mirror_shifts(g, 0, mx.shifts, &mx.nshifts);
  mx.pitch = 0;
  const size_t need = mirror_bytes((long long)g.N * g.R * g.IH, g.IW, mx.nshifts);
  void* buf = ws.take(need);
  if (!buf) return workspace_short(need, ws);
  // the bf16 mirror copies of x (36 us for the 8-channel input) do not depend on the BatchNorm reductions (22 us):
  // they run on a forked stream next to them
  ForkJoin fj(st, 1);
  rc = launch_cast_bf16_mirror(x, buf, (long long)g.N * g.R * g.IH, g.IW, mirror_pitch(g.IW), mx.shifts, mx.nshifts, fj.side());
  if (rc == SELDQ_OK) rc = launch_cnn_tail_bwd_reduce(p, dsums, st);
  const float2* dmean = reinterpret_cast<const float2*>(dsums + 2 * (size_t)t->c);
  const size_t wbytes = (size_t)g.Oc * g.Ic * g.KH * g.KW * sizeof(float);
  for (int i = 0; i < g.tab.nw && rc == SELDQ_OK && !accumulate; ++i) {
    const cudaError_t e = cudaMemsetAsync(host_gw[i], 0, wbytes, st);
    if (e != cudaSuccess) rc = fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  if (!fj.join() && rc == SELDQ_OK) rc = fail(SELDQ_ERR_CUDA, "seldq_cnn_first_bwd: join: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc) return rc;
  mx.data = buf;
  return launch_first_layer_bwd(g, mx, p, dmean, host_gw, st);
}

extern "C" int seldq_tcn_glue(int32_t op, const seldq_tcn_glue_t* a, const seldq_conv_desc_t* layout_of, int32_t which,
                              void* stream) {
  if (!a || a->n <= 0 || a->c <= 0 || a->t <= 0 || a->count <= 0 || a->drop_p < 0.f || a->drop_p >= 1.f)
    return fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: bad arguments");
  tcn::GlueParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->n; p.C = a->c; p.T = a->t; p.C2 = a->c2;
  p.count = a->count; p.inv_count = 1.0 / a->count; p.eps = a->eps; p.momentum = a->momentum;
  p.drop_p = a->drop_p; p.salt = a->salt; p.seed_ptr = reinterpret_cast<const long long*>(a->seed);
  if (a->drop_p > 0.f && !a->seed) return fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: dropout needs a seed pointer");
  for (int i = 0; i < 2; ++i) {
    p.bn[i].sums = a->bn[i].sums; p.bn[i].gamma = a->bn[i].gamma; p.bn[i].beta = a->bn[i].beta;
    p.bn[i].running_mean = a->bn[i].running_mean; p.bn[i].running_var = a->bn[i].running_var;
    p.out_cl[i] = reinterpret_cast<__nv_bfloat16*>(a->out_cl[i]);
    p.out_t16[i] = reinterpret_cast<__nv_bfloat16*>(a->out_t16[i]);
    p.stats_out[i] = a->stats_out[i];
  }
  for (int i = 0; i < 5; ++i) p.in[i] = a->in[i];
  p.out32 = a->out32; p.dsums = a->dsums; p.accum = a->accum;
  p.out32b = a->out32b; p.sync = a->sync;
  p.cc = a->c; p.cpad = (a->c + 63) / 64 * 64; p.Cp = p.cpad;
  p.pitch = nchw16_pitch(a->t);
  int rc;
  if (layout_of) {
    if (which != 0 && which != 1) return fail(SELDQ_ERR_INVALID, "operand selector must be 0 (x) or 1 (gy)");
    ConvGeom g;
    if ((rc = make_conv_geom(layout_of, SELDQ_PASS_FWD, &g))) return rc;
    const OperandInfo o = operand_info(g, which);
    if (o.c != a->c || o.n != a->n || o.h != 1 || o.w != a->t)
      return fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: tensor (%d, %d, %d) is not operand %d of the given convolution",
                  a->n, a->c, a->t, which);
    p.cc = o.lay.cc; p.cpad = o.lay.cpad; p.Cp = o.lay.Cp;
    if (o.lay.nc == 1) { p.cc = a->c; p.cpad = o.lay.Cp; }
  } else if (a->out_cl[0] || a->out_cl[1]) {
    return fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: channels-last outputs need the consuming convolution's descriptor");
  }
  // required pointers per step
  auto need = [&](bool ok) { return ok ? SELDQ_OK : fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: null pointer for op %d", op); };
  switch (op) {
    case SELDQ_TCN_PREACT_FWD: rc = need(a->in[0] && a->out32 && a->out_cl[0] && a->bn[0].sums); break;
    case SELDQ_TCN_ROW_STATS:
      rc = need(a->flag >= 1 && a->flag <= 2 && a->in[0] && a->stats_out[0] && (a->flag < 2 || (a->in[1] && a->stats_out[1])));
      break;
    case SELDQ_TCN_GATE_FWD: rc = need(a->in[0] && a->in[1] && a->out_cl[0] && a->bn[0].sums && a->bn[1].sums); break;
    case SELDQ_TCN_RESIDUAL_FWD:
      rc = need((a->in[1] == nullptr || (a->in[0] && a->out32)) && (a->in[2] == nullptr || (a->accum && a->c2 > 0)) &&
                (a->in[1] || a->in[2]));
      break;
    case SELDQ_TCN_GATE_BWD_REDUCE:
      rc = need(a->in[0] && a->in[1] && a->in[2] && a->dsums && a->bn[0].sums && a->bn[1].sums);
      break;
    case SELDQ_TCN_GATE_BWD_APPLY:
      rc = need(a->in[0] && a->in[1] && a->in[2] && a->dsums && a->bn[0].sums && a->bn[1].sums && a->out_cl[0] &&
                a->out_cl[1] && a->out_t16[0] && a->out_t16[1]);
      break;
    case SELDQ_TCN_PREACT_BWD_REDUCE:
      rc = need(a->in[1] && a->in[2] && a->in[3] && a->in[4] && a->dsums && a->bn[0].sums);
      break;
    case SELDQ_TCN_PREACT_BWD_APPLY:
      rc = need(a->in[1] && a->in[2] && a->in[3] && a->in[4] && a->dsums && a->bn[0].sums && a->out32);
      break;
    case SELDQ_TCN_GATE_BWD:
      rc = need(a->in[0] && a->in[1] && a->in[2] && a->dsums && a->bn[0].sums && a->bn[1].sums && a->out_cl[0] &&
                a->out_cl[1] && a->out_t16[0] && a->out_t16[1] && a->sync);
      break;
    case SELDQ_TCN_PREACT_BWD:
      rc = need(a->in[1] && a->in[2] && a->in[3] && a->in[4] && a->dsums && a->bn[0].sums && a->out32 && a->sync);
      break;
    case SELDQ_TCN_GATE_FWD_STATS:
      rc = need(a->in[0] && a->in[1] && a->out_cl[0] && a->stats_out[0] && a->stats_out[1] && a->sync);
      break;
    case SELDQ_TCN_RESIDUAL_PREACT_FWD:
      rc = need(a->in[0] && a->in[1] && a->in[2] && a->out32 && a->out32b && a->out_cl[0] && a->accum && a->dsums &&
                a->c2 == a->c && a->sync);
      break;
    default: return fail(SELDQ_ERR_INVALID, "seldq_tcn_glue: unknown op %d", op);
  }
  if (rc) return rc;
  if ((rc = cuda_ready())) return rc;
  return launch_tcn_glue(op, p, a->flag, (cudaStream_t)stream);
}

extern "C" int seldq_tcn_glue_fused_supported(int32_t n, int32_t c, int32_t t) {
  if (cuda_ready()) return 0;
  return tcn_glue_fused_supported(n, c, t);
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
                                  float* gbias, int32_t accumulate, void*, size_t, void* stream) {
  simt::WgradParams p{};
  int rc = make_linear_geom(d, SELDQ_PASS_WGRAD, &p.g);
  if (rc) return rc;
  if (!x || !gy || !host_gw) return fail(SELDQ_ERR_INVALID, "seldq_linear_wgrad: null pointer");
  if ((rc = cuda_ready())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const ConvGeom& g = p.g;
  for (int i = 0; i < g.tab.nw; ++i) {
    if (!host_gw[i]) return fail(SELDQ_ERR_INVALID, "seldq_linear_wgrad: gradient %d is null", i);
    p.gw[i] = host_gw[i];
    if (accumulate) continue;
    const cudaError_t e = cudaMemsetAsync(host_gw[i], 0, (size_t)g.Oc * g.Ic * sizeof(float), st);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  p.x = x; p.gy = gy;
  if (gbias && (rc = launch_bias_grad(gy, gbias, g.P, 1, 1, g.OW, 0, g.out_sC, 0, g.out_sW, accumulate, st))) return rc;
  return launch_wgrad_simt(p, st);
}

// ---- STFT ------------------------------------------------------------------------------------------
// ---- TC_Block tail ---------------------------------------------------------------------------------------------
extern "C" int seldq_act_pool1d_fwd(const float* x, int64_t rows, int32_t t, int32_t pool, int32_t act, float* y,
                                    void* stream) {
  if (!x || !y || rows < 1 || t < 1 || (act != SELDQ_ACT_RELU && act != SELDQ_ACT_TANH))
    return fail(SELDQ_ERR_INVALID, "seldq_act_pool1d_fwd: bad argument");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_act_pool_fwd(x, y, rows, t, pool, act, (cudaStream_t)stream);
}
extern "C" int seldq_act_pool1d_bwd(const float* x, const float* y, const float* gy, int64_t rows, int32_t t,
                                    int32_t pool, int32_t act, float* gx, void* stream) {
  if (!x || !y || !gy || !gx || rows < 1 || t < 1 || (act != SELDQ_ACT_RELU && act != SELDQ_ACT_TANH))
    return fail(SELDQ_ERR_INVALID, "seldq_act_pool1d_bwd: bad argument");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_act_pool_bwd(x, y, gy, gx, rows, t, pool, act, (cudaStream_t)stream);
}

extern "C" int seldq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                               double b1, double b2, double eps, float* step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step) return fail(SELDQ_ERR_INVALID, "seldq_adam_step: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, b1, b2, eps, step, 1, (cudaStream_t)stream);
}

extern "C" int seldq_adam_step_part(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                                    double b1, double b2, double eps, float* step, int32_t advance, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step) return fail(SELDQ_ERR_INVALID, "seldq_adam_step_part: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, b1, b2, eps, step, advance ? 1 : 0, (cudaStream_t)stream);
}

// ---- rotation variants and quaternion point-wise operators ------------------------------------------
extern "C" int seldq_rotation_weight(const float* const* host_w, int64_t d0, int64_t d1, int64_t taps,
                                     int32_t quaternion_format, int32_t transpose_out, float* out, void* stream) {
  if (!host_w || !host_w[0] || !host_w[1] || !host_w[2] || !host_w[3] || !out)
    return fail(SELDQ_ERR_INVALID, "seldq_rotation_weight: null pointer");
  return launch_rotation_weight(host_w, d0, d1, taps, quaternion_format, transpose_out, out, (cudaStream_t)stream);
}

extern "C" int seldq_rotation_weight_bwd(const float* const* host_w, const float* g_out, int64_t d0, int64_t d1, int64_t taps,
                                         int32_t quaternion_format, int32_t transpose_out, float* const* host_gw,
                                         void* stream) {
  if (!host_w || !host_w[0] || !host_w[1] || !host_w[2] || !host_w[3] || !g_out || !host_gw || !host_gw[0] || !host_gw[1] ||
      !host_gw[2] || !host_gw[3])
    return fail(SELDQ_ERR_INVALID, "seldq_rotation_weight_bwd: null pointer");
  return launch_rotation_weight_bwd(host_w, g_out, d0, d1, taps, quaternion_format, transpose_out, host_gw,
                                    (cudaStream_t)stream);
}

extern "C" int seldq_quaternion_pointwise(int32_t op, const float* a, const float* b, float* out, int64_t outer, int64_t m,
                                          void* stream) {
  if (!a || !out) return fail(SELDQ_ERR_INVALID, "seldq_quaternion_pointwise: null pointer");
  return launch_quaternion_pointwise(op, a, b, out, outer, m, (cudaStream_t)stream);
}

// ---- evaluation path -------------------------------------------------------------------------------------------
extern "C" int seldq_seld_events(const float* sed, const float* doa, int32_t clips, int32_t frames, int32_t classes,
                                 int32_t overlaps, float max_loc, float* rows, int32_t* counts, void* stream) {
  if (!sed || !doa || !rows || !counts) return fail(SELDQ_ERR_INVALID, "seldq_seld_events: null pointer");
  int rc = cuda_ready();
  if (rc) return rc;
  return launch_seld_events(sed, doa, clips, frames, classes, overlaps, max_loc, rows, counts, (cudaStream_t)stream);
}

// ---- attention ------------------------------------------------------------------------------------------------
namespace {
struct AttnBufs { size_t rm, tr; char *q_rm, *k_rm, *v_rm, *q_tr, *k_tr, *v_tr; };
AttnBufs attn_saved_layout(const seldq_attention_desc_t* d, void* saved) {
  AttnBufs b;
  b.rm = (attention_rm_bytes(d->batch, d->heads, d->seq) + 255) & ~(size_t)255;
  b.tr = (attention_tr_bytes(d->batch, d->heads, d->seq, d->head_dim) + 255) & ~(size_t)255;
  char* p = (char*)saved;
  b.q_rm = p; b.k_rm = p + b.rm; b.v_rm = p + 2 * b.rm;
  b.q_tr = p + 3 * b.rm; b.k_tr = b.q_tr + b.tr; b.v_tr = b.q_tr + 2 * b.tr;
  return b;
}
}  // namespace

extern "C" int seldq_attention_supported(const seldq_attention_desc_t* d) {
  return d && attention_supported(d->batch, d->heads, d->seq, d->head_dim) ? 1 : 0;
}
extern "C" size_t seldq_attention_saved_bytes(const seldq_attention_desc_t* d) {
  if (!seldq_attention_supported(d)) return 0;
  const AttnBufs b = attn_saved_layout(d, nullptr);
  return 3 * b.rm + 3 * b.tr;
}
extern "C" size_t seldq_attention_bwd_workspace_bytes(const seldq_attention_desc_t* d) {
  if (!seldq_attention_supported(d)) return 0;
  const AttnBufs b = attn_saved_layout(d, nullptr);
  return b.rm + b.tr + (((size_t)d->batch * d->heads * d->seq * sizeof(float) + 255) & ~(size_t)255);
}
extern "C" int seldq_attention_fwd(const seldq_attention_desc_t* d, const float* q, const float* k, const float* v,
                                   float* out, float* lse, void* saved, void* stream) {
  if (!seldq_attention_supported(d)) return fail(SELDQ_ERR_UNSUPPORTED, "attention: head_dim must be 16 / 32 / 48 and the sequence a multiple of 8");
  if (!q || !k || !v || !out || !lse || !saved) return fail(SELDQ_ERR_INVALID, "seldq_attention_fwd: null pointer");
  if (reinterpret_cast<uintptr_t>(saved) & 255) return fail(SELDQ_ERR_INVALID, "seldq_attention_fwd: `saved` must be 256-byte aligned");
  int rc = cuda_ready();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const AttnBufs b = attn_saved_layout(d, saved);
  // softmax(q k^T / sqrt(d)) through exp2: q carries log2(e) / sqrt(d)
  const float qs = 1.4426950408889634f / sqrtf((float)d->head_dim);
  // the three operand-staging launches are independent: k and v go to forked streams next to q
  ForkJoin fk(st, 2), fv(st, 3);
  rc = launch_attention_stage(q, nullptr, b.q_rm, b.q_tr, nullptr, d->batch, d->heads, d->seq, d->head_dim, 0, qs, st);
  if (rc == SELDQ_OK) rc = launch_attention_stage(k, nullptr, b.k_rm, b.k_tr, nullptr, d->batch, d->heads, d->seq, d->head_dim, 0, 1.f, fk.side());
  if (rc == SELDQ_OK) rc = launch_attention_stage(v, nullptr, b.v_rm, b.v_tr, nullptr, d->batch, d->heads, d->seq, d->head_dim, 0, 1.f, fv.side());
  const bool jk = fk.join(), jv = fv.join();
  if ((!jk || !jv) && rc == SELDQ_OK) rc = fail(SELDQ_ERR_CUDA, "seldq_attention_fwd: join: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc) return rc;
  return launch_attention_fwd(d->batch, d->heads, d->seq, d->head_dim, b.q_rm, b.k_rm, b.v_tr, out, lse, st);
}
extern "C" int seldq_attention_bwd(const seldq_attention_desc_t* d, const void* saved, const float* out, const float* lse,
                                   const float* d_out, float* dq, float* dk, float* dv, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!seldq_attention_supported(d)) return fail(SELDQ_ERR_UNSUPPORTED, "attention: unsupported shape");
  if (!saved || !out || !lse || !d_out || !dq || !dk || !dv || !workspace) return fail(SELDQ_ERR_INVALID, "seldq_attention_bwd: null pointer");
  if (workspace_bytes < seldq_attention_bwd_workspace_bytes(d) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(SELDQ_ERR_WORKSPACE, "seldq_attention_bwd: workspace of %zu B (256-byte aligned) needed", seldq_attention_bwd_workspace_bytes(d));
  int rc = cuda_ready();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const AttnBufs b = attn_saved_layout(d, const_cast<void*>(saved));
  char* w = (char*)workspace;
  char* do_rm = w; char* do_tr = w + b.rm; float* delta = reinterpret_cast<float*>(w + b.rm + b.tr);
  if ((rc = launch_attention_stage(d_out, out, do_rm, do_tr, delta, d->batch, d->heads, d->seq, d->head_dim, 1, 1.f, st))) return rc;
  return launch_attention_bwd(d->batch, d->heads, d->seq, d->head_dim, b.q_rm, b.k_rm, b.v_rm, b.q_tr, b.k_tr, do_rm, do_tr,
                              const_cast<float*>(lse), delta, dq, dk, dv, st);
}

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
  p.norm_mul[0] = p.norm_mul[1] = 1.f;
  return launch_stft(p, n_batch * n_ch, (cudaStream_t)stream);
}

extern "C" int seldq_stft_features(const void* x, int32_t n_batch, int32_t n_ch, int64_t n_samples, int32_t nperseg,
                                   int32_t noverlap, int32_t cut_dc, int32_t output_phase, int32_t cut_last,
                                   const seldq_stft_options_t* opt, float* out, void* stream) {
  if (!x || !opt || (!out && !opt->stats) || n_batch <= 0 || n_ch <= 0)
    return fail(SELDQ_ERR_INVALID, "seldq_stft_features: bad arguments");
  stft::Params p{};
  int rc = stft_shape(n_samples, nperseg, noverlap, cut_dc, cut_last, &p.n_bins, &p.n_frames);
  if (rc) return rc;
  if ((rc = cuda_ready())) return rc;
  if (opt->input_int16) p.x16 = reinterpret_cast<const short*>(x);
  else p.x = reinterpret_cast<const float*>(x);
  p.out = out; p.n_samples = n_samples; p.n_ch = n_ch; p.hop = nperseg - noverlap;
  p.bin0 = cut_dc ? 1 : 0; p.output_phase = output_phase ? 1 : 0;
  for (int k = 0; k < 2; ++k) { p.norm_sub[k] = opt->mean[k]; p.norm_mul[k] = opt->inv_std[k]; }
  p.stats = opt->stats;
  return launch_stft(p, n_batch * n_ch, (cudaStream_t)stream);
}

// Host-side translation of the C-ABI descriptors (include/seldq.h) into kernel geometry.
// Shared by the library (seldq_api.cu) and the CPU emulation harness (tests/host_emul).
#pragma once
#include "common.cuh"

namespace seldq {

inline int out_extent(int in, int k, int s, int p, int d) { return (in + 2 * p - d * (k - 1) - 1) / s + 1; }

inline int validate_conv(const seldq_conv_desc_t* d, BlockTable* tab, int* oh, int* ow) {
  if (!d) return fail(SELDQ_ERR_INVALID, "null convolution descriptor");
  if (!make_block_table(d->algebra, false, tab)) return fail(SELDQ_ERR_INVALID, "unknown algebra %d", d->algebra);
  if (d->precision != SELDQ_PREC_FP32 && d->precision != SELDQ_PREC_BF16)
    return fail(SELDQ_ERR_INVALID, "unknown precision %d", d->precision);
  if (d->ndim != 1 && d->ndim != 2)
    return fail(SELDQ_ERR_UNSUPPORTED, "only 1-d (NCW) and 2-d (NCHW) convolutions are implemented, got ndim=%d", d->ndim);
  if (d->ndim == 1 && (d->in_h != 1 || d->k_h != 1 || d->stride_h != 1 || d->pad_h != 0 || d->dil_h != 1))
    return fail(SELDQ_ERR_INVALID, "1-d convolution must have in_h=k_h=stride_h=dil_h=1, pad_h=0");
  if (d->batch <= 0 || d->cin <= 0 || d->cout <= 0 || d->in_h <= 0 || d->in_w <= 0 || d->k_h <= 0 || d->k_w <= 0 ||
      d->stride_h <= 0 || d->stride_w <= 0 || d->pad_h < 0 || d->pad_w < 0 || d->dil_h <= 0 || d->dil_w <= 0)
    return fail(SELDQ_ERR_INVALID, "non-positive size in convolution descriptor");
  if (d->cin % tab->nc || d->cout % tab->nc)
    // quaternion_ops.py:53-65 raises the same complaint
    return fail(SELDQ_ERR_INVALID, "Quaternion Tensors must be divisible by %d. cin=%d cout=%d", tab->nc, d->cin, d->cout);
  *oh = out_extent(d->in_h, d->k_h, d->stride_h, d->pad_h, d->dil_h);
  *ow = out_extent(d->in_w, d->k_w, d->stride_w, d->pad_w, d->dil_w);
  if (*oh <= 0 || *ow <= 0) return fail(SELDQ_ERR_INVALID, "convolution output would be empty (%d x %d)", *oh, *ow);
  return SELDQ_OK;
}

// pass: SELDQ_PASS_FWD, _DGRAD (transposed), _WGRAD (forward orientation, in = x, out = gy)
inline int make_conv_geom(const seldq_conv_desc_t* d, int pass, ConvGeom* g) {
  int oh, ow;
  memset(g, 0, sizeof(*g));
  const int rc = validate_conv(d, &g->tab, &oh, &ow);
  if (rc) return rc;
  const int nc = g->tab.nc;
  g->N = d->batch;
  g->Oc = d->cout / nc;
  g->Ic = d->cin / nc;
  g->KH = d->k_h; g->KW = d->k_w;
  g->sh = d->stride_h; g->sw = d->stride_w;
  g->ph = d->pad_h; g->pw = d->pad_w;
  g->dh = d->dil_h; g->dw = d->dil_w;
  g->wsT = 1; g->wsI = d->k_h * d->k_w; g->wsO = g->Ic * g->wsI;
  if (d->algebra == SELDQ_ALG_Q_LINEAR_IO || d->algebra == SELDQ_ALG_DQ_LINEAR_IO) {
    // the linear layer's own compact tensors, (in/nc, out/nc) row-major
    if (d->k_h != 1 || d->k_w != 1) return fail(SELDQ_ERR_INVALID, "the linear-layer algebras serve 1 x 1 kernels only");
    g->wsO = 1; g->wsI = g->Oc; g->wsT = 0;
  }
  const long long xs[4] = {(long long)d->cin * d->in_h * d->in_w, (long long)d->in_h * d->in_w, d->in_w, 1};
  const long long ys[4] = {(long long)d->cout * oh * ow, (long long)oh * ow, ow, 1};
  if (pass == SELDQ_PASS_DGRAD) {
    g->transposed = 1;
    g->P = d->cin; g->R = d->cout;
    g->OH = d->in_h; g->OW = d->in_w; g->IH = oh; g->IW = ow;
    g->in_sN = ys[0]; g->in_sC = ys[1]; g->in_sH = ys[2]; g->in_sW = ys[3];
    g->out_sN = xs[0]; g->out_sC = xs[1]; g->out_sH = xs[2]; g->out_sW = xs[3];
  } else {
    g->transposed = 0;
    g->P = d->cout; g->R = d->cin;
    g->OH = oh; g->OW = ow; g->IH = d->in_h; g->IW = d->in_w;
    g->in_sN = xs[0]; g->in_sC = xs[1]; g->in_sH = xs[2]; g->in_sW = xs[3];
    g->out_sN = ys[0]; g->out_sC = ys[1]; g->out_sH = ys[2]; g->out_sW = ys[3];
  }
  return SELDQ_OK;
}

inline int make_linear_geom(const seldq_linear_desc_t* d, int pass, ConvGeom* g) {
  memset(g, 0, sizeof(*g));
  if (!d) return fail(SELDQ_ERR_INVALID, "null linear descriptor");
  if (d->algebra == SELDQ_ALG_REAL || !make_block_table(d->algebra, true, &g->tab))
    return fail(SELDQ_ERR_INVALID, "linear layers exist for the Q and DQ algebras only, got %d", d->algebra);
  if (d->rows <= 0 || d->in_features <= 0 || d->out_features <= 0)
    return fail(SELDQ_ERR_INVALID, "non-positive size in linear descriptor");
  const int nc = g->tab.nc;
  if (d->in_features % nc || d->out_features % nc)
    return fail(SELDQ_ERR_INVALID, "Quaternion Tensors must be divisible by %d. in=%d out=%d", nc, d->in_features,
                d->out_features);
  g->N = 1;
  g->Oc = d->out_features / nc;
  g->Ic = d->in_features / nc;
  g->KH = g->KW = 1; g->sh = g->sw = 1; g->dh = g->dw = 1;
  g->OH = g->IH = 1; g->OW = g->IW = d->rows;
  // compact linear weights are stored (in/nc, out/nc): quaternion_layers.py:235-238
  g->wsO = 1; g->wsI = g->Oc; g->wsT = 0;
  if (pass == SELDQ_PASS_DGRAD) {
    g->transposed = 1;
    g->P = d->in_features; g->R = d->out_features;
    g->in_sC = 1; g->in_sW = d->out_features;
    g->out_sC = 1; g->out_sW = d->in_features;
  } else {
    g->transposed = 0;
    g->P = d->out_features; g->R = d->in_features;
    g->in_sC = 1; g->in_sW = d->in_features;
    g->out_sC = 1; g->out_sW = d->out_features;
  }
  return SELDQ_OK;
}

inline int stft_shape(long long n_samples, int nperseg, int noverlap, int cut_dc, int cut_last, int* n_bins,
                      int* n_frames) {
  if (nperseg != 512) return fail(SELDQ_ERR_UNSUPPORTED, "only nperseg = 512 is implemented, got %d", nperseg);
  if (noverlap < 0 || noverlap >= nperseg) return fail(SELDQ_ERR_INVALID, "noverlap must be in [0, nperseg)");
  if (n_samples <= 0) return fail(SELDQ_ERR_INVALID, "empty signal");
  const int hop = nperseg - noverlap;
  // scipy.signal.stft: boundary='zeros' extends by nperseg/2 on both sides, padded=True zero-pads
  // the tail so that (len - nperseg) % hop == 0
  const long long ext = n_samples + nperseg;
  const long long nadd = ((-(ext - nperseg)) % hop + hop) % hop;
  const long long frames = (ext + nadd - nperseg) / hop + 1;
  *n_bins = nperseg / 2 + 1 - (cut_dc ? 1 : 0);
  *n_frames = (int)(frames - (cut_last ? 1 : 0));
  if (*n_frames <= 0) return fail(SELDQ_ERR_INVALID, "no frames left after cut_last");
  return SELDQ_OK;
}

}  // namespace seldq

// TEST INFRASTRUCTURE: runs the __host__ __device__ phase functions of the SIMT kernels
// (csrc/conv_simt.cuh, csrc/stft.cuh) on the CPU, block by block and thread by thread, with the
// same launch geometry the library uses.  Lets the index arithmetic be checked against the oracle
// without a GPU.  Built by tests/test_host_emul.py with g++.
#include <cmath>
#include <vector>

#include "../../sound-event-localization-and-detection_b200/csrc/conv_simt.cuh"
#include "../../sound-event-localization-and-detection_b200/csrc/geom.h"
#include "../../sound-event-localization-and-detection_b200/csrc/stft.cuh"

using namespace seldq;

static void run_conv(const simt::ConvParams& p) {
  const ConvGeom& g = p.g;
  const int gx = (g.OW + simt::BN - 1) / simt::BN, gy = (g.P + simt::BM - 1) / simt::BM, gz = g.N * g.OH;
  std::vector<simt::ConvThread> th(simt::NT);
  simt::ConvShared sh;
  for (int bz = 0; bz < gz; ++bz)
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        for (auto& t : th) simt::conv_init(t);
        for (int tap = 0; tap < g.KH * g.KW; ++tap)
          for (int r0 = 0; r0 < g.R; r0 += simt::BK) {
            if (simt::conv_chunk_is_zero(g, by * simt::BM, r0)) continue;
            for (int t = 0; t < simt::NT; ++t) simt::conv_load(p, sh, t, bx, by, bz, tap, r0);
            for (int t = 0; t < simt::NT; ++t) simt::conv_mac(sh, th[t], t, simt::conv_swap(g));
          }
        for (int t = 0; t < simt::NT; ++t) simt::conv_store(p, th[t], t, bx, by, bz);
      }
}

static void run_wgrad(const simt::WgradParams& p) {
  const ConvGeom& g = p.g;
  const int ntile = ((g.Oc + simt::WT - 1) / simt::WT) * ((g.Ic + simt::WT - 1) / simt::WT);
  const int ntap = g.KH * g.KW;
  std::vector<simt::WgradThread> th(simt::NT);
  simt::WgradShared sh;
  for (int bz = 0; bz < p.splits; ++bz)
    for (int by = 0; by < g.tab.nw * ntap; ++by)
      for (int bx = 0; bx < ntile; ++bx) {
        const long long units = simt::wgrad_units(g, by / ntap);
        const long long per = (units + p.splits - 1) / p.splits;
        const long long u0 = bz * per, u1 = std::min(units, u0 + per);
        for (auto& t : th) simt::wgrad_init(t);
        for (long long u = u0; u < u1; ++u) {
          for (int t = 0; t < simt::NT; ++t) simt::wgrad_load(p, sh, t, bx, by, u);
          for (int t = 0; t < simt::NT; ++t) simt::wgrad_mac(sh, th[t], t);
        }
        for (int t = 0; t < simt::NT; ++t)
          simt::wgrad_store(p, th[t], t, bx, by, [](float* a, float v) { *a += v; });
      }
}

extern "C" {

const char* emul_last_error() { return error_buffer(); }

int emul_conv(const seldq_conv_desc_t* d, int pass, const float* in, const float* const* w, const float* bias,
              float* out) {
  simt::ConvParams p{};
  int rc = make_conv_geom(d, pass, &p.g);
  if (rc) return rc;
  p.in = in; p.out = out; p.bias = bias;
  for (int i = 0; i < p.g.tab.nw; ++i) p.w[i] = w[i];
  run_conv(p);
  return 0;
}

int emul_conv_wgrad(const seldq_conv_desc_t* d, const float* x, const float* gy, float* const* gw, int splits) {
  simt::WgradParams p{};
  int rc = make_conv_geom(d, SELDQ_PASS_WGRAD, &p.g);
  if (rc) return rc;
  p.x = x; p.gy = gy; p.splits = splits;
  for (int i = 0; i < p.g.tab.nw; ++i) p.gw[i] = gw[i];
  run_wgrad(p);
  return 0;
}

int emul_linear(const seldq_linear_desc_t* d, int pass, const float* in, const float* const* w, const float* bias,
                float* out) {
  simt::ConvParams p{};
  int rc = make_linear_geom(d, pass, &p.g);
  if (rc) return rc;
  p.in = in; p.out = out; p.bias = bias;
  for (int i = 0; i < p.g.tab.nw; ++i) p.w[i] = w[i];
  run_conv(p);
  return 0;
}

int emul_linear_wgrad(const seldq_linear_desc_t* d, const float* x, const float* gy, float* const* gw, int splits) {
  simt::WgradParams p{};
  int rc = make_linear_geom(d, SELDQ_PASS_WGRAD, &p.g);
  if (rc) return rc;
  p.x = x; p.gy = gy; p.splits = splits;
  for (int i = 0; i < p.g.tab.nw; ++i) p.gw[i] = gw[i];
  run_wgrad(p);
  return 0;
}

int emul_stft(const float* x, int n_batch, int n_ch, long long n_samples, int nperseg, int noverlap, int cut_dc,
              int output_phase, int cut_last, float* out) {
  int n_bins, n_frames;
  int rc = stft_shape(n_samples, nperseg, noverlap, cut_dc, cut_last, &n_bins, &n_frames);
  if (rc) return rc;
  stft::Params p{};
  p.x = x; p.out = out; p.n_samples = n_samples; p.n_ch = n_ch; p.hop = nperseg - noverlap;
  p.n_frames = n_frames; p.bin0 = cut_dc ? 1 : 0; p.n_bins = n_bins; p.output_phase = output_phase;
  p.groups = (n_frames + stft::FRB - 1) / stft::FRB;
  p.total = (long long)n_batch * n_ch * p.groups;
  stft::Shared* sh = new stft::Shared;
  std::vector<stft::Thread> th(stft::NT);
  for (int t = 0; t < stft::NT; ++t) stft::init_tables(*sh, t);
  for (long long batch = 0; batch < p.total; ++batch) {
    int signal, t0;
    stft::batch_decode(p, batch, &signal, &t0);
    for (int t = 0; t < stft::NT; ++t) stft::phase_a(p, *sh, th[t], t, signal, t0);
    for (int t = 0; t < stft::NT; ++t) stft::phase_b(*sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_b2(*sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_c(p, *sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_d(p, *sh, t, signal, t0);
  }
  delete sh;
  return 0;
}

}  // extern "C"

// TEST INFRASTRUCTURE: runs the __host__ __device__ phase functions of the SIMT kernels
// (csrc/conv_simt.cuh, csrc/stft.cuh) on the CPU, block by block and thread by thread, with the
// same launch geometry the library uses.  Lets the index arithmetic be checked against the oracle
// without a GPU.  Built by tests/test_host_emul.py with g++.
#include <cmath>
#include <vector>

#include "../../sound-event-localization-and-detection_b200/csrc/conv_cl_plan.h"
#include "../../sound-event-localization-and-detection_b200/csrc/conv_simt.cuh"
#include "../../sound-event-localization-and-detection_b200/csrc/geom.h"
#include "../../sound-event-localization-and-detection_b200/csrc/rotation.cuh"
#include "../../sound-event-localization-and-detection_b200/csrc/stft.cuh"
#include "../../sound-event-localization-and-detection_b200/csrc/stft_pair.cuh"

using namespace seldq;


// ---- channels-last tensor-core path (csrc/conv_cl.cu), data flow on the CPU -----------------------------------------
// Runs the REAL host plan (cl::plan_fprop: fusion sets, op table, epilogue column / sign table, stage geometry, chunk
// masks, group schedule) and the REAL pack mapping (cl::pack_dst_offset) through a plain-loop model of what the kernel
// does with them: TMA boxes with zero fill, one "MMA" per op-table entry (A rows x B rows, K = 16, sign = negate-B
// bit, every accumulator starting from zero as the epilogue warps leave it), accumulator column sets, epilogue combination.  No rounding to bf16: the result
// must match the expanded-weight convolution to float accuracy, so any table or offset error shows as O(1).
static int run_cl_fprop(const ConvGeom& g, const float* in_nchw, const float* const* w, float* out, int n_sms,
                        int* info, int nprob = 1) {
  cl::FpropParams* pp = new cl::FpropParams;
  cl::FpropParams& p = *pp;
  size_t smem = 0;
  int rc = cl::plan_fprop(g, &p, &smem, n_sms, nprob);
  if (rc) { delete pp; return rc; }
  {
    // the kernel's schedule: grid = nprob * min(units, n_sms / nprob) CTAs, CTA x serves problem x % nprob as
    // number x / nprob of its G CTAs; every unit of every problem must be visited exactly once
    const int per = n_sms / nprob, G = p.total_units < per ? p.total_units : per;
    std::vector<int> seen((size_t)nprob * p.total_units, 0);
    for (int x = 0; x < nprob * G; ++x) {
      const int prob = nprob == 2 ? (x & 1) : 0, cta = nprob == 2 ? (x >> 1) : x;
      for (int round = 0; round * G < p.total_units; ++round) {
        const int u = cl::unit_of_round(round, p.total_units, G, cta);
        if (u >= 0) ++seen[(size_t)prob * p.total_units + u];
      }
    }
    for (size_t i = 0; i < seen.size(); ++i)
      if (seen[i] != 1) { delete pp; return fail(SELDQ_ERR_INVALID, "schedule: unit %zu visited %d times", i, seen[i]); }
  }
  if (info) { info[0] = p.fuse; info[1] = p.pair_xor; info[2] = p.gc; info[3] = p.ngroups; info[4] = p.rs; info[5] = p.tps;
              info[6] = p.acc_cols; info[7] = p.acc_stages; info[8] = p.nstages; info[9] = (int)smem; }
  const cl::WeightPlan wp = cl::weight_plan(g);
  // packed weights as floats, 2 "bytes" per element like the bf16 image
  std::vector<float> packed(wp.total / 2, 0.f);
  cl::PackParams pk{};
  cl::fill_pack_params(g, wp, &pk);
  const int rows = cl::pack_rows(pk);
  if (p.dense) {
    // dense mode (K side narrower than 8 channels per component, or the real algebra): the signed EXPANDED tile is
    // built in shared memory by the kernel's prologue, tile (tap, j) = B[n][k], n = out channel, k = in channel 16 j + k
    pk.n_img = 0;
    for (int tap = 0; tap < p.ntaps; ++tap)
      for (int j = 0; j < p.J; ++j)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < p.NBp; ++n)
            for (int jj = 0; jj < 8; ++jj) {
              const int ch = j * 16 + kc * 8 + jj;
              const size_t off = (size_t)(tap * p.J + j) * p.slab_bytes + (size_t)kc * (p.NBp * 16) + (n >> 3) * 128 + (n & 7) * 16;
              packed[off / 2 + jj] = (n < g.P && ch < g.R) ? expanded_weight(g, w, n, ch, tap) : 0.f;
            }
  }
  for (int img = 0; img < pk.n_img; ++img)
    for (int tap = 0; tap < pk.ntaps; ++tap)
      for (int j = 0; j < pk.J; ++j)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < rows; ++n) {
            const size_t off = cl::pack_dst_offset(pk, img, tap, j, kc, n);
            if (off + 16 > wp.total) { delete pp; return fail(SELDQ_ERR_INVALID, "pack offset %zu beyond %zu", off, wp.total); }
            for (int jj = 0; jj < 8; ++jj) {
              const int k = j * 16 + kc * 8 + jj;
              float x = 0.f;
              if (n < pk.rows_real && k < pk.k_real) {
                const int o = pk.transposed ? k : n, i = pk.transposed ? n : k;
                x = w[img][(long long)o * pk.wsO + (long long)i * pk.wsI + (long long)tap * pk.wsT];
              }
              packed[off / 2 + jj] = x;
            }
          }
  // the channels-last operand: [n][h][w][Cp], every component padded to cpad channels
  const cl::OperandLayout l = cl::operand_layout(g.tab.nc, g.R, p.dense != 0);
  std::vector<float> xcl((size_t)g.N * g.IH * g.IW * l.Cp, 0.f);
  for (int n = 0; n < g.N; ++n)
    for (int c = 0; c < g.R; ++c)
      for (int h = 0; h < g.IH; ++h)
        for (int x = 0; x < g.IW; ++x)
          xcl[(((size_t)n * g.IH + h) * g.IW + x) * l.Cp + (c / l.cc) * l.cpad + c % l.cc] =
              in_nchw[n * g.in_sN + c * g.in_sC + h * g.in_sH + x * g.in_sW];
  auto a_elem = [&](int n, int h, int x, int ch) -> double {          // TMA: out-of-range coordinates read zero
    if (h < 0 || h >= g.IH || x < 0 || x >= g.IW) return 0.0;
    return xcl[(((size_t)n * g.IH + h) * g.IW + x) * l.Cp + ch];
  };
  const int lpc = p.slabs_per_chunk * p.mma_per_slab;
  std::vector<double> acc((size_t)cl::kTileM * 512);
  for (int u = 0; u < p.total_units; ++u) {
    const int group = p.group_order[u / p.total_tiles];
    int r = u % p.total_tiles;
    const int wt = r % p.tiles_w; r /= p.tiles_w;
    const int h = r % p.OH, n = r / p.OH, w0 = wt * cl::kTileM;
    std::fill(acc.begin(), acc.end(), 0.0);       // the epilogue warps clear every accumulator (tcgen05.st): all MMAs accumulate
    for (int c = 0; c < p.chunks; ++c) {
      if (!((p.chunk_mask[group] >> c) & 1u)) continue;
      for (int tap0 = 0; tap0 < p.ntaps; tap0 += p.tps)
        for (int tl = 0; tl < p.tps; ++tl)
          for (int e_i = 0; e_i < lpc; ++e_i) {
            const uint2 e = p.op_tbl[((size_t)group * p.chunks + c) * lpc + e_i];
            if (!((int)e.x < 0)) continue;
            const int tap = tap0 + tl;
            const int col = (e.x >> 20) & 0x1ff, slab = (e.x >> 16) & 3;
            const size_t tile = ((size_t)(e.x & 0x3fffu) + (size_t)tap * p.tap_stride16) * 16;   // bytes
            const int nmma = (e.y >> 17) & 0x3f, neg = (e.y >> 14) & 1;
            if (nmma * 8 != p.NBmma) { delete pp; return fail(SELDQ_ERR_INVALID, "idesc N %d != NBmma %d", nmma * 8, p.NBmma); }
            if (tile + (size_t)p.NBmma * 32 > wp.total) { delete pp; return fail(SELDQ_ERR_INVALID, "weight tile beyond the packed buffer"); }
            if (col + p.NBmma > p.acc_cols) { delete pp; return fail(SELDQ_ERR_INVALID, "accumulator columns %d+%d beyond %d", col, p.NBmma, p.acc_cols); }
            // where the tap's A rows start: its own box, or rs_row rows into the kernel row's shared box
            const int x_base = p.rs ? w0 + p.rs_min_off + p.rs_row[tl] : w0 + p.off_w[tap];
            const int hh = h + p.off_h[p.rs ? tap0 : tap];
            for (int m = 0; m < cl::kTileM; ++m)
              for (int nn = 0; nn < p.NBmma; ++nn) {
                double sacc = 0.0;
                for (int k = 0; k < 16; ++k) {
                  const double a = a_elem(n, hh, x_base + m, c * p.BK + slab * 16 + k);
                  const double b = packed[(tile + (size_t)(k >> 3) * (p.NBmma * 16) + (nn >> 3) * 128 + (nn & 7) * 16) / 2 + (k & 7)];
                  sacc += a * b;
                }
                acc[(size_t)m * 512 + col + nn] += neg ? -sacc : sacc;
              }
          }
    }
    // epilogue
    for (int al = 0; al < p.gc; ++al) {
      const int ch_base = p.comp_of[group][al] * p.Pc;
      for (int o = 0; o < p.Pc; ++o)
        for (int m = 0; m < cl::kTileM; ++m) {
          if (w0 + m >= p.OW) continue;
          double v = 0.0;
          if (!p.fuse) v = acc[(size_t)m * 512 + al * p.NBp + o];
          else
            for (int st = 0; st < 4; ++st)
              if (p.epi_sgn[group][al][st] != 0)
                v += p.epi_sgn[group][al][st] * acc[(size_t)m * 512 + p.epi_col[group][al][st] + (o >> 3) * 8 * p.fuse + (o & 7)];
          out[n * g.out_sN + (ch_base + o) * g.out_sC + h * g.out_sH + (w0 + m) * g.out_sW] = (float)v;
        }
    }
  }
  delete pp;
  return 0;
}

// ---- weight-gradient kernel (csrc/wgrad_cl.cu), data flow on the CPU ---------------------------------------------------
// Real host plan (cl::plan_wgrad: o tiles, tap groups, split-K ranges, the (a, b) -> compact tensor fold table) through
// a plain-loop model of the kernel: per CTA the dense [128 rows = (component, out channel)] x [taps x Cp in channels]
// accumulator over its slice of the position axis (operand boxes with zero fill), then the sign-weighted fold onto
// the compact gradients, accumulated like the kernel's atomicAdd.
static int run_cl_wgrad(const ConvGeom& g, const float* x_nchw, const float* gy_nchw, float* const* gw, int n_sms,
                        int* info) {
  cl::WgradParams* pp = new cl::WgradParams;
  size_t smem = 0;
  int rc = cl::plan_wgrad(g, 1, n_sms, pp, &smem);
  if (rc) { delete pp; return rc; }
  const cl::WgradParams& p = *pp;
  if (info) { info[0] = p.o_tiles; info[1] = p.tap_groups; info[2] = p.taps_per_group; info[3] = p.splits;
              info[4] = p.nstages; info[5] = (int)smem; info[6] = p.tmem_cols; info[7] = p.Cp; }
  const cl::OperandLayout l = cl::operand_layout(g.tab.nc, g.tab.nc * g.Ic, false);
  std::vector<float> xcl((size_t)g.N * g.IH * g.IW * l.Cp, 0.f);
  for (int n = 0; n < g.N; ++n)
    for (int c = 0; c < g.R; ++c)
      for (int h = 0; h < g.IH; ++h)
        for (int x = 0; x < g.IW; ++x)
          xcl[(((size_t)n * g.IH + h) * g.IW + x) * l.Cp + (c / l.cc) * l.cpad + c % l.cc] =
              x_nchw[n * g.in_sN + c * g.in_sC + h * g.in_sH + x * g.in_sW];
  const int tiles = p.nprob * p.tap_groups * p.o_tiles;
  std::vector<double> acc((size_t)128 * 512);
  for (int bx = 0; bx < tiles; ++bx)
    for (int by = 0; by < p.splits; ++by) {
      const int o_tile = bx % p.o_tiles, tg = (bx / p.o_tiles) % p.tap_groups;
      const int tap0 = tg * p.taps_per_group;
      const int ntap = p.taps_per_group < p.ntaps - tap0 ? p.taps_per_group : p.ntaps - tap0;
      const int o0 = o_tile * p.OS;
      const long long per = (p.ksteps + p.splits - 1) / p.splits;
      const long long k_begin = (long long)by * per, k_end = k_begin + per < p.ksteps ? k_begin + per : p.ksteps;
      if (k_end <= k_begin) continue;
      if (ntap * p.Cp > p.tmem_cols) { delete pp; return fail(SELDQ_ERR_INVALID, "accumulator beyond the TMEM allocation"); }
      std::fill(acc.begin(), acc.end(), 0.0);
      for (long long ks = k_begin; ks < k_end; ++ks) {
        long long u = ks;
        const int wc = (int)(u % p.chunks_w); u /= p.chunks_w;
        const int h = (int)(u % p.OH), n = (int)(u / p.OH), w0 = wc * 64;
        for (int a = 0; a < p.ncomp; ++a)
          for (int ol = 0; ol < p.OS; ++ol) {
            const int o = o0 + ol;
            if (o >= g.Oc) continue;                                    // rows beyond the component read zero
            for (int t = 0; t < 64; ++t) {
              if (w0 + t >= g.OW) continue;
              const double gyv = gy_nchw[n * g.out_sN + (a * g.Oc + o) * g.out_sC + h * g.out_sH + (w0 + t) * g.out_sW];
              if (gyv == 0.0) continue;
              for (int tl = 0; tl < ntap; ++tl) {
                const int hh = h + p.off_h[tap0 + tl], xx = w0 + t + p.off_w[tap0 + tl];
                if (hh < 0 || hh >= g.IH || xx < 0 || xx >= g.IW) continue;
                const float* xr = &xcl[(((size_t)n * g.IH + hh) * g.IW + xx) * l.Cp];
                double* ar = &acc[(size_t)(a * p.OS + ol) * 512 + tl * p.Cp];
                for (int ch = 0; ch < p.Cp; ++ch) ar[ch] += gyv * xr[ch];
              }
            }
          }
      }
      // epilogue: fold the (a, b) blocks onto the compact tensors
      for (int tl = 0; tl < ntap; ++tl)
        for (int e = 0; e < g.tab.nw; ++e)
          for (int ol = 0; ol < p.OS; ++ol)
            for (int i = 0; i < g.Ic; ++i) {
              if (o0 + ol >= g.Oc) continue;
              double v = 0.0;
              for (int k = 0; k < p.pair_n[e]; ++k) {
                const double val = acc[(size_t)(p.pair_a[e][k] * p.OS + ol) * 512 + tl * p.Cp + p.pair_b[e][k] * p.cpad_in + i];
                v += p.pair_neg[e][k] ? -val : val;
              }
              gw[e][(long long)(o0 + ol) * g.wsO + (long long)i * g.wsI + (long long)(tap0 + tl) * g.wsT] += (float)v;
            }
    }
  delete pp;
  return 0;
}

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

// frame-pair kernel (csrc/stft_pair.cuh): PAIRS = 28 / 16 magnitude only, PAIRS = 12 with the phase plane
template <int PAIRS, int PLANES>
static int run_stft_pairs(stft::Params& p, int n_signals) {
  p.groups = (p.n_frames + 2 * PAIRS - 1) / (2 * PAIRS);
  p.total = (long long)n_signals * p.groups;
  constexpr int NT = 16 * PAIRS;
  stft2::Shared<PAIRS, PLANES>* sh = new stft2::Shared<PAIRS, PLANES>;
  std::vector<stft2::Thread> th(NT);
  std::vector<stft2::Raw> raw(NT);
  for (int t = 0; t < NT; ++t) stft2::init_tables(*sh, t);
  for (long long batch = 0; batch < p.total; ++batch) {
    int signal, t0;
    stft2::batch_decode<PAIRS>(p, batch, &signal, &t0);
    for (int t = 0; t < NT; ++t) stft2::load_raw(p, raw[t], t, signal, t0);
    for (int t = 0; t < NT; ++t) stft2::phase_a(*sh, raw[t], th[t], t);
    for (int t = 0; t < NT; ++t) stft2::phase_b(*sh, th[t], t);
    for (int t = 0; t < NT; ++t) stft2::phase_b2(*sh, th[t], t);
    for (int t = 0; t < NT; ++t) stft2::phase_c(p, *sh, th[t], t);
    for (int t = 0; t < NT; ++t) stft2::phase_d(p, *sh, t, signal, t0);
  }
  delete sh;
  return 0;
}

extern "C" {

// forward (pass 0) or dgrad (pass 1) of a Q / DQ convolution through the emulated tensor-core data flow; info (10
// ints, may be NULL) reports the plan: fuse, pair_xor, gc, ngroups, rs, tps, acc_cols, acc_stages, nstages, smem bytes
int emul_cl_conv(const seldq_conv_desc_t* d, int pass, const float* in, const float* const* w, float* out, int n_sms,
                 int* info) {
  ConvGeom g;
  int rc = make_conv_geom(d, pass, &g);
  if (rc) return rc;
  return run_cl_fprop(g, in, w, out, n_sms, info);
}
// the plan of a SIBLING launch (two problems, n_sms / 2 CTAs each; conv_cl.h) run for one of its problems: the data
// flow per problem is the single launch's, the plan (groups, accumulator buffering) and the CTA schedule differ
int emul_cl_conv_pair(const seldq_conv_desc_t* d, int pass, const float* in, const float* const* w, float* out,
                      int n_sms, int* info) {
  ConvGeom g;
  int rc = make_conv_geom(d, pass, &g);
  if (rc) return rc;
  return run_cl_fprop(g, in, w, out, n_sms, info, 2);
}
// compact weight gradients (accumulated into gw, which the caller zeroes) through the emulated weight-gradient kernel;
// info (8 ints, may be NULL): o_tiles, tap_groups, taps_per_group, splits, nstages, smem bytes, tmem_cols, Cp
int emul_cl_conv_wgrad(const seldq_conv_desc_t* d, const float* x, const float* gy, float* const* gw, int n_sms,
                       int* info) {
  ConvGeom g;
  int rc = make_conv_geom(d, SELDQ_PASS_WGRAD, &g);
  if (rc) return rc;
  return run_cl_wgrad(g, x, gy, gw, n_sms, info);
}
int emul_cl_linear(const seldq_linear_desc_t* d, int pass, const float* in, const float* const* w, float* out,
                   int n_sms, int* info) {
  ConvGeom g;
  int rc = make_linear_geom(d, pass, &g);
  if (rc) return rc;
  return run_cl_fprop(g, in, w, out, n_sms, info);
}

// rotation weights and quaternion point-wise operators (csrc/rotation.cuh): the kernels' element functions, one call
// per thread index
int emul_rotation_weight(const float* const* w, long long d0, long long d1, long long taps, int qf, int tr, float* out) {
  rot::RotGeom g{d0, d1, taps, qf ? 4 : 3, tr ? 1 : 0};
  for (long long idx = 0; idx < d0 * d1 * taps; ++idx) rot::rot_fwd_element(g, idx, w[0], w[1], w[2], w[3], out);
  return 0;
}
int emul_rotation_weight_bwd(const float* const* w, const float* gout, long long d0, long long d1, long long taps, int qf,
                             int tr, float* const* gw) {
  rot::RotGeom g{d0, d1, taps, qf ? 4 : 3, tr ? 1 : 0};
  for (long long idx = 0; idx < d0 * d1 * taps; ++idx)
    rot::rot_bwd_element(g, idx, w[0], w[1], w[2], w[3], gout, gw[0], gw[1], gw[2], gw[3]);
  return 0;
}
int emul_qpointwise(int op, const float* a, const float* b, float* out, long long outer, long long m) {
  for (long long idx = 0; idx < outer * m; ++idx) rot::qpointwise_element(op, a, b, out, idx, m);
  return 0;
}

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
  p.norm_mul[0] = p.norm_mul[1] = 1.f;
  stft::Shared* sh = new stft::Shared;
  std::vector<stft::Thread> th(stft::NT);
  stft::Stats acc = {0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < stft::NT; ++t) stft::init_tables(*sh, t);
  for (long long batch = 0; batch < p.total; ++batch) {
    int signal, t0;
    stft::batch_decode(p, batch, &signal, &t0);
    for (int t = 0; t < stft::NT; ++t) stft::phase_a(p, *sh, th[t], t, signal, t0);
    for (int t = 0; t < stft::NT; ++t) stft::phase_b(*sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_b2(*sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_c(p, *sh, th[t], t);
    for (int t = 0; t < stft::NT; ++t) stft::phase_d(p, *sh, acc, t, signal, t0);
  }
  delete sh;
  return 0;
}

int emul_stft_pairs(const float* x, int n_batch, int n_ch, long long n_samples, int nperseg, int noverlap, int cut_dc,
                    int output_phase, int cut_last, int pairs, float* out) {
  int n_bins, n_frames;
  int rc = stft_shape(n_samples, nperseg, noverlap, cut_dc, cut_last, &n_bins, &n_frames);
  if (rc) return rc;
  stft::Params p{};
  p.x = x; p.out = out; p.n_samples = n_samples; p.n_ch = n_ch; p.hop = nperseg - noverlap;
  p.n_frames = n_frames; p.bin0 = cut_dc ? 1 : 0; p.n_bins = n_bins; p.output_phase = output_phase;
  p.norm_mul[0] = p.norm_mul[1] = 1.f;
  if (output_phase) return run_stft_pairs<12, 2>(p, n_batch * n_ch);
  return pairs == 16 ? run_stft_pairs<16, 1>(p, n_batch * n_ch) : run_stft_pairs<28, 1>(p, n_batch * n_ch);
}

}  // extern "C"

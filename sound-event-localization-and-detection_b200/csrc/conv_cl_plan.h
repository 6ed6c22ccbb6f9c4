// Host-side planner of the channels-last tensor-core convolution (conv_cl.cu): which components fuse into one
// tcgen05.mma, where every weight tile lives in shared memory, the per-(group, chunk) MMA op table, the epilogue's
// column / sign table, stage geometry.  Pure host C++ (no device code), so that tests/host_emul can run the plan
// through a CPU emulation of the kernel's data flow and check it against the expanded-weight convolution.
#pragma once
#include <cstdlib>

#include "conv_cl.h"
#include "umma_desc.h"

namespace seldq {
namespace cl {

// Unit schedule of a CTA (shared by the kernel and its CPU model): units are ordered heaviest group first.
// G = CTAs that share the problem, b = this CTA's index among them.  Returns -1 when the CTA sits a round out.
//   * many rounds (the CNN layers): round r hands unit r*G + b to CTA b in even rounds and r*G + (G-1-b) in odd rounds
//     (snake order), so every CTA gets a mix of heavy and light units;
//   * two or three rounds (the TCN launches: 152 units on 74 CTAs): a CTA's time is set by the NUMBER of its units -- each
//     ends in a ~2.3 us epilogue, heavy or light, and the epilogues of a CTA do not overlap each other (tools/
//     fprop_trace.py) -- so the `rem` CTAs that must take one unit more than the others take the LIGHTEST units (the tail
//     of the list) in every round, and the others share the rest in snake order: 64 x (heavy, light), 6 x (heavy, heavy),
//     4 x (light, light, light) instead of 4 x (heavy, light, light) finishing 2 us after everybody else.
SELDQ_HD int unit_of_round(int round, int total_units, int G, int b) {
  const int R = (total_units + G - 1) / G;            // rounds of the longest CTAs
  const int rem = total_units - (R - 1) * G;          // CTAs with R units
  if (R <= 1 || R > 3 || rem == G) {
    const int u = round * G + ((round & 1) ? G - 1 - b : b);
    return u < total_units ? u : -1;
  }
  if (b < rem) return round < R ? total_units - rem * R + round * rem + b : -1;
  if (round >= R - 1) return -1;
  const int Gs = G - rem, bb = b - rem;               // the head of the list holds exactly (R - 1) * Gs units
  return round * Gs + ((round & 1) ? Gs - 1 - bb : bb);
}

// compact fp32 weights -> bf16 UMMA B tiles [img][tap][j][NBp x 16] (K-major, no swizzle; see the dense
// prologue above for the byte layout).  Row n / column k of tile (img, tap, j):
//   forward : W_img[o = n][i = 16 j + k][tap]        dgrad : W_img[o = 16 j + k][i = n][tap]
// Fusion (pair_xor = m != 0; F = 2 for m = 1 | 2, F = 4 for m = 3): tile (image set q, tap, j) is [F NB8 x 16]; its
// 8-row groups cycle through the F images {i0 | sub : sub a submask of m} of the set:
// row (n / 8) * 8 F + slot * 8 + n % 8, slot = the bits of the image index under m.
struct PackParams {
  const float* w[8];
  uint8_t* dst;
  int n_img, ntaps, J, NBp;
  int rows_real, k_real;        // real extent of the row (N side) and K side
  int transposed;
  int wsO, wsI, wsT;
  int pair_xor, NB8;            // pair fusion; rows per image = NB8 then, NBp otherwise
};
SELDQ_HD int pack_rows(const PackParams& p) { return p.pair_xor ? p.NB8 : p.NBp; }
// position of image (or component) index i inside its fusion set, and the index of that set: the bits of i under
// the mask m, and the remaining bits squeezed together
SELDQ_HD int fuse_slot(int i, int m) { return m == 1 ? (i & 1) : m == 2 ? ((i >> 1) & 1) : (i & 3); }
SELDQ_HD int fuse_set(int i, int m) { return m == 1 ? (i >> 1) : m == 2 ? (((i >> 2) << 1) | (i & 1)) : (i >> 2); }

// byte offset of the 16-byte item (8 consecutive k of row n) of image `img`, tap, K slab j, K half kc in the packed
// weight buffer (the mapping pack_item in conv_cl.cu writes through)
SELDQ_HD size_t pack_dst_offset(const PackParams& p, int img, int tap, int j, int kc, int n) {
  if (p.pair_xor) {
    const int m = p.pair_xor, F = m == 3 ? 4 : 2;
    const int slot = fuse_slot(img, m), q = fuse_set(img, m);
    const int nf = (n >> 3) * (8 * F) + slot * 8 + (n & 7), NBf = F * p.NB8;
    return ((size_t)(q * p.ntaps + tap) * p.J + j) * ((size_t)NBf * 32) + (size_t)kc * (NBf * 16) + (nf >> 3) * 128 +
           (nf & 7) * 16;
  }
  return ((size_t)(img * p.ntaps + tap) * p.J + j) * ((size_t)p.NBp * 32) + (size_t)kc * (p.NBp * 16) + (n >> 3) * 128 +
         (n & 7) * 16;
}

// block (pass-out component a, pass-in component b) of the expanded weight: image index (-1: structural zero) and sign
inline void pass_block(const ConvGeom& g, int a, int b, int* img, int* neg) {
  const int fa = g.transposed ? b : a, fb = g.transposed ? a : b;   // forward-sense (out, in) components
  *img = g.tab.widx[fa][fb];
  *neg = g.tab.sign[fa][fb] < 0;
}

// Fusion (conv_cl.h): with the out components grouped as {a0 | sub : sub a submask of m} (pairs for m = 1 | 2, quads
// for m = 3), does every group see at most `max_sets` (slot order, relative signs) classes over the in components?
// Fills set_of[a0][b] (a0 = lowest member) with the class index of block column b, or -1 where the group has a
// structural zero, and returns the number of classes needed (0: this grouping does not work).
inline int fusion_sets(const ConvGeom& g, int m, int max_sets, int8_t set_of[8][8]) {
  const int nc = g.tab.nc, F = m == 3 ? 4 : 2;
  int need = 0;
  for (int a0 = 0; a0 < nc; ++a0) {
    if (a0 & m) continue;
    int nclass = 0, key[4] = {0, 0, 0, 0};
    for (int b = 0; b < nc; ++b) {
      int e[4], sg[4], nz = 0, used = 0, k = 0;
      for (int t = 0, sub = 0; t < F; ++t, sub = (sub - m) & m) {      // submasks of m in increasing order
        pass_block(g, a0 | sub, b, &e[t], &sg[t]);
        nz += e[t] >= 0;
      }
      set_of[a0][b] = -1;
      if (nz == 0) continue;
      if (nz != F) return 0;
      for (int t = 0; t < F; ++t) {
        if (fuse_set(e[t], m) != fuse_set(e[0], m)) return 0;
        used |= 1 << fuse_slot(e[t], m);
        k = (k << 3) | (fuse_slot(e[t], m) << 1) | (sg[t] ^ sg[0]);
      }
      if (used != (1 << F) - 1) return 0;
      int c = 0;
      while (c < nclass && key[c] != k) ++c;
      if (c == nclass) {
        if (nclass == max_sets) return 0;
        key[nclass++] = k;
      }
      set_of[a0][b] = (int8_t)c;
    }
    need = nclass > need ? nclass : need;
  }
  return need;
}

// geometry of the resident weight tiles of one pass
struct WeightPlan {
  int dense, n_img, ntaps, J, NBp, rows_real, k_real;
  int fuse, pair_xor, nsets, NB8, NBmma;   // fusion factor F (0 | 2 | 4): a tile holds the F images of a set, N = F * NB8
  int n_tilesets;                   // images, or image sets when fused
  size_t slab_bytes, img_bytes, total;
};
inline WeightPlan weight_plan(const ConvGeom& g) {
  WeightPlan w;
  memset(&w, 0, sizeof(w));
  w.dense = is_dense(g) ? 1 : 0;
  w.ntaps = g.KH * g.KW;
  if (w.dense) {
    const OperandLayout l = operand_layout(1, g.R, true);
    w.n_img = 1; w.J = l.Cp / 16; w.NBp = round_up(g.P, 16); w.rows_real = g.P; w.k_real = g.R;
  } else {
    const int kc = g.transposed ? g.Oc : g.Ic, pc = g.transposed ? g.Ic : g.Oc;
    w.n_img = g.tab.nw; w.J = round_up(kc, 16) / 16; w.NBp = round_up(pc, 16); w.rows_real = pc; w.k_real = kc;
    const bool enabled = getenv("SELDQ_PAIR_FUSE") == nullptr || atoi(getenv("SELDQ_PAIR_FUSE")) != 0;
    const bool quads = getenv("SELDQ_QUAD_FUSE") == nullptr || atoi(getenv("SELDQ_QUAD_FUSE")) != 0;
    if (enabled && g.tab.nc >= 4 && (pc & 7) == 0 && pc <= 128 && (w.n_img & 3) == 0) {
      int8_t scratch[8][8];
      // quads need four column sets of 4 * pc accumulator columns each: 16 * pc <= 512.  SELDQ_QUAD_FUSE=2 takes them
      // only where that fits TWICE (pc <= 16), i.e. where MMA and epilogue phases of successive units still overlap
      const int quad_cols = (getenv("SELDQ_QUAD_FUSE") && atoi(getenv("SELDQ_QUAD_FUSE")) == 2) ? 256 : 512;
      if (quads && 16 * pc <= quad_cols && (w.nsets = fusion_sets(g, 3, 4, scratch)) > 0) { w.fuse = 4; w.pair_xor = 3; }
      for (int m = 1; m <= 2 && !w.fuse; ++m)
        if ((w.nsets = fusion_sets(g, m, 2, scratch)) > 0) { w.fuse = 2; w.pair_xor = m; }
    }
  }
  w.NB8 = round_up(w.rows_real, 8);
  w.NBmma = w.fuse ? w.fuse * w.NB8 : w.NBp;
  w.n_tilesets = w.fuse ? w.n_img / w.fuse : w.n_img;
  w.slab_bytes = (size_t)w.NBmma * 32;
  w.img_bytes = (size_t)w.ntaps * w.J * w.slab_bytes;
  w.total = (size_t)w.n_tilesets * w.img_bytes;
  return w;
}

// fills the pack parameters of one layer / pass (the weight pointers and the destination are the caller's)
inline void fill_pack_params(const ConvGeom& g, const WeightPlan& w, PackParams* p) {
  p->n_img = w.n_img; p->ntaps = w.ntaps; p->J = w.J; p->NBp = w.NBp;
  p->rows_real = w.rows_real; p->k_real = w.k_real; p->transposed = g.transposed;
  p->wsO = g.wsO; p->wsI = g.wsI; p->wsT = g.wsT;
  p->pair_xor = w.fuse ? w.pair_xor : 0; p->NB8 = w.NB8;
}

// nprob = 2 plans a sibling launch (conv_cl.h): each problem has n_sms / 2 CTAs to itself
inline int plan_fprop(const ConvGeom& g, FpropParams* p, size_t* smem_bytes, int n_sms, int nprob = 1) {
  memset(p, 0, sizeof(*p));
  p->nprob = nprob;
  if (nprob == 2) n_sms /= 2;
  if (g.sh != 1 || g.sw != 1)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 tensor-core path implements stride 1 only (got %dx%d)", g.sh, g.sw);
  const int ntaps = g.KH * g.KW;
  if (ntaps > kMaxTaps) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most %d taps", kMaxTaps);
  const WeightPlan w = weight_plan(g);
  const int nc = g.tab.nc;
  p->g = g;
  p->ntaps = ntaps;
  p->N = g.N; p->OH = g.OH; p->OW = g.OW;
  for (int t = 0; t < ntaps; ++t) {
    const int kh = t / g.KW, kw = t % g.KW;
    // forward: in = out - pad + k*dil ; dgrad: gy position = gx position + pad - k*dil   (stride 1)
    p->off_h[t] = g.transposed ? g.ph - kh * g.dh : kh * g.dh - g.ph;
    p->off_w[t] = g.transposed ? g.pw - kw * g.dw : kw * g.dw - g.pw;
  }
  p->dense = w.dense;
  OperandLayout l;
  for (int b = 0; b < 8; ++b)
    for (int a = 0; a < 8; ++a) p->op_img[b][a] = -1;
  if (w.dense) {
    if (g.P > 256)
      return fail(SELDQ_ERR_UNSUPPORTED, "bf16 dense mode (K side < 8 channels per component) needs <= 256 out channels, got %d",
                  g.P);
    l = operand_layout(1, g.R, true);
    p->ncomp_out = 1; p->Pc = g.P;
    p->op_img[0][0] = 0; p->op_neg[0][0] = 0;
  } else {
    l = operand_layout(nc, g.R, false);
    p->ncomp_out = nc; p->Pc = g.transposed ? g.Ic : g.Oc;
    for (int b = 0; b < nc; ++b)
      for (int a = 0; a < nc; ++a) {
        const int fa = g.transposed ? b : a, fb = g.transposed ? a : b;   // forward-sense (out, in) components
        const int e = g.tab.widx[fa][fb];
        if (e < 0) continue;
        p->op_img[b][a] = (int8_t)e;
        p->op_neg[b][a] = (int8_t)(g.tab.sign[fa][fb] < 0);
      }
  }
  p->NBp = w.NBp; p->J = w.J; p->n_img = w.n_img;
  p->slab_bytes = (uint32_t)w.slab_bytes; p->img_bytes = (uint32_t)w.img_bytes; p->w_bytes = (uint32_t)w.total;
  p->BK = l.BK; p->chunks = l.Cp / l.BK; p->slabs_per_chunk = l.BK / 16; p->cpad_in = l.cpad;
  if (p->chunks > 32) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most 2048 padded input channels");
  p->box_bytes = (uint32_t)kTileM * l.BK * 2;
  p->tps = 1;
  p->a_sbo = 8u * l.BK * 2;
  p->a_swz = l.BK == 64 ? ptx::kSwizzle128B : (l.BK == 32 ? ptx::kSwizzle64B : ptx::kSwizzle32B);

  p->tiles_w = (g.OW + kTileM - 1) / kTileM;
  const long long tiles = (long long)g.N * g.OH * p->tiles_w;
  if (tiles > 0x0fffffffLL) return fail(SELDQ_ERR_UNSUPPORTED, "too many tiles");
  p->total_tiles = (int)tiles;
  // out-component groups: as few as TMEM allows, more while that helps to fill the SMs.  A component costs NBp
  // accumulator columns, or 2 * NB8 when pairs are fused (two column sets per pair); fused groups hold whole pairs.
  p->fuse = w.fuse; p->pair_xor = w.pair_xor; p->NB8 = w.NB8; p->NBmma = w.NBmma;
  const int F = w.fuse ? w.fuse : 1, S = w.fuse ? w.nsets : 1;
  const int cols_per_comp = w.fuse ? S * w.NB8 : w.NBp;                // a fused set of F components: S * F * NB8
  const int max_groups = p->ncomp_out / F;
  int ngroups = 1;
  while (p->ncomp_out / ngroups * cols_per_comp > 512 && ngroups < max_groups) ngroups *= 2;
  if (p->ncomp_out / ngroups * cols_per_comp > 512)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most 512 out channels per component");
  while (ngroups < max_groups && tiles * ngroups * 2 <= n_sms + n_sms / 16) ngroups *= 2;
  // SELDQ_ACC_DOUBLE=1: split once more if that lets the accumulators double-buffer (the epilogue of a unit then
  // overlaps the MMAs of the next; costs re-loading the shared channel chunks per group)
  // sibling launches take that split by default: their CTAs run two or more units each, so an exposed epilogue
  // would be paid per unit
  if (((getenv("SELDQ_ACC_DOUBLE") && atoi(getenv("SELDQ_ACC_DOUBLE")) != 0) ||
       (nprob == 2 && !(getenv("SELDQ_ACC_DOUBLE") && atoi(getenv("SELDQ_ACC_DOUBLE")) == 0))) && ngroups < max_groups &&
      p->ncomp_out / ngroups * cols_per_comp * 2 > 512 && p->ncomp_out / (ngroups * 2) * cols_per_comp * 2 <= 512 &&
      tiles * ngroups >= n_sms)
    ngroups *= 2;
  p->ngroups = ngroups;
  p->gc = p->ncomp_out / ngroups;
  p->mma_per_slab = p->gc / F;
  p->acc_cols = p->gc * cols_per_comp;
  p->total_units = (int)(tiles * ngroups);
  // members of the groups.  Unfused: consecutive components.  Fused: consecutive PAIRS {a0, a0 ^ pair_xor} in the
  // order of their lower members; local component 2 q + t is member t of the group's pair q.
  int8_t set_of[8][8];
  if (w.fuse) {
    if (fusion_sets(g, w.pair_xor, S, set_of) != S) return fail(SELDQ_ERR_INVALID, "fusion: inconsistent grouping");
    int lower[4], nl = 0;
    for (int a = 0; a < p->ncomp_out; ++a)
      if (!(a & w.pair_xor)) lower[nl++] = a;
    const int spg = p->gc / F;                       // fused sets per group; local component F q + t = member t of set q
    for (int gi = 0; gi < ngroups; ++gi)
      for (int q = 0; q < spg; ++q)
        for (int t = 0, sub = 0; t < F; ++t, sub = (sub - w.pair_xor) & w.pair_xor)
          p->comp_of[gi][F * q + t] = (int8_t)(lower[gi * spg + q] | sub);
  } else {
    for (int gi = 0; gi < ngroups; ++gi)
      for (int al = 0; al < p->gc; ++al) p->comp_of[gi][al] = (int8_t)(gi * p->gc + al);
  }
  // cost (number of non-zero blocks) and needed chunks per group
  int cost[8];
  for (int gi = 0; gi < ngroups; ++gi) {
    cost[gi] = 0;
    uint32_t mask = 0;
    for (int c = 0; c < p->chunks; ++c)
      for (int s = 0; s < p->slabs_per_chunk; ++s) {
        const int b = (c * l.BK + s * 16) / l.cpad;
        for (int al = 0; al < p->gc; ++al)
          if (p->op_img[b][p->comp_of[gi][al]] >= 0) { mask |= 1u << c; ++cost[gi]; }
      }
    p->chunk_mask[gi] = mask;
    p->group_order[gi] = gi;
  }
  for (int i = 1; i < ngroups; ++i)                  // insertion sort, heaviest first
    for (int j = i; j > 0 && cost[p->group_order[j]] > cost[p->group_order[j - 1]]; --j) {
      const int t = p->group_order[j]; p->group_order[j] = p->group_order[j - 1]; p->group_order[j - 1] = t;
    }

  // MMA op table (conv_cl.h): per (group, chunk) the valid MMAs in (slab, component | pair) order, zero-padded to
  // slabs_per_chunk * mma_per_slab entries.  `first` = the first MMA of a unit into its accumulator columns
  // (chunks outside the group's mask hold no valid entry for its components by construction).
  const int lps = p->slabs_per_chunk * p->mma_per_slab;
  p->op_entries = ngroups * p->chunks * lps;
  if (p->op_entries > kOpTableEntries)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: too many (slab, component) pairs for the MMA op table");
  if (lps > 32) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: more than 32 MMAs per stage");
  if (p->acc_cols > 512) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: accumulators do not fit in tensor memory");
  // Narrow K side (first CNN layer: one 16-channel slab per tap): a stage per tap is a 4 KB box and one MMA, so the
  // producer's and the issuers' per-stage bookkeeping (~0.3 us) would bound the layer; all taps share one stage then.
  if (l.BK <= 16 && ntaps * lps <= 32 && !getenv("SELDQ_NO_TPS")) p->tps = ntaps;
  p->stage_bytes = p->box_bytes * (uint32_t)p->tps;
  p->box_rows = kTileM;
  // Row-shared taps (conv_cl.h): the KW taps of a kernel row from one box, if their shifts span at most 128 rows (a TMA
  // box has at most 256): every dilation of the TCN qualifies -- a launch's time grows by ~1.4 ns per 128-byte row it
  // loads (ncu, per-launch list), so 128 + 2 d rows instead of 3 x 128 is worth 1-2 us of a 12-14 us launch
  // (a stage's MMAs are issued by the lanes of ONE warp: at most 32 per stage)
  if (p->tps == 1 && l.BK == 64 && g.KW > 1 && (g.KW - 1) * g.dw <= 128 && g.KW * lps <= 32 &&
      !getenv("SELDQ_NO_RS")) {
    int mn = p->off_w[0];
    for (int t = 0; t < g.KW; ++t) mn = p->off_w[t] < mn ? p->off_w[t] : mn;
    p->rs = 1; p->tps = g.KW; p->rs_min_off = mn;
    for (int t = 0; t < g.KW; ++t) p->rs_row[t] = p->off_w[t] - mn;        // the same for every kernel row
    p->box_rows = kTileM + (g.KW - 1) * g.dw;
    p->stage_bytes = (uint32_t)round_up(p->box_rows * l.BK * 2, 1024);
  }
  p->stage_tx = p->rs ? (uint32_t)(p->box_rows * l.BK * 2) : p->stage_bytes;
  p->tap_stride16 = (uint32_t)(((size_t)p->J * p->slab_bytes) >> 4);
  {
    const uint32_t idesc = ptx::make_idesc_bf16(kTileM, (uint32_t)p->NBmma, 0, 0, 0, 0);
    for (int gi = 0; gi < ngroups; ++gi) {
      bool seen[8][4];
      memset(seen, 0, sizeof(seen));
      for (int al = 0; al < 8; ++al)
        for (int st = 0; st < 4; ++st) { p->epi_col[gi][al][st] = 0; p->epi_sgn[gi][al][st] = 0; }
      for (int c = 0; c < p->chunks; ++c) {
        uint2* dst = p->op_tbl + ((size_t)gi * p->chunks + c) * lps;
        int n = 0;
        for (int s = 0; s < p->slabs_per_chunk; ++s)
          for (int ml = 0; ml < p->mma_per_slab; ++ml) {
            const int ch0 = (c * p->slabs_per_chunk + s) * 16;
            const int b = ch0 / l.cpad;
            const int j = (ch0 - b * l.cpad) >> 4;
            uint32_t col, tile16, neg;
            int fi, fs;                                       // which `seen` flag this MMA initialises
            if (!w.fuse) {
              const int a = p->comp_of[gi][ml];
              const int img = p->op_img[b][a];
              if (img < 0) continue;
              col = (uint32_t)(ml * p->NBp);
              tile16 = (uint32_t)(((size_t)img * p->img_bytes + (size_t)j * p->slab_bytes) >> 4);
              neg = (uint32_t)p->op_neg[b][a];
              fi = ml; fs = 0;
            } else {
              const int a0 = p->comp_of[gi][F * ml], m = w.pair_xor;
              const int e0 = p->op_img[b][a0];
              if (e0 < 0) continue;
              const int st = set_of[a0][b];
              col = (uint32_t)((ml * S + st) * p->NBmma);
              tile16 = (uint32_t)(((size_t)fuse_set(e0, m) * p->img_bytes + (size_t)j * p->slab_bytes) >> 4);
              neg = (uint32_t)p->op_neg[b][a0];                // a0's product enters with sign +
              fi = ml; fs = st;
              // epilogue: member t reads the 8-column slot of ITS image in this column set, with its sign relative to a0's
              for (int t = 0; t < F; ++t) {
                const int at = p->comp_of[gi][F * ml + t];
                p->epi_col[gi][F * ml + t][st] = (uint16_t)(col + fuse_slot(p->op_img[b][at], m) * 8);
                p->epi_sgn[gi][F * ml + t][st] = (int8_t)((p->op_neg[b][a0] ^ p->op_neg[b][at]) ? -1 : 1);
              }
            }
            if (col > 0x1ffu || tile16 > 0x3fffu) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path: op table field overflow");
            uint2 e;
            e.x = (1u << 31) | (seen[fi][fs] ? 0u : (1u << 30)) | (col << 20) | ((uint32_t)s << 16) | tile16;
            e.y = idesc | (neg << 14);
            seen[fi][fs] = true;
            dst[n++] = e;
          }
        if (n > 0) dst[n - 1].x |= 1u << 29;                 // the lane that issues last commits the stage
        for (; n < lps; ++n) dst[n] = make_uint2(0u, 0u);
      }
    }
  }
  const int acc_cols = p->acc_cols;
  p->acc_stages = acc_cols * 2 <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < acc_cols * p->acc_stages) cols <<= 1;
  p->tmem_cols = cols;
  const size_t fixed = 1024 /* barriers */ + w.total + 4096 /* slack behind the tiles */;
  const size_t budget = 218 * 1024;   // + 8 KB static (op table, output staging) stays under the 227 KB limit
  if (fixed + 2 * (size_t)p->stage_bytes > budget)
    return fail(SELDQ_ERR_UNSUPPORTED, "compact weights (%zu B as bf16 tiles) do not fit in shared memory", w.total);
  size_t ns = (budget - fixed) / p->stage_bytes;
  if (ns > (size_t)kMaxStages) ns = kMaxStages;
  p->nstages = (int)ns;
  *smem_bytes = ns * p->stage_bytes + fixed;
  return SELDQ_OK;
}


// weight-gradient kernel (wgrad_cl.cu): tile decode, tap groups, split-K ranges, the (a, b) -> compact-tensor fold table
inline int plan_wgrad(const ConvGeom& g, int nprob, int n_sms, WgradParams* pp, size_t* smem_bytes) {
  if (g.sh != 1 || g.sw != 1) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 tensor-core path implements stride 1 only");
  const int ntaps = g.KH * g.KW;
  if (ntaps > kMaxTaps) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most %d taps", kMaxTaps);
  const int nc = g.tab.nc;
  const OperandLayout lx = operand_layout(g.tab.nc, g.tab.nc * g.Ic, g.tab.nc == 1 || g.Ic < 8);      // = x_operand_layout
  if (lx.nc != nc || lx.Cp % 64) return fail(SELDQ_ERR_INVALID, "channels-last wgrad needs a component-padded x operand");
  if (lx.Cp > 512) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 wgrad supports at most 512 padded input channels, got %d", lx.Cp);
  WgradParams& p = *pp;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.nprob = nprob;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    p.off_h[t] = (t / g.KW) * g.dh - g.ph;
    p.off_w[t] = (t % g.KW) * g.dw - g.pw;
  }
  p.OH = g.OH; p.OW = g.OW; p.N = g.N;
  p.ncomp = nc;
  p.OS = 128 / nc;
  p.o_tiles = (g.Oc + p.OS - 1) / p.OS;
  p.cpad_in = lx.cpad; p.Cp = lx.Cp; p.nchunks = lx.Cp / 64;
  for (int a = 0; a < nc; ++a)
    for (int b = 0; b < nc; ++b) {
      const int e = g.tab.widx[a][b];
      if (e < 0) continue;
      const int k = p.pair_n[e]++;
      p.pair_a[e][k] = (int8_t)a; p.pair_b[e][k] = (int8_t)b; p.pair_neg[e][k] = (int8_t)(g.tab.sign[a][b] < 0);
    }
  p.taps_per_group = 256 / lx.Cp;
  if (p.taps_per_group < 1) p.taps_per_group = 1;
  if (p.taps_per_group > ntaps) p.taps_per_group = ntaps;
  p.tap_groups = (ntaps + p.taps_per_group - 1) / p.taps_per_group;
  int cols = 32;
  while (cols < p.taps_per_group * lx.Cp) cols <<= 1;
  p.tmem_cols = cols;
  p.chunks_w = (g.OW + 63) / 64;
  p.ksteps = (long long)g.N * g.OH * p.chunks_w;
  const int tiles = p.nprob * p.tap_groups * p.o_tiles;
  // split-K over positions: one wave of CTAs (every split pays a TMEM round trip and one atomicAdd per compact
  // element in its epilogue), at least 4 K steps each, and no empty splits
  long long splits = n_sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > (p.ksteps + 3) / 4) splits = (p.ksteps + 3) / 4;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  if (const char* e = getenv("SELDQ_WGRAD_SPLITS")) {      // tuning knob (tools/kprof.py)
    const long long v = atoll(e);
    if (v >= 1 && v <= p.ksteps) splits = v;
  }
  const long long per = (p.ksteps + splits - 1) / splits;
  splits = (p.ksteps + per - 1) / per;
  p.splits = (int)splits;
  p.b_tap_bytes = (uint32_t)lx.Cp * 128u;
  p.stage_bytes = 128u * 128u + (uint32_t)p.taps_per_group * p.b_tap_bytes;
  size_t ns = (208 * 1024) / p.stage_bytes;
  if (ns > (size_t)kWgradStages) ns = kWgradStages;
  if (ns < 2) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 wgrad: operand stage of %u B does not fit twice", p.stage_bytes);
  p.nstages = (int)ns;
  size_t smem = ns * p.stage_bytes;
  const size_t stg = (size_t)128 * (nc * 16 + 1) * 4;
  if (smem < stg) smem = stg;
  smem += 4096;   // slack: the tensor core may fetch past the logical end of the last operand tile

  *smem_bytes = smem;
  return SELDQ_OK;
}

}  // namespace cl
}  // namespace seldq

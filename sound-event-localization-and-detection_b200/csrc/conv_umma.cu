// tcgen05 implicit-GEMM convolution for the quaternion / dual-quaternion layers: forward and dgrad.
//
//   D[t, (a,o)] = sum_b sum_tap sum_i  sign[a][b] * X_b[t + off(tap), i] * W_{widx[a][b]}[o, i, tap]
//
//   * M = 128 consecutive time/width positions of one (n, h) row  -> TMEM lanes
//   * N = out channels of ONE out component a (padded to 16)      -> TMEM columns a*NBp ...
//   * K = (in component b, tap, 8-channel atom)                   -> streamed through a TMA ring
//
// The Hamilton / dual-quaternion expansion (quaternion_ops.py:131-135, dual_quaternion_ops.py:122-140)
// is never materialised: the COMPACT weights are converted to bf16 once per CTA into shared memory
// (UMMA K-major, no swizzle) and stay resident; every (a,b) block is one tcgen05.mma whose
// instruction descriptor carries the block's sign in the negate-B bit; the dual-quaternion zero
// block is simply never issued (25 % fewer MMAs).  Activations are channels-first with time
// contiguous, so the A operand is MN-major: TMA boxes [64 t x Rc channels] land in the
// 128B-swizzled canonical layout as they are; dilation / taps / padding are TMA coordinate offsets
// with out-of-bounds zero fill.
//
// Layers whose channel count per component is not a multiple of 8 but whose whole expanded weight
// is small (the first CNN layer, Cin = 8 or 16) run in "dense" mode: the signed expanded tile is
// built in shared memory only and one MMA covers all out channels.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> coalesced global stores, + bias, + optional bf16 copy).
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "launch.h"
#include "tensor_map.h"
#include "umma_ptx.cuh"

namespace seldq {
namespace umma {

__device__ __forceinline__ float fprop_weight(const FpropParams& p, int img, int row, int tap, int chan) {
  if (p.dense) return row < p.g.P ? expanded_weight(p.g, p.w, row, chan, tap) : 0.f;
  if (row >= p.Pc) return 0.f;
  const int o = p.g.transposed ? chan : row, i = p.g.transposed ? row : chan;
  return p.w[img][(long long)o * p.g.wsO + (long long)i * p.g.wsI + (long long)tap * p.g.wsT];
}

__global__ void __launch_bounds__(kThreads, 1)
qconv_umma_fprop_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t stage_bytes = 2u * p.stage_atoms * 1024u;
  const uint32_t half_bytes = p.stage_atoms * 1024u;
  const uint32_t op_bytes = (uint32_t)p.NBp * 32u;            // one [NBp x 16] bf16 B operand
  // layout: [activation ring][barriers, 1 KB][weight images][slack]; the tensor core may fetch past the
  // logical end of an operand tile (seen on the bring-up probe), so the images are not the last bytes
  uint8_t* a_ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * stage_bytes);
  uint8_t* b_img = smem + (size_t)p.nstages * stage_bytes + 1024;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  // ---- one-time setup ---------------------------------------------------------------------------
  // (1) zero the activation ring: pad atoms are never written by TMA and meet zero weights, so
  //     whatever they hold only has to be finite
  for (uint32_t i = threadIdx.x; i < (uint32_t)p.nstages * stage_bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(a_ring)[i] = make_uint4(0, 0, 0, 0);
  // (2) compact weights -> bf16 operand tiles.  Tile (img, kpair) holds B[n][k], n = out channel of
  //     the component, k = 16 consecutive entries of the component's K sequence; layout = UMMA
  //     K-major / no swizzle: core matrix = 8 rows x 16 B, LBO (K step) = NBp*16, SBO (8 rows) = 128
  {
    const int rows8 = p.NBp / 8;
    const int items = p.n_img * p.kpairs * 2 * p.NBp;   // one item = 8 consecutive k of one row
    for (int it = threadIdx.x; it < items; it += kThreads) {
      int r = it;
      const int n = r % p.NBp; r /= p.NBp;
      const int kc = r & 1; r >>= 1;
      const int kp = r % p.kpairs;
      const int img = r / p.kpairs;
      const int atom = kp * 2 + kc;
      const int tap = p.atom_tap[atom];
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = __float2bfloat16_rn(tap >= 0 ? fprop_weight(p, img, n, tap, p.atom_chan[atom] + j) : 0.f);
      uint8_t* dst = b_img + (size_t)(img * p.kpairs + kp) * op_bytes + (size_t)kc * (rows8 * 128) + (n >> 3) * 128 +
                     (n & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
    }
  }
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 4); }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tm_in);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = (uint32_t)(p.ncomp_out * p.NBp);

  if (warp == 0) {
    // ===== TMA producer ============================================================================
    if (ptx::elect_one()) {
      uint32_t slot = 0, parity = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int r = tile;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int h = r % p.OH;
        const int n = r / p.OH;
        const int w0 = wt * kTileM;
        for (int b = 0; b < p.ncomp_in; ++b)
          for (int s = 0; s < p.stages_per_comp; ++s) {
            const int t0 = s * p.G;
            const int nt = min(p.G, p.ntaps - t0);
            ptx::mbar_wait(&empty_bar[slot], parity ^ 1);
            if (p.debug & 2) {
              ptx::mbar_arrive(&full_bar[slot]);
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[slot], 2u * nt * p.Rc * 128u);
              uint8_t* st = a_ring + (size_t)slot * stage_bytes;
              // off_w already contains the tap's mirror shift, so the inner coordinate is a multiple of 8
              for (int mh = 0; mh < 2; ++mh)
                for (int ti = 0; ti < nt; ++ti)
                  ptx::tma_load_5d(st + mh * half_bytes + (size_t)ti * p.Rc * 128, &tm_in, &full_bar[slot],
                                   w0 + mh * 64 + p.off_w[t0 + ti], h + p.off_h[t0 + ti], b * p.Rc, n,
                                   p.tap_sidx[t0 + ti]);
            }
            if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ==============================================================================
    if (ptx::elect_one()) {
      const uint64_t a_hi = ptx::make_smem_desc_hi(half_bytes, 1024, ptx::kSwizzle128B);   // MN-major: LBO = M-half stride
      const uint64_t b_hi = ptx::make_smem_desc_hi((uint32_t)p.NBp * 16u, 128, ptx::kSwizzleNone);
      const uint32_t idesc = ptx::make_idesc_bf16(kTileM, (uint32_t)p.NBp, /*A MN-major*/ 1, /*B K-major*/ 0, 0, 0);
      const uint32_t a_base = ptx::smem_u32(a_ring), b_base = ptx::smem_u32(b_img);
      uint32_t slot = 0, parity = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t use = p.acc_stages == 2 ? (it >> 1) : it;
        ptx::mbar_wait(&tempty_bar[as], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_base = tmem_base + as * acc_cols;
        uint32_t written = 0;
        for (int b = 0; b < p.ncomp_in; ++b)
          for (int s = 0; s < p.stages_per_comp; ++s) {
            const int nt = min(p.G, p.ntaps - s * p.G);
            const int npairs = (nt * p.atoms_per_tap + 1) >> 1;
            ptx::mbar_wait(&full_bar[slot], parity);
            ptx::tc_fence_after();
            const uint32_t a_stage = a_base + slot * stage_bytes;
            for (int j = 0; j < npairs; ++j) {
              const uint64_t a_desc = ptx::smem_desc(a_hi, a_stage + (uint32_t)j * 2048u);
              const int kp = s * (p.stage_atoms >> 1) + j;
              for (int o = 0; o < p.nops[b]; ++o) {
                const int oc = p.op_out[b][o];
                const uint64_t b_desc =
                    ptx::smem_desc(b_hi, b_base + (uint32_t)(p.op_img[b][o] * p.kpairs + kp) * op_bytes);
                if (!(p.debug & 1))
                  ptx::umma_f16(d_base + (uint32_t)(oc * p.NBp), a_desc, b_desc,
                                idesc | ((uint32_t)p.op_neg[b][o] << 14), (written >> oc) & 1u);
                written |= 1u << oc;
              }
            }
            if (p.debug & 1) ptx::mbar_arrive(&empty_bar[slot]);
            else ptx::umma_commit(&empty_bar[slot]);    // frees the ring slot once these MMAs retire
            if (++slot == (uint32_t)p.nstages) { slot = 0; parity ^= 1; }
          }
        if (p.debug & 1) ptx::mbar_arrive(&tfull_bar[as]);
        else ptx::umma_commit(&tfull_bar[as]);     // accumulator complete -> epilogue
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global ===================================================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t use = p.acc_stages == 2 ? (it >> 1) : it;
      int r = tile;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int h = r % p.OH;
      const int n = r / p.OH;
      const int w = wt * kTileM + row;
      const bool w_ok = w < p.OW;
      ptx::mbar_wait(&tfull_bar[as], use & 1);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols;
      float* out_row = p.out + (long long)n * p.out_sN + (long long)h * p.out_sH + w;
      __nv_bfloat16* out16_row =
          p.out_bf16 ? p.out_bf16 + (long long)n * p.o16_sN + (long long)h * p.o16_sH + w : nullptr;
      for (int a = 0; a < p.ncomp_out; ++a)
        for (int c0 = 0; c0 < p.Pc; c0 += 16) {
          uint32_t v[16];
          if (p.debug & 4) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0;
          } else {
            ptx::tmem_ld16(t_row + (uint32_t)(a * p.NBp + c0), v);
            ptx::tmem_ld_wait();
          }
          if (w_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c < p.Pc) {
                const int ch = a * p.Pc + c;
                float y = __uint_as_float(v[j]);
                if (p.bias) y += __ldg(p.bias + ch);
                out_row[(long long)ch * p.out_sC] = y;
                if (out16_row) out16_row[(long long)ch * p.o16_sC] = __float2bfloat16_rn(y);
              }
            }
          }
        }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace umma

// ---- host side: plan + launch ---------------------------------------------------------------------
using umma::FpropParams;

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

int plan_umma_fprop(const ConvGeom& g, FpropParams* p, size_t* smem_bytes) {
  memset(p, 0, sizeof(*p));
  if (g.sh != 1 || g.sw != 1)
    return fail(SELDQ_ERR_UNSUPPORTED, "bf16 tensor-core path implements stride 1 only (got %dx%d)", g.sh, g.sw);
  const int ntaps = g.KH * g.KW;
  if (ntaps > umma::kMaxTaps) return fail(SELDQ_ERR_UNSUPPORTED, "bf16 path supports at most %d taps", umma::kMaxTaps);
  const int nc = g.tab.nc;
  const int pc = g.transposed ? g.Ic : g.Oc, rc = g.transposed ? g.Oc : g.Ic;
  p->g = g;
  p->ntaps = ntaps;
  p->N = g.N; p->OH = g.OH; p->OW = g.OW; p->P = g.P;
  for (int t = 0; t < ntaps; ++t) {
    const int kh = t / g.KW, kw = t % g.KW;
    // forward: in = out - pad + k*dil ; dgrad: gy position = gx position + pad - k*dil   (stride 1)
    p->off_h[t] = g.transposed ? g.ph - kh * g.dh : kh * g.dh - g.ph;
    p->off_w[t] = g.transposed ? g.pw - kw * g.dw : kw * g.dw - g.pw;
  }
  if (rc % 8 == 0 && rc <= 64 && nc * round_up(pc, 16) <= 512) {
    p->dense = 0;
    p->ncomp_in = p->ncomp_out = nc;
    p->Rc = rc; p->Pc = pc; p->NBp = round_up(pc, 16);
    p->n_img = g.tab.nw;
    for (int b = 0; b < nc; ++b) {
      int k = 0;
      for (int a = 0; a < nc; ++a) {
        const int fa = g.transposed ? b : a, fb = g.transposed ? a : b;   // forward-sense (out, in) components
        const int e = g.tab.widx[fa][fb];
        if (e < 0) continue;
        p->op_img[b][k] = (int8_t)e;
        p->op_neg[b][k] = (int8_t)(g.tab.sign[fa][fb] < 0);
        p->op_out[b][k] = (int8_t)a;
        ++k;
      }
      p->nops[b] = k;
    }
  } else if (g.R % 8 == 0 && g.R <= 64 && g.P <= 256) {
    p->dense = 1;
    p->ncomp_in = p->ncomp_out = 1;
    p->Rc = g.R; p->Pc = g.P; p->NBp = round_up(g.P, 16);
    p->n_img = 1;
    p->nops[0] = 1;
  } else {
    return fail(SELDQ_ERR_UNSUPPORTED,
                "bf16 tensor-core path needs channels per component to be a multiple of 8 (<= 64), or a layer with "
                "<= 64 input and <= 256 output channels; got %d -> %d channels in %d components",
                g.R, g.P, nc);
  }
  p->atoms_per_tap = p->Rc / 8;
  p->G = (6 + p->atoms_per_tap - 1) / p->atoms_per_tap;
  if (p->G > ntaps) p->G = ntaps;
  p->stages_per_comp = (ntaps + p->G - 1) / p->G;
  p->stage_atoms = round_up(p->G * p->atoms_per_tap, 2);
  p->kpairs = p->stages_per_comp * (p->stage_atoms / 2);
  if (p->kpairs * 2 > umma::kMaxAtoms) return fail(SELDQ_ERR_UNSUPPORTED, "K sequence too long for the bf16 path");
  for (int a = 0; a < p->kpairs * 2; ++a) {
    const int s = a / p->stage_atoms, la = a % p->stage_atoms;
    const int nt = (ntaps - s * p->G) < p->G ? (ntaps - s * p->G) : p->G;
    if (la < nt * p->atoms_per_tap) {
      p->atom_tap[a] = (int8_t)(s * p->G + la / p->atoms_per_tap);
      p->atom_chan[a] = (int8_t)((la % p->atoms_per_tap) * 8);
    } else {
      p->atom_tap[a] = -1;
      p->atom_chan[a] = 0;
    }
  }
  const int acc_cols = p->ncomp_out * p->NBp;
  p->acc_stages = acc_cols * 2 <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < acc_cols * p->acc_stages) cols <<= 1;
  p->tmem_cols = cols;
  p->tiles_w = (g.OW + umma::kTileM - 1) / umma::kTileM;
  const long long tiles = (long long)g.N * g.OH * p->tiles_w;
  if (tiles > 0x7fffffffLL) return fail(SELDQ_ERR_UNSUPPORTED, "too many tiles");
  p->total_tiles = (int)tiles;
  const size_t img_bytes = (size_t)p->n_img * p->kpairs * p->NBp * 32;
  const size_t fixed = 1024 /* barriers */ + img_bytes + 8192 /* slack behind the images */;
  const size_t stage_bytes = 2u * p->stage_atoms * 1024u;
  const size_t budget = 225 * 1024;
  if (fixed + 2 * stage_bytes > budget)
    return fail(SELDQ_ERR_UNSUPPORTED, "compact weights (%zu B as bf16) do not fit in shared memory", img_bytes);
  size_t ns = (budget - fixed) / stage_bytes;
  if (ns > (size_t)umma::kMaxStages) ns = umma::kMaxStages;
  p->nstages = (int)ns;
  *smem_bytes = ns * stage_bytes + fixed;
  return SELDQ_OK;
}

// shifts[0] = 0; then one entry per distinct (-off_w) mod 8 over the taps of the pass
void mirror_shifts(const ConvGeom& fwd_geom, int which, int* shifts, int* nshifts) {
  int n = 0;
  shifts[n++] = 0;
  for (int kw = 0; kw < fwd_geom.KW; ++kw) {
    const int off = which == 0 ? kw * fwd_geom.dw - fwd_geom.pw : fwd_geom.pw - kw * fwd_geom.dw;
    const int s = ((-off) % 8 + 8) % 8;
    bool seen = false;
    for (int i = 0; i < n; ++i) seen |= shifts[i] == s;
    if (!seen && n < 8) shifts[n++] = s;
  }
  *nshifts = n;
}

// fills tap_sidx / folds the shift into off_w; false if the mirror set lacks a needed shift
static bool bind_mirror(const MirrorSet& m, int ntaps, const int* off_w_in, int* off_w_out, int* tap_sidx) {
  for (int t = 0; t < ntaps; ++t) {
    const int s = ((-off_w_in[t]) % 8 + 8) % 8;
    int idx = -1;
    for (int i = 0; i < m.nshifts; ++i)
      if (m.shifts[i] == s) idx = i;
    if (idx < 0) return false;
    tap_sidx[t] = idx;
    off_w_out[t] = off_w_in[t] + s;
  }
  return true;
}

// (W, H, C, N, shift) view of a mirror set with box {64, 1, rows, 1, 1}, 128-byte swizzle
int encode_mirror_map(CUtensorMap* tm, const MirrorSet& m, int w, int h, int c, int n, int box_rows) {
  const uint64_t pitch = (uint64_t)mirror_pitch(w);
  const uint64_t dims[5] = {pitch, (uint64_t)h, (uint64_t)c, (uint64_t)n, (uint64_t)m.nshifts};
  const uint64_t strides[4] = {pitch * 2, pitch * h * 2, pitch * h * c * 2, pitch * h * c * n * 2};
  const uint32_t box[5] = {64, 1, (uint32_t)box_rows, 1, 1};
  return encode_tensor_map(tm, m.data, 2, 5, dims, strides, box, /*128B*/ 3);
}

int launch_umma_fprop(const ConvGeom& g, const MirrorSet& in, const float* const* host_w, const float* bias,
                      float* out, cudaStream_t st) {
  FpropParams p;
  size_t smem = 0;
  int rc = plan_umma_fprop(g, &p, &smem);
  if (rc) return rc;
  for (int i = 0; i < g.tab.nw; ++i) p.w[i] = host_w[i];
  if (const char* dbg = getenv("SELDQ_DEBUG")) p.debug = atoi(dbg);
  p.bias = bias;
  p.out = out;
  p.out_sN = g.out_sN; p.out_sC = g.out_sC; p.out_sH = g.out_sH;
  p.out_bf16 = nullptr;
  int off_w[umma::kMaxTaps];
  for (int t = 0; t < p.ntaps; ++t) off_w[t] = p.off_w[t];
  if (!bind_mirror(in, p.ntaps, off_w, p.off_w, p.tap_sidx))
    return fail(SELDQ_ERR_INVALID, "bf16 mirror set lacks a shift this convolution's taps need");
  alignas(64) CUtensorMap tm;
  rc = encode_mirror_map(&tm, in, g.IW, g.IH, g.R, g.N, p.Rc);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(umma::qconv_umma_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "fprop smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  const int grid = p.total_tiles < umma::num_sms() ? p.total_tiles : umma::num_sms();
  umma::qconv_umma_fprop_kernel<<<grid, umma::kThreads, smem, st>>>(tm, p);
  return check_launch("qconv_umma_fprop_kernel");
}

}  // namespace seldq

// Frame-pair form of the STFT magnitude kernel (stft.cuh has the per-frame form and the reference citations).
//
// stft.cuh is bound by its total instruction issue (ALU + shared-memory instructions: ~800 warp instructions per frame
// at 54 % issue utilisation, ncu), and packing (re, im) of one frame into the f32x2 instructions of sm_100 did not help
// (a complex multiply crosses the halves).  Here the two halves of every packed register hold the SAME quantity of two
// CONSECUTIVE FRAMES f, f + 1 of one signal: the two frames run the identical instruction stream with identical window
// values and twiddles, FMUL2 / FFMA2 take the twiddle as a broadcast scalar operand (`R.F32`), FADD2 negates operands
// for free -- so every butterfly, twiddle and untangle instruction, every exchange through shared memory (16-byte
// accesses {re_f, re_f+1, im_f, im_f+1}), every table load and every staging store (8 bytes: two adjacent frames of a
// bin) serves two frames.  16 threads own a frame pair; the pipeline is stft.cuh's: DFT16 over r, twiddle, exchange,
// DFT16 over n2, exchange, real-input untangle, |.|, [bin][frame] staging tile, row stores.  The factor 1/2 of the
// untangle step is folded into the window table.
//
// Magnitude (PLANES = 1) or magnitude + phase (PLANES = 2), float32 or int16 input, optional normalisation; the
// statistics pass stays with stft.cuh (launch_stft picks).  PAIRS frame pairs per batch (block = 16 * PAIRS threads).
#pragma once
#include "stft.cuh"

#if defined(__CUDA_ARCH__)
#include <cuda_runtime.h>
#endif

namespace seldq {
namespace stft2 {

using stft::MAXBINS;
using stft::NFFT;
using stft::Params;

// two frames' values of one quantity
struct P2 {
  float a, b;
};
SELDQ_HD P2 mk(float a, float b) { P2 r; r.a = a; r.b = b; return r; }
#if defined(__CUDA_ARCH__)
SELDQ_HD P2 from2(float2 v) { return mk(v.x, v.y); }
SELDQ_HD float2 to2(P2 v) { return make_float2(v.a, v.b); }
SELDQ_HD P2 add(P2 x, P2 y) { return from2(__fadd2_rn(to2(x), to2(y))); }
SELDQ_HD P2 sub(P2 x, P2 y) { return from2(__fadd2_rn(to2(x), make_float2(-y.a, -y.b))); }
SELDQ_HD P2 mul_s(P2 x, float s) { return from2(__fmul2_rn(to2(x), make_float2(s, s))); }
SELDQ_HD P2 fma_s(P2 x, float s, P2 c) { return from2(__ffma2_rn(to2(x), make_float2(s, s), to2(c))); }       // x s + c
SELDQ_HD P2 nfma_s(P2 x, float s, P2 c) { return from2(__ffma2_rn(make_float2(-x.a, -x.b), make_float2(s, s), to2(c))); }  // c - x s
SELDQ_HD P2 mul(P2 x, P2 y) { return from2(__fmul2_rn(to2(x), to2(y))); }
SELDQ_HD P2 fma(P2 x, P2 y, P2 c) { return from2(__ffma2_rn(to2(x), to2(y), to2(c))); }
#else
SELDQ_HD P2 add(P2 x, P2 y) { return mk(x.a + y.a, x.b + y.b); }
SELDQ_HD P2 sub(P2 x, P2 y) { return mk(x.a - y.a, x.b - y.b); }
SELDQ_HD P2 mul_s(P2 x, float s) { return mk(x.a * s, x.b * s); }
SELDQ_HD P2 fma_s(P2 x, float s, P2 c) { return mk(fmaf(x.a, s, c.a), fmaf(x.b, s, c.b)); }
SELDQ_HD P2 nfma_s(P2 x, float s, P2 c) { return mk(fmaf(-x.a, s, c.a), fmaf(-x.b, s, c.b)); }
SELDQ_HD P2 mul(P2 x, P2 y) { return mk(x.a * y.a, x.b * y.b); }
SELDQ_HD P2 fma(P2 x, P2 y, P2 c) { return mk(fmaf(x.a, y.a, c.a), fmaf(x.b, y.b, c.b)); }
#endif

// keeps the compiler from hoisting every shared-memory load of an unrolled loop to its top (at 512 threads the kernel
// has 128 registers: the 64 data registers of a frame pair plus the loads of FOUR iterations fit, those of sixteen spill)
#if defined(__CUDA_ARCH__)
#define SELDQ_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define SELDQ_SCHED_FENCE()
#endif

template <int PAIRS, int PLANES = 1>
struct Shared {
  float4 xch[PAIRS][16][17];                 // exchange: one 16 x 16 complex matrix (pitch 17) per frame pair
  float tile[PLANES][MAXBINS][2 * PAIRS + 2];   // staging [plane][bin][frame]; even pitch: a frame pair is one 8-byte store
  float2 tw256[16][16];                      // [k1][j] = W256^(j k1)
  float2 tw512[17][16];                      // [q][j]  = W512^(j + 16 q)
  float2 win[16][16];                        // [r][j]  = HALF the window at samples 2 (j + 16 r), + 1, / sum(w)
};

struct Thread {
  P2 re[16], im[16];
};
struct Raw {                                 // the thread's 16 complex points of both frames, as loaded
  float ax[16], ay[16], bx[16], by[16];
};

template <int PAIRS, int PLANES>
SELDQ_HD void init_tables(Shared<PAIRS, PLANES>& s, int tid) {
  for (int idx = tid; idx < 17 * 16; idx += 16 * PAIRS) {
    const int q = idx >> 4, j = idx & 15;
    float sn, cs;
    stft::sincospi_f((float)(j + 16 * q) / 256.0f, &sn, &cs);
    s.tw512[q][j] = make_float2(cs, -sn);
    if (q < 16) {
      stft::sincospi_f((float)(j * q) / 128.0f, &sn, &cs);
      s.tw256[q][j] = make_float2(cs, -sn);
      float w[2];
      for (int h = 0; h < 2; ++h) {
        stft::sincospi_f((float)(2 * (j + 16 * q) + h) / 256.0f, &sn, &cs);
        w[h] = 0.5f * ((0.54f - 0.46f * cs) * (1.0f / (0.54f * NFFT)));      // 1/2: the untangle step's factor
      }
      s.win[q][j] = make_float2(w[0], w[1]);
    }
  }
}

#define SELDQ_P2_DFT4(r0, i0, r1, i1, r2, i2, r3, i3)                                  \
  {                                                                                     \
    const P2 t0r = add(r0, r2), t0i = add(i0, i2), t1r = sub(r0, r2), t1i = sub(i0, i2); \
    const P2 t2r = add(r1, r3), t2i = add(i1, i3), t3r = sub(i1, i3), t3i = sub(r3, r1); \
    r0 = add(t0r, t2r); i0 = add(t0i, t2i); r1 = add(t1r, t3r); i1 = add(t1i, t3i);       \
    r2 = sub(t0r, t2r); i2 = sub(t0i, t2i); r3 = sub(t1r, t3r); i3 = sub(t1i, t3i);       \
  }

// forward 16-point DFT of both frames, natural order in and out (stft::dft16 on packed values)
SELDQ_HD void dft16(P2 (&re)[16], P2 (&im)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) SELDQ_P2_DFT4(re[n2], im[n2], re[4 + n2], im[4 + n2], re[8 + n2], im[8 + n2], re[12 + n2], im[12 + n2]);
  // y[k1][n2] *= (wr + i wi)
#define SELDQ_P2_CMUL(idx, wr, wi)                          \
  {                                                         \
    const P2 a = re[idx], b = im[idx];                      \
    re[idx] = nfma_s(b, (wi), mul_s(a, (wr)));              \
    im[idx] = fma_s(b, (wr), mul_s(a, (wi)));               \
  }
  SELDQ_P2_CMUL(4 * 1 + 1, c1, -s1)
  SELDQ_P2_CMUL(4 * 1 + 2, h, -h)
  SELDQ_P2_CMUL(4 * 1 + 3, s1, -c1)
  SELDQ_P2_CMUL(4 * 2 + 1, h, -h)
  {                                                          // * (-i)
    const P2 a = re[4 * 2 + 2], b = im[4 * 2 + 2];
    re[4 * 2 + 2] = b; im[4 * 2 + 2] = mul_s(a, -1.f);
  }
  SELDQ_P2_CMUL(4 * 2 + 3, -h, -h)
  SELDQ_P2_CMUL(4 * 3 + 1, s1, -c1)
  SELDQ_P2_CMUL(4 * 3 + 2, -h, -h)
  SELDQ_P2_CMUL(4 * 3 + 3, -c1, s1)
#undef SELDQ_P2_CMUL
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    SELDQ_P2_DFT4(re[4 * k1], im[4 * k1], re[4 * k1 + 1], im[4 * k1 + 1], re[4 * k1 + 2], im[4 * k1 + 2], re[4 * k1 + 3],
                  im[4 * k1 + 3]);
#define SELDQ_P2_SWAP(x, y)                                                  \
  {                                                                          \
    const P2 tr = re[x], ti = im[x];                                         \
    re[x] = re[y]; im[x] = im[y]; re[y] = tr; im[y] = ti;                    \
  }
  SELDQ_P2_SWAP(1, 4) SELDQ_P2_SWAP(2, 8) SELDQ_P2_SWAP(3, 12) SELDQ_P2_SWAP(6, 9) SELDQ_P2_SWAP(7, 13) SELDQ_P2_SWAP(11, 14)
#undef SELDQ_P2_SWAP
}
#undef SELDQ_P2_DFT4

// batch index -> signal, first frame
template <int PAIRS>
SELDQ_HD void batch_decode(const Params& p, long long batch, int* signal, int* t0) {
  const unsigned b32 = (unsigned)batch, g32 = (unsigned)p.groups;      // the launcher keeps total below 2^31
  *signal = (int)(b32 / g32);
  *t0 = (int)(b32 - (unsigned)(*signal) * g32) * (2 * PAIRS);
}

SELDQ_HD float sample_at(const Params& p, int signal, long long g) {
  if (g < 0 || g >= p.n_samples) return 0.f;
  if (p.x16 != nullptr) return (float)p.x16[(long long)signal * p.n_samples + g] * (1.0f / 32768.0f);
  return p.x[(long long)signal * p.n_samples + g];
}

// raw samples of this thread's 16 complex points z[j + 16 r] of frames t0 + 2 fp and t0 + 2 fp + 1
SELDQ_HD void load_raw(const Params& p, Raw& raw, int tid, int signal, int t0) {
  const int fp = tid >> 4, j = tid & 15;
  const long long gA = (long long)(t0 + 2 * fp) * p.hop - NFFT / 2, gB = gA + p.hop;
  const bool interior = gA >= 0 && gB + NFFT <= p.n_samples;
#if defined(__CUDA_ARCH__)
  if (interior && p.x16 == nullptr) {
    const float* src = p.x + (long long)signal * p.n_samples;
    if (((reinterpret_cast<unsigned long long>(src + gA) | reinterpret_cast<unsigned long long>(src + gB)) & 7ull) == 0) {
      const float2* a = reinterpret_cast<const float2*>(src + gA) + j;
      const float2* b = reinterpret_cast<const float2*>(src + gB) + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float2 va = __ldg(a + 16 * r), vb = __ldg(b + 16 * r);
        raw.ax[r] = va.x; raw.ay[r] = va.y; raw.bx[r] = vb.x; raw.by[r] = vb.y;
      }
      return;
    }
  }
  if (interior && p.x16 != nullptr) {
    const short* src = p.x16 + (long long)signal * p.n_samples;
    if (((reinterpret_cast<unsigned long long>(src + gA) | reinterpret_cast<unsigned long long>(src + gB)) & 3ull) == 0) {
      const short2* a = reinterpret_cast<const short2*>(src + gA) + j;
      const short2* b = reinterpret_cast<const short2*>(src + gB) + j;
      const float sc = 1.0f / 32768.0f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const short2 va = __ldg(a + 16 * r), vb = __ldg(b + 16 * r);
        raw.ax[r] = (float)va.x * sc; raw.ay[r] = (float)va.y * sc; raw.bx[r] = (float)vb.x * sc; raw.by[r] = (float)vb.y * sc;
      }
      return;
    }
  }
#endif
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const long long m = 2 * (j + 16 * r);
    raw.ax[r] = sample_at(p, signal, gA + m); raw.ay[r] = sample_at(p, signal, gA + m + 1);
    raw.bx[r] = sample_at(p, signal, gB + m); raw.by[r] = sample_at(p, signal, gB + m + 1);
  }
}

// window, DFT16 over r, twiddle, into the exchange buffer
template <int PAIRS, int PLANES>
SELDQ_HD void phase_a(Shared<PAIRS, PLANES>& s, const Raw& raw, Thread& th, int tid) {
  const int fp = tid >> 4, j = tid & 15;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 w = s.win[r][j];
    th.re[r] = mk(raw.ax[r] * w.x, raw.bx[r] * w.x);
    th.im[r] = mk(raw.ay[r] * w.y, raw.by[r] * w.y);
    if ((r & 3) == 3) SELDQ_SCHED_FENCE();
  }
  dft16(th.re, th.im);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const float2 w = s.tw256[k1][j];
    const P2 xr = nfma_s(th.im[k1], w.y, mul_s(th.re[k1], w.x)), xi = fma_s(th.im[k1], w.x, mul_s(th.re[k1], w.y));
    s.xch[fp][k1][j] = make_float4(xr.a, xr.b, xi.a, xi.b);
    if ((k1 & 3) == 3) SELDQ_SCHED_FENCE();
  }
}

template <int PAIRS, int PLANES>
SELDQ_HD void phase_b(const Shared<PAIRS, PLANES>& s, Thread& th, int tid) {
  const int fp = tid >> 4, j = tid & 15;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const float4 v = s.xch[fp][j][n2];
    th.re[n2] = mk(v.x, v.y);
    th.im[n2] = mk(v.z, v.w);
  }
  dft16(th.re, th.im);          // th[q] = Z[j + 16 q] (half scale)
}

template <int PAIRS, int PLANES>
SELDQ_HD void phase_b2(Shared<PAIRS, PLANES>& s, const Thread& th, int tid) {
  const int fp = tid >> 4, j = tid & 15;
#pragma unroll
  for (int q = 0; q < 16; ++q) s.xch[fp][q][j] = make_float4(th.re[q].a, th.re[q].b, th.im[q].a, th.im[q].b);
}

SELDQ_HD float mag_of(float m2) {
#if defined(__CUDA_ARCH__)
  float mag;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(m2));
  return mag;
#else
  return sqrtf(m2);
#endif
}

// atan2 of both frames: Abramowitz & Stegun 4.4.49 (atan(t) / t as a degree-8 polynomial in t^2 on [0, 1], |error| <=
// 2e-8) on packed registers, then the octant / quadrant reflections per half.  atan2f costs ~35 scalar instructions per
// value; this is ~30 for the pair.  Zero magnitude gives 0 (as np.angle(0)), xi = +0 with xr < 0 gives pi.
SELDQ_HD float refl_(float r, float ay, float ax, float y, float x) {
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.f) r = 3.14159265358979324f - r;
  return y < 0.f ? -r : r;
}
SELDQ_HD P2 atan2_p2(P2 y, P2 x) {
  const float axa = fabsf(x.a), aya = fabsf(y.a), axb = fabsf(x.b), ayb = fabsf(y.b);
  const float mxa = fmaxf(axa, aya), mna = fminf(axa, aya), mxb = fmaxf(axb, ayb), mnb = fminf(axb, ayb);
#if defined(__CUDA_ARCH__)
  const P2 t = mk(mxa > 0.f ? __fdividef(mna, mxa) : 0.f, mxb > 0.f ? __fdividef(mnb, mxb) : 0.f);
#else
  const P2 t = mk(mxa > 0.f ? mna / mxa : 0.f, mxb > 0.f ? mnb / mxb : 0.f);
#endif
  const P2 u = mul(t, t);
  P2 q = mk(0.0028662257f, 0.0028662257f);
  q = fma(q, u, mk(-0.0161657367f, -0.0161657367f));
  q = fma(q, u, mk(0.0429096138f, 0.0429096138f));
  q = fma(q, u, mk(-0.0752896400f, -0.0752896400f));
  q = fma(q, u, mk(0.1065626393f, 0.1065626393f));
  q = fma(q, u, mk(-0.1420889944f, -0.1420889944f));
  q = fma(q, u, mk(0.1999355085f, 0.1999355085f));
  q = fma(q, u, mk(-0.3333314528f, -0.3333314528f));
  q = fma(q, u, mk(1.f, 1.f));
  const P2 r = mul(q, t);
  return mk(refl_(r.a, aya, axa, y.a, x.a), refl_(r.b, ayb, axb, y.b, x.b));
}

template <int PAIRS, int PLANES>
SELDQ_HD void emit_bin(const Params& p, Shared<PAIRS, PLANES>& s, int fp, int kb, P2 xr, P2 xi) {
  if (kb < 0) return;
  const P2 m2 = fma(xi, xi, mul(xr, xr));
  *reinterpret_cast<float2*>(&s.tile[0][kb][2 * fp]) = make_float2(mag_of(m2.a), mag_of(m2.b));
  if (PLANES == 2) {
    const P2 ph = atan2_p2(xi, xr);
    *reinterpret_cast<float2*>(&s.tile[PLANES - 1][kb][2 * fp]) = make_float2(ph.a, ph.b);
  }
}

// real-input untangle (window at half scale): R[k] = E + W512^k O, E = Z[k] + conj Z[256-k], O = (Z[k] - conj Z[256-k]) / i
template <int PAIRS, int PLANES>
SELDQ_HD void phase_c(const Params& p, Shared<PAIRS, PLANES>& s, const Thread& th, int tid) {
  const int fp = tid >> 4, j = tid & 15;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 pz = j == 0 ? s.xch[fp][(16 - q) & 15][0] : s.xch[fp][15 - q][16 - j];
    const P2 a = th.re[q], bb = th.im[q], cc = mk(pz.x, pz.y), d = mk(pz.z, pz.w);
    const P2 er = add(a, cc), ei = sub(bb, d), orr = add(bb, d), oi = sub(cc, a);
    const float2 w = s.tw512[q][j];
    emit_bin(p, s, fp, j + 16 * q - p.bin0, nfma_s(oi, w.y, fma_s(orr, w.x, er)), fma_s(orr, w.y, fma_s(oi, w.x, ei)));
    if ((q & 3) == 3) SELDQ_SCHED_FENCE();
  }
  if (j == 0) {                                                        // Nyquist bin: 2 (Re Z0 - Im Z0) at half scale
    const P2 n = sub(th.re[0], th.im[0]);
    emit_bin(p, s, fp, 256 - p.bin0, add(n, n), mk(0.f, 0.f));
  }
}

// rows of 2 * PAIRS consecutive frames leave the tile, a frame pair (8 bytes) per lane; warp w takes rows w, w + NW, ...
// (the row loop is the whole cost of this phase -- one shared-memory load and one store per row and warp -- so the
// common case carries nothing else: pointers advance by constants, alignment and range are decided once per batch)
template <int PAIRS, int PLANES>
SELDQ_HD void phase_d(const Params& p, const Shared<PAIRS, PLANES>& s, int tid, int signal, int t0) {
  constexpr int NW = PAIRS / 2;                       // warps per block
  constexpr int PITCH = 2 * PAIRS + 2;
  constexpr int LW = PAIRS <= 8 ? 8 : (PAIRS <= 16 ? 16 : 32);      // lanes per row: a warp stores 32 / LW rows at once
  constexpr int RG = 32 / LW;
  const int warp = tid >> 5, lane = tid & 31;
  const int rg = lane / LW, fp = lane - rg * LW;
  if (fp >= PAIRS) return;
  const int t = t0 + 2 * fp;
  if (t >= p.n_frames) return;
  const bool both = t + 1 < p.n_frames;
  const int b = signal / p.n_ch, c = signal - b * p.n_ch;
  const int row0 = warp * RG + rg;
  const long long step = (long long)(NW * RG) * p.n_frames;
#pragma unroll
  for (int pl = 0; pl < PLANES; ++pl) {
    const bool plain = p.norm_sub[pl] == 0.f && p.norm_mul[pl] == 1.f;
    float* dst = p.out + ((long long)((b * PLANES + pl) * p.n_ch + c) * p.n_bins + row0) * p.n_frames + t;
    const float* src = &s.tile[pl][row0][2 * fp];
    const bool vec = both && ((reinterpret_cast<unsigned long long>(dst) | (unsigned long long)(step * 4) |
                               (unsigned long long)((long long)p.n_frames * 4)) & 7ull) == 0;
    if (vec && plain) {
#pragma unroll 4
      for (int kb = row0; kb < p.n_bins; kb += NW * RG, dst += step, src += NW * RG * PITCH)
        *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(src);
      continue;
    }
    const float sub_ = p.norm_sub[pl], mul_ = p.norm_mul[pl];
    for (int kb = row0; kb < p.n_bins; kb += NW * RG, dst += step, src += NW * RG * PITCH) {
      float2 v = *reinterpret_cast<const float2*>(src);
      v.x = (v.x - sub_) * mul_;
      v.y = (v.y - sub_) * mul_;
      if (vec) {
        *reinterpret_cast<float2*>(dst) = v;
      } else {
        dst[0] = v.x;
        if (both) dst[1] = v.y;
      }
    }
  }
}

}  // namespace stft2
}  // namespace seldq

// Batched STFT magnitude / phase front end (utility_functions.py:129-155, spectrum_fast on top of
// scipy.signal.stft defaults): periodic Hamming window, nperseg = 512, zero boundary extension of
// 256 samples, hop = nperseg - noverlap, Z = rfft(frame*w)/sum(w), |Z| and atan2(Im,Re), DC bin
// and last frame optionally dropped, output planes laid out (batch, [mag ch..., phase ch...], bin, frame).
//
// One block works on batches of FRB = 32 consecutive frames of one signal, 16 threads per frame.  A 512-point
// real FFT is a 256-point complex FFT of the packed sequence z[m] = x[2m] + i x[2m+1] followed by the real-input
// untangle step; 256 = 16 x 16, so the complex FFT is two 16-point DFTs held entirely in registers with ONE
// exchange through shared memory between them (four-step FFT):
//   phase_a   thread j loads z[j + 16 r] (16 lanes read 128 contiguous bytes), windows it, DFT16 over r,
//             multiplies by W256^(j k1) and drops column k1 into the exchange buffer
//   phase_b   thread j picks up row k1 = j, DFT16 over n2  ->  Z[j + 16 q]
//   phase_b2 / phase_c   Z goes back through the exchange buffer so that thread j can pair Z[k] with Z[256 - k]
//             (owned by thread 16 - j), untangle, |.| and atan2, into a [bin][frame] staging tile
//   phase_d   the tile leaves as rows of 32 consecutive frames (128-byte stores)
// The staging tile is double-buffered over batches and the raw samples of the next batch are requested before the
// stores of the current one (stft.cu), so a block pays ONE block-wide barrier per batch and the global-load latency of
// batch i + 1 hides behind the stores of batch i.
//
// Phases are __host__ __device__ so tests/host_emul can execute them on the CPU.
//
// Tried and measured in round 2 (tools/kernel_bench.py, profiles/r2r_stft_*.json): the same pipeline with complex
// numbers as float2 and the packed fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: half the ALU instructions of the
// butterflies) needs the twiddles as two pre-arranged pairs each, i.e. twice the shared-memory loads, and came out at
// 73.7 us per clip against 63.5 us for this scalar form -- the kernel is bound by its total instruction issue
// (ALU + LDS), not by the ALU alone.
#pragma once
#include <cmath>
#include <vector_types.h>
#include <vector_functions.h>
#include "common.cuh"

namespace seldq {
namespace stft {

constexpr int NFFT = 512;
constexpr int FRB = 32;          // frames per batch
constexpr int TPF = 16;          // threads per frame
constexpr int NT = FRB * TPF;    // 512 threads
constexpr int MAXBINS = NFFT / 2 + 1;

struct Params {
  const float* x;       // (n_signals, n_samples) float32 ...
  const short* x16;     // ... or, when not null, int16 PCM (scaled by 1 / 32768 on load: half the input bytes)
  float* out;           // (n_batch, planes*n_ch, n_bins, n_frames); may be null with `stats` (statistics pass only)
  long long n_samples;
  int n_ch;             // signals per batch item
  int hop;
  int n_frames;         // frames kept
  int bin0;             // 1 if the DC bin is dropped
  int n_bins;           // 257 - bin0
  int output_phase;
  int groups;           // ceil(n_frames / FRB)
  long long total;      // n_signals * groups
  // dataset normalisation fused into the store (train.py:374-408): out = (v - norm_sub[plane]) * norm_mul[plane]
  // (plane 0 = magnitude, 1 = phase; sub = mean, mul = 1 / std; 0 and 1 leave the features as they are)
  float norm_sub[2], norm_mul[2];
  // when not null: sum and sum of squares of the (un-normalised) features per plane are ADDED to
  // stats[2 * plane], stats[2 * plane + 1] -- the reduction train.py runs with np.mean / np.std over the stored array
  double* stats;
};

struct Shared {
  float2 xch[FRB][16][17];           // exchange buffer, one 16 x 16 complex matrix (pitch 17) per frame
  float tile[2][2][MAXBINS][FRB + 1];   // staging, double-buffered over batches: [buffer][plane][bin][frame]
  float2 tw256[16][16];              // [k1][j] = W256^(j k1)
  float2 tw512[17][16];              // [q][j]  = W512^(j + 16 q)
  float2 win[16][16];                // [r][j]  = window at samples 2 (j + 16 r), 2 (j + 16 r) + 1, / sum(w)
};

struct Thread {
  float re[16], im[16];
};
// running feature statistics of a thread (Params::stats): separate scalars -- a dynamically indexed member would
// send the whole register-resident Thread to local memory
struct Stats {
  float mag1, mag2, ph1, ph2;
};

SELDQ_HD void sincospi_f(float x, float* sn, float* cs) {
#if defined(__CUDA_ARCH__)
  sincospif(x, sn, cs);
#else
  *sn = (float)sin(3.14159265358979323846 * (double)x);
  *cs = (float)cos(3.14159265358979323846 * (double)x);
#endif
}

// twiddles and the periodic Hamming window w[n] = 0.54 - 0.46 cos(2 pi n / 512)
// (scipy.signal.get_window('hamming', 512)), pre-divided by sum(w) = 0.54 * 512 (scaling='spectrum')
SELDQ_HD void init_tables(Shared& s, int tid) {
  for (int idx = tid; idx < 17 * 16; idx += NT) {
    const int q = idx >> 4, j = idx & 15;
    float sn, cs;
    sincospi_f((float)(j + 16 * q) / 256.0f, &sn, &cs);     // exp(-2 pi i k / 512)
    s.tw512[q][j] = make_float2(cs, -sn);
    if (q < 16) {
      sincospi_f((float)(j * q) / 128.0f, &sn, &cs);        // exp(-2 pi i j k1 / 256)
      s.tw256[q][j] = make_float2(cs, -sn);
      float w[2];
      for (int h = 0; h < 2; ++h) {
        sincospi_f((float)(2 * (j + 16 * q) + h) / 256.0f, &sn, &cs);
        w[h] = (0.54f - 0.46f * cs) * (1.0f / (0.54f * NFFT));
      }
      s.win[q][j] = make_float2(w[0], w[1]);
    }
  }
}

// forward 4-point DFT of (a0, a1, a2, a3) in place
#define SELDQ_DFT4(r0, i0, r1, i1, r2, i2, r3, i3)                                   \
  {                                                                                   \
    const float t0r = r0 + r2, t0i = i0 + i2, t1r = r0 - r2, t1i = i0 - i2;           \
    const float t2r = r1 + r3, t2i = i1 + i3, t3r = i1 - i3, t3i = -(r1 - r3);        \
    r0 = t0r + t2r; i0 = t0i + t2i; r1 = t1r + t3r; i1 = t1i + t3i;                   \
    r2 = t0r - t2r; i2 = t0i - t2i; r3 = t1r - t3r; i3 = t1i - t3i;                   \
  }

// forward 16-point DFT in registers, natural order in and out:  X[k] = sum_n x[n] W16^(n k)
// (n = 4 n1 + n2, k = k1 + 4 k2:  DFT4 over n1, twiddle W16^(n2 k1), DFT4 over n2)
SELDQ_HD void dft16(float (&re)[16], float (&im)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  // step 1: for each n2, DFT4 over n1 of x[4 n1 + n2]; result index k1 lands at position 4 k1 + n2
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) SELDQ_DFT4(re[n2], im[n2], re[4 + n2], im[4 + n2], re[8 + n2], im[8 + n2], re[12 + n2], im[12 + n2]);
  // step 2: y[k1][n2] *= W16^(n2 k1), W16^m = (cos, -sin)(2 pi m / 16)
#define SELDQ_CMUL(idx, wr, wi)                         \
  {                                                     \
    const float a = re[idx], b = im[idx];               \
    re[idx] = a * (wr) - b * (wi);                      \
    im[idx] = a * (wi) + b * (wr);                      \
  }
  SELDQ_CMUL(4 * 1 + 1, c1, -s1)    // m = 1
  SELDQ_CMUL(4 * 1 + 2, h, -h)      // m = 2
  SELDQ_CMUL(4 * 1 + 3, s1, -c1)    // m = 3
  SELDQ_CMUL(4 * 2 + 1, h, -h)      // m = 2
  {                                  // m = 4: * (-i)
    const float a = re[4 * 2 + 2], b = im[4 * 2 + 2];
    re[4 * 2 + 2] = b; im[4 * 2 + 2] = -a;
  }
  SELDQ_CMUL(4 * 2 + 3, -h, -h)     // m = 6
  SELDQ_CMUL(4 * 3 + 1, s1, -c1)    // m = 3
  SELDQ_CMUL(4 * 3 + 2, -h, -h)     // m = 6
  SELDQ_CMUL(4 * 3 + 3, -c1, s1)    // m = 9
#undef SELDQ_CMUL
  // step 3: for each k1, DFT4 over n2 of y[k1][n2]; result k2 lands at position 4 k1 + k2 = X[k1 + 4 k2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    SELDQ_DFT4(re[4 * k1], im[4 * k1], re[4 * k1 + 1], im[4 * k1 + 1], re[4 * k1 + 2], im[4 * k1 + 2], re[4 * k1 + 3],
               im[4 * k1 + 3]);
  // positions hold X[k1 + 4 k2] at 4 k1 + k2: transpose the 4 x 4 index grid into natural order
#define SELDQ_SWAP(a, b)                                                     \
  {                                                                          \
    const float tr = re[a], ti = im[a];                                      \
    re[a] = re[b]; im[a] = im[b]; re[b] = tr; im[b] = ti;                    \
  }
  SELDQ_SWAP(1, 4) SELDQ_SWAP(2, 8) SELDQ_SWAP(3, 12) SELDQ_SWAP(6, 9) SELDQ_SWAP(7, 13) SELDQ_SWAP(11, 14)
#undef SELDQ_SWAP
}
#undef SELDQ_DFT4

// batch index -> signal, first frame
SELDQ_HD void batch_decode(const Params& p, long long batch, int* signal, int* t0) {
  *signal = (int)(batch / p.groups);
  *t0 = (int)(batch - (long long)(*signal) * p.groups) * FRB;
}

// raw samples of this thread's 16 complex points z[j + 16 r] = x[2 m] + i x[2 m + 1] (zero outside the signal)
SELDQ_HD void phase_a_load(const Params& p, Thread& th, int tid, int signal, int t0) {
  const int f = tid >> 4, j = tid & 15;
  const long long g0 = (long long)(t0 + f) * p.hop - NFFT / 2;
  const bool interior = g0 >= 0 && g0 + NFFT <= p.n_samples;
  if (p.x16 != nullptr) {            // int16 PCM ingest
    const short* s16 = p.x16 + (long long)signal * p.n_samples;
    const float sc = 1.0f / 32768.0f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const long long g = g0 + 2 * (j + 16 * r);
#if defined(__CUDA_ARCH__)
      if (interior && ((reinterpret_cast<unsigned long long>(s16 + g0) & 3ull) == 0)) {
        const short2 v = __ldg(reinterpret_cast<const short2*>(s16 + g0) + (j + 16 * r));
        th.re[r] = (float)v.x * sc;
        th.im[r] = (float)v.y * sc;
        continue;
      }
#endif
      th.re[r] = (interior || (g >= 0 && g < p.n_samples)) ? (float)s16[g] * sc : 0.f;
      th.im[r] = (interior || (g + 1 >= 0 && g + 1 < p.n_samples)) ? (float)s16[g + 1] * sc : 0.f;
    }
    return;
  }
  const float* src = p.x + (long long)signal * p.n_samples;
#if defined(__CUDA_ARCH__)
  if (interior && ((reinterpret_cast<unsigned long long>(src + g0) & 7ull) == 0)) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src + g0) + (j + 16 * r));
      th.re[r] = v.x;
      th.im[r] = v.y;
    }
  } else
#endif
  {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const long long g = g0 + 2 * (j + 16 * r);
      th.re[r] = (interior || (g >= 0 && g < p.n_samples)) ? src[g] : 0.f;
      th.im[r] = (interior || (g + 1 >= 0 && g + 1 < p.n_samples)) ? src[g + 1] : 0.f;
    }
  }
}

// window, DFT16 over r, twiddle, into the exchange buffer
SELDQ_HD void phase_a_compute(Shared& s, Thread& th, int tid) {
  const int f = tid >> 4, j = tid & 15;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 w = s.win[r][j];
    th.re[r] *= w.x;
    th.im[r] *= w.y;
  }
  dft16(th.re, th.im);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    const float2 w = s.tw256[k1][j];
    s.xch[f][k1][j] = make_float2(th.re[k1] * w.x - th.im[k1] * w.y, th.re[k1] * w.y + th.im[k1] * w.x);
  }
}

SELDQ_HD void phase_a(const Params& p, Shared& s, Thread& th, int tid, int signal, int t0) {
  phase_a_load(p, th, tid, signal, t0);
  phase_a_compute(s, th, tid);
}

SELDQ_HD void phase_b(const Shared& s, Thread& th, int tid) {
  const int f = tid >> 4, j = tid & 15;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    const float2 v = s.xch[f][j][n2];
    th.re[n2] = v.x;
    th.im[n2] = v.y;
  }
  dft16(th.re, th.im);          // th[q] = Z[j + 16 q]
}

SELDQ_HD void phase_b2(Shared& s, const Thread& th, int tid) {
  const int f = tid >> 4, j = tid & 15;
#pragma unroll
  for (int q = 0; q < 16; ++q) s.xch[f][q][j] = make_float2(th.re[q], th.im[q]);
}

SELDQ_HD void emit_bin(const Params& p, Shared& s, int buf, int f, int kb, float xr, float xi) {
  if (kb < 0) return;
#if defined(__CUDA_ARCH__)
  float mag;                                               // sqrt.approx: one MUFU (+ a multiply), relative error 2^-22 against a
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(xr * xr + xi * xi));     // tolerance of 1e-4; 0, inf and NaN behave as in sqrtf
  s.tile[buf][0][kb][f] = mag;
#else
  s.tile[buf][0][kb][f] = sqrtf(xr * xr + xi * xi);
#endif
  if (p.output_phase) s.tile[buf][1][kb][f] = atan2f(xi, xr);
}

// real-input untangle: R[k] = E[k] + W512^k O[k], E = (Z[k] + conj Z[256-k]) / 2, O = (Z[k] - conj Z[256-k]) / 2i
SELDQ_HD void phase_c(const Params& p, Shared& s, const Thread& th, int tid, int buf = 0) {
  const int f = tid >> 4, j = tid & 15;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float2 pz = j == 0 ? s.xch[f][(16 - q) & 15][0] : s.xch[f][15 - q][16 - j];
    const float a = th.re[q], bb = th.im[q], cc = pz.x, d = pz.y;
    const float er = 0.5f * (a + cc), ei = 0.5f * (bb - d);
    const float orr = 0.5f * (bb + d), oi = -0.5f * (a - cc);
    const float2 w = s.tw512[q][j];
    emit_bin(p, s, buf, f, j + 16 * q - p.bin0, er + w.x * orr - w.y * oi, ei + w.x * oi + w.y * orr);
  }
  if (j == 0) emit_bin(p, s, buf, f, 256 - p.bin0, th.re[0] - th.im[0], 0.f);     // Nyquist bin
}

// rows of FRB consecutive frames; warp w takes rows w, w + 16, ...; the dataset normalisation is applied on the way
// out and the running statistics of the un-normalised features are kept per thread
SELDQ_HD void phase_d(const Params& p, const Shared& s, Stats& st, int tid, int signal, int t0, int buf = 0) {
  const int b = signal / p.n_ch, c = signal - b * p.n_ch;
  const int planes = p.output_phase ? 2 : 1;
  const int warp = tid >> 5, lane = tid & 31;
  const int t = t0 + lane;
  if (t >= p.n_frames) return;
  if (p.stats == nullptr && p.norm_sub[0] == 0.f && p.norm_mul[0] == 1.f && p.norm_sub[1] == 0.f && p.norm_mul[1] == 1.f) {
    for (int row = warp; row < planes * p.n_bins; row += NT / 32) {      // plain features (spectrum_fast)
      const int plane = row >= p.n_bins ? 1 : 0, kb = row - plane * p.n_bins;
      float* dst = p.out + ((long long)(b * planes * p.n_ch + plane * p.n_ch + c) * p.n_bins + kb) * p.n_frames;
      dst[t] = s.tile[buf][plane][kb][lane];
    }
    return;
  }
  for (int row = warp; row < planes * p.n_bins; row += NT / 32) {
    const int plane = row >= p.n_bins ? 1 : 0, kb = row - plane * p.n_bins;
    const float v = s.tile[buf][plane][kb][lane];
    if (p.stats != nullptr) {
      if (plane == 0) { st.mag1 += v; st.mag2 = fmaf(v, v, st.mag2); }
      else { st.ph1 += v; st.ph2 = fmaf(v, v, st.ph2); }
    }
    if (p.out != nullptr) {
      float* dst = p.out + ((long long)(b * planes * p.n_ch + plane * p.n_ch + c) * p.n_bins + kb) * p.n_frames;
      dst[t] = (v - p.norm_sub[plane]) * p.norm_mul[plane];
    }
  }
}

}  // namespace stft
}  // namespace seldq

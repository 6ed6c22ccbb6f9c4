// Batched STFT magnitude / phase front end (utility_functions.py:129-155, spectrum_fast on top of
// scipy.signal.stft defaults): periodic Hamming window, nperseg = 512, zero boundary extension of
// 256 samples, hop = nperseg - noverlap, Z = rfft(frame*w)/sum(w), |Z| and atan2(Im,Re), DC bin
// and last frame optionally dropped, output planes laid out (batch, [mag ch..., phase ch...], bin, frame).
//
// One block handles STFT_FR consecutive frames of one signal.  The overlapping frames are staged
// once in shared memory; each 512-point real FFT is a 256-point complex radix-4 Stockham FFT
// (4 passes, 64 threads per frame) followed by the real-input untangle step.  Results of 8
// frames at a time are written with consecutive lanes on consecutive frames.
//
// Phases are __host__ __device__ so tests/host_emul can execute them on the CPU.
#pragma once
#include <cmath>
#include <vector_types.h>
#include <vector_functions.h>
#include "common.cuh"

namespace seldq {
namespace stft {

constexpr int NFFT = 512;
constexpr int NC = 256;          // complex FFT length
constexpr int FR = 32;           // frames per block
constexpr int FPR = 8;           // frames per round
constexpr int NT = 64 * FPR;     // 512 threads

struct Params {
  const float* x;       // (n_signals, n_samples)
  float* out;           // (n_batch, planes*n_ch, n_bins, n_frames)
  long long n_samples;
  int n_ch;             // signals per batch item
  int hop;
  int n_frames;         // frames kept
  int bin0;             // 1 if the DC bin is dropped
  int n_bins;           // 257 - bin0
  int output_phase;
};

struct Shared {
  float* samples;                // (FR-1)*hop + NFFT floats
  float2 tw[NFFT];
  float win[NFFT];
  float re[2][FPR][NC + 4];
  float im[2][FPR][NC + 4];
};

SELDQ_HD void sincospi_f(float x, float* sn, float* cs) {
#if defined(__CUDA_ARCH__)
  sincospif(x, sn, cs);
#else
  *sn = (float)sin(3.14159265358979323846 * (double)x);
  *cs = (float)cos(3.14159265358979323846 * (double)x);
#endif
}

SELDQ_HD int span(int hop) { return (FR - 1) * hop + NFFT; }

SELDQ_HD void load(const Params& p, Shared& s, int tid, int bx, int by) {
  const long long g0 = (long long)bx * FR * p.hop - NFFT / 2;
  const float* src = p.x + (long long)by * p.n_samples;
  const int n = span(p.hop);
  for (int i = tid; i < n; i += NT) {
    const long long g = g0 + i;
    s.samples[i] = (g >= 0 && g < p.n_samples) ? src[g] : 0.f;
  }
  // twiddles exp(-2*pi*i*k/512) and the periodic Hamming window w[n] = 0.54 - 0.46 cos(2*pi*n/512)
  // (scipy.signal.get_window('hamming', 512)), pre-divided by sum(w) = 0.54 * 512 (scaling='spectrum')
  for (int i = tid; i < NFFT; i += NT) {
    float sn, cs;
    sincospi_f(i / 256.0f, &sn, &cs);
    s.tw[i] = make_float2(cs, -sn);
    s.win[i] = (0.54f - 0.46f * cs) * (1.0f / (0.54f * NFFT));
  }
}

// window and pack two real samples into one complex point, 8 frames per round
SELDQ_HD void pack(const Params& p, Shared& s, int tid, int round) {
#pragma unroll
  for (int q = 0; q < (FPR * NC) / NT; ++q) {
    const int idx = tid + q * NT;
    const int fr = idx / NC, m = idx - fr * NC;
    const float* f = s.samples + (round * FPR + fr) * p.hop;
    s.re[0][fr][m] = f[2 * m] * s.win[2 * m];
    s.im[0][fr][m] = f[2 * m + 1] * s.win[2 * m + 1];
  }
}

// one radix-4 Stockham pass; Ns = 1, 4, 16, 64; reads buffer src, writes buffer 1-src
SELDQ_HD void fft_pass(Shared& s, int tid, int Ns, int src) {
  const int fr = tid >> 6, j = tid & 63;
  const int k = j & (Ns - 1);
  const int tstep = k * (128 / Ns);          // W_{4Ns}^{k r} = W_512^{k r 128/Ns}
  float vr[4], vi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float a = s.re[src][fr][j + 64 * r], b = s.im[src][fr][j + 64 * r];
    const float2 w = s.tw[tstep * r];
    vr[r] = a * w.x - b * w.y;
    vi[r] = a * w.y + b * w.x;
  }
  const float t0r = vr[0] + vr[2], t0i = vi[0] + vi[2];
  const float t1r = vr[0] - vr[2], t1i = vi[0] - vi[2];
  const float t2r = vr[1] + vr[3], t2i = vi[1] + vi[3];
  const float t3r = vi[1] - vi[3], t3i = -(vr[1] - vr[3]);   // (v1 - v3) * (-i)
  const int j0 = (j / Ns) * Ns * 4 + k;
  const int dst = 1 - src;
  s.re[dst][fr][j0] = t0r + t2r;           s.im[dst][fr][j0] = t0i + t2i;
  s.re[dst][fr][j0 + Ns] = t1r + t3r;      s.im[dst][fr][j0 + Ns] = t1i + t3i;
  s.re[dst][fr][j0 + 2 * Ns] = t0r - t2r;  s.im[dst][fr][j0 + 2 * Ns] = t0i - t2i;
  s.re[dst][fr][j0 + 3 * Ns] = t1r - t3r;  s.im[dst][fr][j0 + 3 * Ns] = t1i - t3i;
}

// real-input untangle + magnitude / phase + store; the spectrum of round `round` is in buffer 0
SELDQ_HD void emit(const Params& p, const Shared& s, int tid, int bx, int by, int round) {
  const int b = by / p.n_ch, c = by - b * p.n_ch;
  const int planes = p.output_phase ? 2 : 1;
  float* mag = p.out + ((long long)(b * planes * p.n_ch + c) * p.n_bins) * p.n_frames;
  float* pha = p.out + ((long long)(b * planes * p.n_ch + p.n_ch + c) * p.n_bins) * p.n_frames;
  for (int idx = tid; idx < FPR * p.n_bins; idx += NT) {
    const int fr = idx & (FPR - 1), kb = idx / FPR;
    const int t = bx * FR + round * FPR + fr;
    if (t >= p.n_frames) continue;
    const int k = kb + p.bin0;             // 0..256
    const int k1 = k & (NC - 1), k2 = (NC - k) & (NC - 1);
    const float a = s.re[0][fr][k1], bb = s.im[0][fr][k1];
    const float cc = s.re[0][fr][k2], d = s.im[0][fr][k2];
    const float er = 0.5f * (a + cc), ei = 0.5f * (bb - d);
    const float orr = 0.5f * (bb + d), oi = -0.5f * (a - cc);
    const float2 w = s.tw[k];
    const float xr = er + w.x * orr - w.y * oi;
    const float xi = ei + w.x * oi + w.y * orr;
    const long long o = (long long)kb * p.n_frames + t;
    mag[o] = sqrtf(xr * xr + xi * xi);
    if (p.output_phase) pha[o] = atan2f(xi, xr);
  }
}

}  // namespace stft
}  // namespace seldq

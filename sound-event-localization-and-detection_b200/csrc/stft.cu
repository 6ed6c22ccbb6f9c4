// __global__ wrapper of the STFT magnitude / phase kernel (phase functions in stft.cuh).
#include <cuda_runtime.h>

#include "conv_cl.h"
#include "launch.h"
#include <cstdlib>

#include "stft.cuh"
#include "stft_pair.cuh"

namespace seldq {
namespace stft {

__global__ void __launch_bounds__(NT, 1) stft_magphase_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& s = *reinterpret_cast<Shared*>(smem_raw);
  init_tables(s, threadIdx.x);
  __syncthreads();
  Thread th;
  Stats acc = {0.f, 0.f, 0.f, 0.f};
  int buf = 0;
  long long batch = blockIdx.x;
  int signal = 0, t0 = 0;
  if (batch < p.total) {
    batch_decode(p, batch, &signal, &t0);
    phase_a_load(p, th, threadIdx.x, signal, t0);
  }
  for (; batch < p.total; batch += gridDim.x, buf ^= 1) {
    phase_a_compute(s, th, threadIdx.x);
    __syncwarp();                                  // a frame lives in one half-warp
    phase_b(s, th, threadIdx.x);
    __syncwarp();
    phase_b2(s, th, threadIdx.x);
    __syncwarp();
    phase_c(p, s, th, threadIdx.x, buf);
    // The only block-wide barrier of a batch.  It also orders this batch's writes of tile[buf] after the reads
    // phase_d made of it two batches ago: every thread passed the previous batch's barrier after those reads.
    __syncthreads();
    const int cur_signal = signal, cur_t0 = t0;
    const long long next = batch + gridDim.x;
    if (next < p.total) {                          // the next batch's samples are in flight during the stores below
      batch_decode(p, next, &signal, &t0);
      phase_a_load(p, th, threadIdx.x, signal, t0);
    }
    phase_d(p, s, acc, threadIdx.x, cur_signal, cur_t0, buf);
  }
  if (p.stats != nullptr) {
    // feature statistics (sum, sum of squares per plane): warp shuffle, then one double atomicAdd per warp and quantity
    const int planes = p.output_phase ? 2 : 1;
    for (int pl = 0; pl < planes; ++pl) {
      float a = pl == 0 ? acc.mag1 : acc.ph1, b = pl == 0 ? acc.mag2 : acc.ph2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(p.stats + 2 * pl, (double)a);
        atomicAdd(p.stats + 2 * pl + 1, (double)b);
      }
    }
  }
}

}  // namespace stft

namespace stft2 {

// frame-pair kernel (stft_pair.cuh): one persistent block per SM, 16 threads per pair of consecutive frames
template <int PAIRS, int BPS, int PLANES>
__global__ void __launch_bounds__(16 * PAIRS, BPS) stft_pair_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared<PAIRS, PLANES>& s = *reinterpret_cast<Shared<PAIRS, PLANES>*>(smem_raw);
  init_tables(s, threadIdx.x);
  __syncthreads();
  Raw raw;
  Thread th;
  long long batch = blockIdx.x;
  int signal = 0, t0 = 0;
  if (batch < p.total) {
    batch_decode<PAIRS>(p, batch, &signal, &t0);
    load_raw(p, raw, threadIdx.x, signal, t0);
  }
  for (; batch < p.total; batch += gridDim.x) {
    phase_a(s, raw, th, threadIdx.x);
    __syncwarp();                                  // a frame pair lives in one half-warp
    phase_b(s, th, threadIdx.x);
    __syncwarp();
    phase_b2(s, th, threadIdx.x);
    __syncthreads();                               // the row stores of the previous batch have read the tile
    phase_c(p, s, th, threadIdx.x);
    __syncthreads();                               // the tile is complete
    const int cur_signal = signal, cur_t0 = t0;
    const long long next = batch + gridDim.x;
    if (next < p.total) {                          // the next batch's samples are in flight during the stores below
      batch_decode<PAIRS>(p, next, &signal, &t0);
      load_raw(p, raw, threadIdx.x, signal, t0);
    }
    phase_d(p, s, threadIdx.x, cur_signal, cur_t0);
  }
}

// BPS blocks per SM (each with its own tables, exchange buffer and tile): two or three smaller blocks decouple the
// block-wide barriers and overlap one block's row stores with another's butterflies
template <int PAIRS, int BPS, int PLANES = 1>
static int launch_pairs(Params& p, int n_signals, cudaStream_t st) {
  const size_t smem = sizeof(Shared<PAIRS, PLANES>);
  const cudaError_t e = cudaFuncSetAttribute(stft_pair_kernel<PAIRS, BPS, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "stft (frame pairs) smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
  p.groups = (p.n_frames + 2 * PAIRS - 1) / (2 * PAIRS);
  p.total = (long long)n_signals * p.groups;
  if (p.total >= (1LL << 31)) return fail(SELDQ_ERR_UNSUPPORTED, "stft: too many frame batches (%lld)", p.total);
  const long long slots = (long long)cl::num_sms() * BPS;
  const unsigned grid = (unsigned)(p.total < slots ? p.total : slots);
  stft_pair_kernel<PAIRS, BPS, PLANES><<<grid, 16 * PAIRS, smem, st>>>(p);
  return check_launch("stft_pair_kernel");
}

// Frame pairs per batch and blocks per SM.  Registers (168 per thread without spills) allow 384 threads per SM, shared
// memory (6.4 KB per frame pair + 6 KB of tables per block) about 34 pairs.  Measured on 8 signals x 4800 frames per
// clip, us per clip at batch 1 / 4 / 16 (tools/kernel_bench.py, gpurun_out/r2af): 3 blocks of 8 pairs 41.0 / 33.3 / 31.0,
// 2 blocks of 12 pairs 43.0 / 32.8 / 30.2, one block of 24 pairs 45.1 / 35.3 / 32.5, of 16 pairs 51.2 / 37.9 / 34.9, of
// 32 pairs (128 registers: spills) 63.5 / 47.1 / 41.8; the per-frame kernel of stft.cuh 63.5 / 51.4 / 47.8.  Several small
// blocks decouple the block-wide barriers and overlap one block's row stores with another's butterflies.
// SELDQ_STFT_PAIRS overrides.
static int pick_pairs(const Params& p, int n_signals) {
  if (const char* e = getenv("SELDQ_STFT_PAIRS")) {
    const int v = atoi(e);
    if (v == 8 || v == 12 || v == 16 || v == 20 || v == 24 || v == 28 || v == 32) return v;
  }
  return (long long)n_signals * p.n_frames >= 150000 ? 12 : 8;
}

}  // namespace stft2

int launch_stft(stft::Params& p, int n_signals, cudaStream_t st) {
  // magnitude-only feature extraction takes the frame-pair kernel (SELDQ_STFT_PAIR=0: always the per-frame kernel)
  static const bool pair_ok = [] { const char* e = getenv("SELDQ_STFT_PAIR"); return !(e && e[0] == '0'); }();
  if (pair_ok && p.output_phase && p.stats == nullptr && p.out != nullptr)      // two staging planes: 2 blocks of 12 pairs per SM
    return stft2::launch_pairs<12, 2, 2>(p, n_signals, st);
  if (pair_ok && !p.output_phase && p.stats == nullptr && p.out != nullptr) {
    switch (stft2::pick_pairs(p, n_signals)) {
      case 8: return stft2::launch_pairs<8, 3>(p, n_signals, st);
      case 12: return stft2::launch_pairs<12, 2>(p, n_signals, st);
      case 16: return stft2::launch_pairs<16, 1>(p, n_signals, st);
      case 20: return stft2::launch_pairs<20, 1>(p, n_signals, st);
      case 28: return stft2::launch_pairs<28, 1>(p, n_signals, st);
      case 32: return stft2::launch_pairs<32, 1>(p, n_signals, st);
      default: return stft2::launch_pairs<24, 1>(p, n_signals, st);
    }
  }
  const size_t smem = sizeof(stft::Shared);
  static thread_local bool configured = false;
  if (!configured) {
    const cudaError_t e = cudaFuncSetAttribute(stft::stft_magphase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem);
    if (e != cudaSuccess) return fail(SELDQ_ERR_CUDA, "stft smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  p.groups = (p.n_frames + stft::FRB - 1) / stft::FRB;
  p.total = (long long)n_signals * p.groups;
  const long long sms = cl::num_sms();
  const unsigned grid = (unsigned)(p.total < sms ? p.total : sms);     // one resident block per SM
  stft::stft_magphase_kernel<<<grid, stft::NT, smem, st>>>(p);
  return check_launch("stft_magphase_kernel");
}

}  // namespace seldq

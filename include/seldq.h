/*
 * seldq.h -- C ABI of libseldq.so: the B200 (sm_100a) implementation of the quaternion /
 * dual-quaternion convolution + linear stack and the STFT magnitude/phase front end of
 * AuroraEchos/Sound-Event-Localization-and-Detection.
 *
 * Each entry point replaces one call site of the reference (file:line relative to the
 * reference tree):
 *   seldq_conv_fwd / _dgrad / _wgrad   quaternion/quaternion_ops.py:125-147 (quaternion_conv)
 *                                      dual_quaternion/dual_quaternion_ops.py:111-153 (dual_quaternion_conv)
 *                                      and the autograd of F.conv1d/F.conv2d + torch.cat behind them
 *   seldq_linear_fwd / _dgrad / _wgrad quaternion_ops.py:299-327 (quaternion_linear),
 *                                      :392-464 (QuaternionLinearFunction),
 *                                      dual_quaternion_ops.py:156-203 (dual_quaternion_linear)
 *   seldq_stft_magphase                utility_functions.py:129-155 (spectrum_fast)
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *     the caller (PyTorch) owns all buffers including workspaces and outputs; the library
 *     never allocates or frees device memory and keeps no reference after returning.
 *   - tensors are contiguous, channels-first (NCW / NCHW), float32, exactly as the reference
 *     hands them to F.conv*; compact weights are the reference's Parameters:
 *     conv (Cout/nc, Cin/nc, k...) and linear (in/nc, out/nc), nc = 1 | 4 | 8.
 *   - weight pointer arrays are HOST arrays of nc device pointers in the reference's order
 *     r, i, j, k [, r_2, i_2, j_2, k_2].
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous, nothing in the
 *     library synchronises the device.
 *   - return 0 on success, a negative seldq_status_t otherwise; seldq_last_error() returns a
 *     thread-local message.  The library is re-entrant (no mutable globals besides that string
 *     and a host-side cache of immutable driver entry points).
 *   - there is no CPU fallback: with no CUDA device every compute call returns SELDQ_ERR_CUDA.
 */
#ifndef SELDQ_H_
#define SELDQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELDQ_ABI_VERSION 5

typedef enum {
  SELDQ_OK = 0,
  SELDQ_ERR_INVALID = -1,     /* bad descriptor / null pointer / shape mismatch           */
  SELDQ_ERR_UNSUPPORTED = -2, /* valid request the chosen precision path cannot serve    */
  SELDQ_ERR_WORKSPACE = -3,   /* workspace smaller than seldq_*_workspace_bytes()        */
  SELDQ_ERR_CUDA = -4         /* CUDA runtime / driver error (message has the details)   */
} seldq_status_t;

/* which block structure expands the compact weights (Hamilton tables) */
typedef enum {
  SELDQ_ALG_REAL = 0,      /* nc = 1, plain convolution / linear                              */
  SELDQ_ALG_Q = 1,         /* nc = 4, Wq[a*O+o, b*I+i] = sign[a][b] * W_{a^b}[o,i]            */
  SELDQ_ALG_DQ = 2,        /* nc = 8, [[Q(w),0],[Q(w2),Q(w)]]; linear uses the transposed form */
  SELDQ_ALG_DQ_LINEAR = 3, /* nc = 8, the block table of dual_quaternion_linear (dual_quaternion_ops.py:170-188)
                              used as a convolution table: lets a DQ linear layer run as a 1x1 convolution over
                              the transposed matrices on the tensor-core path (functional.block_linear)          */
  /* the linear layers as 1x1 convolutions that read (and, in seldq_conv_wgrad, write) the layer's OWN compact
   * tensors, stored (in/nc, out/nc) as quaternion_layers.py:235-238 / dual_quaternion_layers.py keep them: no
   * transposed copies, the weight gradient lands in the parameters' gradient buffers.  Kernel 1 x 1 only. */
  SELDQ_ALG_Q_LINEAR_IO = 4,    /* table of SELDQ_ALG_Q         */
  SELDQ_ALG_DQ_LINEAR_IO = 5    /* table of SELDQ_ALG_DQ_LINEAR */
} seldq_algebra_t;

/* arithmetic the contraction runs in */
typedef enum {
  SELDQ_PREC_FP32 = 0,     /* FFMA, fp32 operands and accumulation (parity gate: rel 1e-4)    */
  SELDQ_PREC_BF16 = 1      /* tcgen05 kind::f16 MMA, bf16 operands, fp32 accumulation in TMEM */
} seldq_precision_t;

typedef enum { SELDQ_PASS_FWD = 0, SELDQ_PASS_DGRAD = 1, SELDQ_PASS_WGRAD = 2 } seldq_pass_t;

typedef struct {
  int32_t algebra;     /* seldq_algebra_t                                            */
  int32_t precision;   /* seldq_precision_t                                          */
  int32_t ndim;        /* 1 (NCW, in_h = k_h = 1) or 2 (NCHW)                        */
  int32_t batch;
  int32_t cin, cout;   /* expanded channel counts, multiples of nc                   */
  int32_t in_h, in_w;
  int32_t k_h, k_w;
  int32_t stride_h, stride_w;
  int32_t pad_h, pad_w;
  int32_t dil_h, dil_w;
} seldq_conv_desc_t;

typedef struct {
  int32_t algebra;     /* SELDQ_ALG_Q -> quaternion_linear table, SELDQ_ALG_DQ -> dual_quaternion_linear */
  int32_t precision;
  int32_t rows;        /* flattened leading dims (T*N)                               */
  int32_t in_features, out_features;   /* expanded                                    */
} seldq_linear_desc_t;

int seldq_abi_version(void);
const char* seldq_last_error(void);
/* number of visible CUDA devices (0 on a CPU-only host); never fails */
int seldq_device_count(void);

/* ---- convolution (A1, A2 in SURVEY.md 8a) ------------------------------------------------ */
int seldq_conv_out_shape(const seldq_conv_desc_t* d, int32_t* out_h, int32_t* out_w);
size_t seldq_conv_workspace_bytes(const seldq_conv_desc_t* d, int32_t pass);

/* bf16 operands of the tensor-core path (SELDQ_PREC_BF16).  The tcgen05 kernels read bf16 copies of the
 * activations / gradients, laid out for TMA:
 *   "CL operand"  [n][h][w][Cp]  channels innermost, every quaternion component padded to a multiple of 16
 *                 channels (pad channels zero); x in this form feeds forward and wgrad, gy feeds dgrad;
 *   "T16 operand" [n][c][h][pitch]  the tensor's own NCHW order in bf16, pitch = w rounded up to 8 (pad
 *                 columns zero); gy in this form feeds wgrad.
 * Every *_cl / *_t16 argument below is optional: when NULL the library stages the operand from the float32
 * tensor into the workspace first.  A caller that stages once (seldq_stage_operand) and reuses the copy
 * for forward and wgrad saves the second pass.  which = 0 selects x, 1 selects gy.
 *   dense != 0: the layer's K side has fewer than 8 channels per component (first CNN layer) or the algebra
 *   is real; its x operand is a single 16-channel-padded component, and its weight gradient reads the
 *   float32 x (argument x of seldq_conv_wgrad must not be NULL). */
int seldq_conv_operand_info(const seldq_conv_desc_t* d, int32_t which, int32_t* padded_channels, int32_t* dense,
                            size_t* cl_bytes, size_t* t16_bytes);
int seldq_stage_operand(const seldq_conv_desc_t* d, int32_t which, const float* src, void* dst_cl, void* dst_t16,
                        void* stream);

/* compact fp32 weights -> bf16 UMMA tiles of the COMPACT tensors (4 | 8 images, never the expanded
 * weight), one set per pass (SELDQ_PASS_FWD, SELDQ_PASS_DGRAD).  Valid until the weights change, i.e.
 * once per optimiser step.  seldq_conv_packed_bytes is 0 for dense layers (nothing to pack). */
size_t seldq_conv_packed_bytes(const seldq_conv_desc_t* d, int32_t pass);
int seldq_conv_pack_weights(const seldq_conv_desc_t* d, int32_t pass, const float* const* host_w, void* packed,
                            void* stream);

/* The same for many layers in ONE launch (a training step re-packs every layer after the optimiser update):
 * seldq_conv_pack_table_fill writes one entry (seldq_conv_pack_table_entry_bytes() bytes, host memory) per
 * (layer, pass) and reports its work items; the caller keeps a device copy of the table while the weight and
 * packed buffers stay in place and replays it with seldq_conv_pack_table_run(dev_table, entries, max items). */
size_t seldq_conv_pack_table_entry_bytes(void);
int seldq_conv_pack_table_fill(const seldq_conv_desc_t* d, int32_t pass, const float* const* host_w, void* packed,
                               void* host_entry, int32_t* items);
int seldq_conv_pack_table_run(const void* dev_table, int32_t count, int32_t max_items, void* stream);

/* y = conv(x, expand(w)) + bias.   x may be NULL when x_cl is given (BF16 path). */
int seldq_conv_fwd(const seldq_conv_desc_t* d, const float* x, const void* x_cl,
                   const float* const* host_w, const void* packed_w, const float* bias, float* y,
                   void* workspace, size_t workspace_bytes, void* stream);

/* same (bf16 operands), but y is written once in 16 bits, IEEE fp16 saturating at +-65504, in the same NCHW order:
 * the form the fused CNN-block glue below consumes.  fp16 and not bf16 because this tensor is only stored, never a
 * tensor-core operand: the 8x finer rounding keeps the fused path inside the bf16 error budget. */
int seldq_conv_fwd_bf16(const seldq_conv_desc_t* d, const float* x, const void* x_cl,
                        const float* const* host_w, const void* packed_w, const float* bias, void* y_f16,
                        void* workspace, size_t workspace_bytes, void* stream);

/* gx = conv_transpose(gy, expand(w)) : gradient w.r.t. the input */
int seldq_conv_dgrad(const seldq_conv_desc_t* d, const float* gy, const void* gy_cl,
                     const float* const* host_w, const void* packed_w, float* gx,
                     void* workspace, size_t workspace_bytes, void* stream);

/* compact weight gradients (nc tensors) and, if gbias != NULL, the bias gradient sum(gy).
 * accumulate == 0: the outputs are overwritten; != 0: the results are ADDED to what the buffers hold, so a
 * caller can point them straight at its gradient bucket (no zero-fill + add pass per parameter). */
int seldq_conv_wgrad(const seldq_conv_desc_t* d, const float* x, const void* x_cl,
                     const float* gy, const void* gy_t16, float* const* host_gw, float* gbias,
                     int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* The weight gradients of TWO convolutions of equal geometry that read the same x (the filter / gate and the
 * skip / residual convolutions of a residual block, model.py:118-132) in one launch; bf16 path, both operands
 * pre-staged (x as CL operand, the gy's as T16 operands). */
int seldq_conv_wgrad_pair(const seldq_conv_desc_t* d, const void* x_cl, const void* gy_t16_a, const void* gy_t16_b,
                          float* const* host_gw_a, float* const* host_gw_b, int32_t accumulate, void* stream);

/* Fused glue of a convolution's fp32 output (bf16 path, stride 1), replacing the element-wise passes model.py
 * runs behind conv2_skip / conv2_residual and in front of the BatchNorms (model.py:114-132, :210-216):
 *   mode 0   y  = conv(x)
 *   mode 1   y += conv(x)              the running sum of the blocks' skip outputs (model.py:210-216)
 *   mode 2   y  = addend + conv(x)     x + conv2_residual(y) (model.py:132); addend has y's layout, may not alias y
 *   stats    when not NULL: per-channel (sum, sum of squares) of the values STORED in y are ADDED to
 *            stats[2 c], stats[2 c + 1] (doubles, zeroed by the caller) -- the batch statistics of the BatchNorm
 *            that follows (batch_filter2 / batch_gate2 / the next block's batch_filter1), so no pass re-reads y */
typedef struct {
  int32_t mode;
  const float* addend;
  double* stats;
} seldq_conv_epilogue_t;

/* one convolution pass (SELDQ_PASS_FWD | SELDQ_PASS_DGRAD) from pre-staged operands with a fused epilogue */
int seldq_conv_epi(const seldq_conv_desc_t* d, int32_t pass, const void* in_cl, const void* packed_w, float* out,
                   const seldq_conv_epilogue_t* epi, void* stream);

/* TWO sibling convolutions of equal geometry in ONE launch: conv1_filter / conv1_gate (same x) or conv2_skip /
 * conv2_residual (same y) of a residual block (model.py:118-119, :130-131), forward or dgrad (then the inputs are
 * the two output gradients).  At the reference's batch size of 1 a TCN launch holds ~2 us of tensor-core work in
 * ~10 us of fixed cost; sharing the launch halves that cost.  Operands pre-staged (CL operands, packed weights of
 * the same pass), fp32 outputs, epilogues as above (epi_a / epi_b may be NULL).  Needs >= 8 channels per component
 * on the K side (seldq_conv_pair_supported returns 1). */
int seldq_conv_pair_supported(const seldq_conv_desc_t* d, int32_t pass);
int seldq_conv_pair(const seldq_conv_desc_t* d, int32_t pass, const void* in_cl_a, const void* in_cl_b,
                    const void* packed_a, const void* packed_b, float* out_a, float* out_b,
                    const seldq_conv_epilogue_t* epi_a, const seldq_conv_epilogue_t* epi_b, void* stream);

/* ---- glue between the convolutions (E1 in SURVEY.md 8a) -------------------------------------------------
 * CNN block, model.py:276-283: conv -> BatchNorm2d(train) -> ReLU -> MaxPool2d([pool,1]) -> Dropout(drop_p).
 *   seldq_bn_stats      sums[c][0] += sum v, sums[c][1] += sum v^2 over an NCHW tensor (plane = H*W elements per
 *                       (n, c)); the caller zeroes sums
 *   seldq_bn_finalize   coef[c] = {a, b, mean, rstd}: BN(v) = a v + b (batch statistics, biased variance, eps);
 *                       running_mean / running_var (may be NULL) get nn.BatchNorm's momentum update
 *   seldq_cnn_tail_fwd  z = dropout(max_{pool rows}(relu(BN(y)))) written as the channels-last bf16 operand of
 *                       the consuming convolution (z_cl, may be NULL) and / or as fp32 NCHW (z_f32, may be NULL);
 *                       idx gets one byte per pooled element (arg-max row | 0x80 if kept) for the backward pass;
 *                       ymax_f16 (pooled shape, may be NULL) gets the conv output at the arg-max, which lets the
 *                       backward reductions stream it instead of gathering from y.
 *                       seed: device counter the caller advances every step (needed iff drop_p > 0)
 *   seldq_cnn_tail_bwd  d(conv out) from gz (fp32, pooled NCHW): written as the pitched NCHW bf16 operand
 *                       (d_t16) and / or the channels-last operand (d_cl) of the producing convolution's
 *                       gradient kernels; dsums: 3*c doubles, zeroed by the caller; on return dsums[2c], dsums[2c+1] =
 *                       sum dy, sum dy*xhat = (d beta, d gamma) of channel c (the last c doubles are scratch) */
typedef struct {
  int32_t n, c, h, w;     /* conv output (N, C, H, W) */
  int32_t pool;           /* rows pooled along h (1..8); the pooled height is h / pool */
  float drop_p;
  uint32_t salt;          /* distinguishes the dropout streams of different layers */
} seldq_cnn_tail_desc_t;
int seldq_bn_stats(const void* src, int32_t is_f16, int32_t n, int32_t c, int64_t plane, double* sums, void* stream);
int seldq_bn_finalize(const double* sums, const float* gamma, const float* beta, int32_t c, double count, float eps,
                      float momentum, float* running_mean, float* running_var, float* coef, void* stream);
int seldq_cnn_tail_fwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* consumer, const void* y_f16,
                       const float* coef, const int64_t* seed, void* z_cl, float* z_f32, uint8_t* idx, void* ymax_f16,
                       void* stream);
int seldq_cnn_tail_bwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* producer, const void* y_f16,
                       const float* coef, const uint8_t* idx, const void* ymax_f16, const float* gz, double* dsums,
                       void* d_t16, void* d_cl, void* stream);

/* First CNN block, backward.  The first convolution needs no input gradient, so d(conv out) has a single consumer,
 * the layer's weight gradient: seldq_cnn_first_bwd runs the two BatchNorm reductions (as seldq_cnn_tail_bwd) and
 * then ONE kernel that forms d(conv out) tile by tile in shared memory and feeds it to the tensor cores, so the
 * largest gradient tensor of the model (472 MB per sample) is never written.  x is the float32 input of the
 * convolution; host_gw / accumulate as in seldq_conv_wgrad; dsums as in seldq_cnn_tail_bwd.  Geometry: stride 1,
 * input channels a multiple of 8 with taps * channels <= 128, at most 256 output channels, output width a multiple
 * of 8 (seldq_cnn_first_bwd_supported returns 1). */
int seldq_cnn_first_bwd_supported(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* conv);
size_t seldq_cnn_first_bwd_workspace_bytes(const seldq_conv_desc_t* conv);
int seldq_cnn_first_bwd(const seldq_cnn_tail_desc_t* t, const seldq_conv_desc_t* conv, const float* x,
                        const void* y_f16, const float* coef, const uint8_t* idx, const void* ymax_f16,
                        const float* gz, double* dsums, float* const* host_gw, int32_t accumulate, void* workspace,
                        size_t workspace_bytes, void* stream);

/* TCN residual block, model.py:109-132:
 *     x = tanh(BN1(r));  y = dropout1d(tanh(BN_f(conv_f x)) * sigmoid(BN_g(conv_g x)));
 *     r' = x + conv_res(y);  skips += conv_skip(y)
 * The four convolutions are the entry points above; seldq_tcn_glue runs the bandwidth-bound steps between them.
 * All tensors are fp32 (N, C, T); T must be a multiple of 8.  BatchNorm is train-mode: statistics travel as
 * per-channel (sum, sum of squares) in double and every step derives mean / rstd from them itself.  bf16 operand
 * outputs use the layouts of `layout_of`'s x (which = 0) or gy (which = 1) operand (seldq_conv_operand_info).
 *   op                     in[0..4]                      writes
 *   PREACT_FWD             r                             out32 = x, out_cl[0] = x operand; bn[0] running stats
 *   ROW_STATS (flag = k)   k <= 2 tensors                stats_out[i] += (sum, sum sq) of in[i]  (caller zeroes)
 *   GATE_FWD               y_f, y_g                      out_cl[0] = y operand; bn[0], bn[1] running stats
 *   RESIDUAL_FWD           x, conv_res(y) | NULL, conv_skip(y)
 *                                                        out32 = r', dsums[c] += (sum, sum sq) of r' (the next
 *                                                        block's BN1 statistics; may be NULL),
 *                                                        accum = (flag ? 0 : accum) + in[2]  (c2 channels)
 *   GATE_BWD_REDUCE        y_f, y_g, gy1, gy2 | NULL     dsums[0..3][c] += sum dz_f, sum dz_f xhat_f, sum dz_g, sum dz_g xhat_g
 *                                                        (= d beta_f, d gamma_f, d beta_g, d gamma_g)
 *   GATE_BWD_APPLY         same                          d y_f -> out_cl[0], out_t16[0];  d y_g -> out_cl[1], out_t16[1]
 *   PREACT_BWD_REDUCE      g_r' | NULL, gx1, gx2, x, r   dsums[0..1][c] += sum dz, sum dz xhat (= d beta, d gamma of BN1)
 *   PREACT_BWD_APPLY       same                          out32 = g_r; out_cl[0], out_t16[0] (may be NULL) = its operands
 * gy = gy1 + gy2 is the gradient w.r.t. y (two dgrad outputs), gx1 + gx2 the one w.r.t. x. */
enum {
  SELDQ_TCN_PREACT_FWD = 0, SELDQ_TCN_ROW_STATS = 1, SELDQ_TCN_GATE_FWD = 2, SELDQ_TCN_RESIDUAL_FWD = 3,
  SELDQ_TCN_GATE_BWD_REDUCE = 4, SELDQ_TCN_GATE_BWD_APPLY = 5, SELDQ_TCN_PREACT_BWD_REDUCE = 6,
  SELDQ_TCN_PREACT_BWD_APPLY = 7,
  /* single-launch forms (one block per 64-channel x 64-t tile, reduce and apply on either side of a grid barrier;
   * `sync` = a zeroed uint32 the call may use once; seldq_tcn_glue_fused_supported tells whether the tensor's tiles
   * can all be resident at once -- the call fails with SELDQ_ERR_UNSUPPORTED otherwise):
   *   GATE_BWD            = GATE_BWD_REDUCE + GATE_BWD_APPLY      (same arguments)
   *   PREACT_BWD          = PREACT_BWD_REDUCE + PREACT_BWD_APPLY  (same arguments)
   *   GATE_FWD_STATS      = ROW_STATS(y_f, y_g) + GATE_FWD: stats_out[i] += statistics of in[i] (caller zeroes;
   *                         bn[i].sums is ignored, the step reads stats_out[i]), then GATE_FWD
   *   RESIDUAL_PREACT_FWD = RESIDUAL_FWD (c2 == c, in[1] and in[2] not NULL, dsums not NULL) + the NEXT block's
   *                         PREACT_FWD: bn[0] = the next block's batch_filter1 (its sums are dsums),
   *                         out32b = tanh(BN(r')), out_cl[0] = its operand */
  SELDQ_TCN_GATE_BWD = 8, SELDQ_TCN_PREACT_BWD = 9, SELDQ_TCN_GATE_FWD_STATS = 10, SELDQ_TCN_RESIDUAL_PREACT_FWD = 11
};
typedef struct {
  const double* sums;                 /* [C][2] batch sum, sum of squares */
  const float* gamma;
  const float* beta;
  float* running_mean;                /* nn.BatchNorm momentum update by the forward steps; may be NULL */
  float* running_var;
} seldq_bn_ref_t;
typedef struct {
  int32_t n, c, t;
  int32_t c2;                         /* channels of the skip tensors (RESIDUAL_FWD) */
  float eps, momentum, drop_p;
  uint32_t salt;                      /* dropout stream of this block */
  double count;                       /* N * T */
  seldq_bn_ref_t bn[2];
  const float* in[5];
  float* out32;
  void* out_cl[2];
  void* out_t16[2];
  double* dsums;
  double* stats_out[2];
  float* accum;
  const int64_t* seed;                /* device counter, advanced by the caller every step (iff drop_p > 0) */
  int32_t flag;
  uint32_t* sync;                     /* single-launch forms: zeroed grid-barrier counter */
  float* out32b;                      /* RESIDUAL_PREACT_FWD: tanh(BN(r')) */
} seldq_tcn_glue_t;
int seldq_tcn_glue(int32_t op, const seldq_tcn_glue_t* args, const seldq_conv_desc_t* layout_of, int32_t which,
                   void* stream);
int seldq_tcn_glue_fused_supported(int32_t n, int32_t c, int32_t t);

/* ---- linear (A3, A4) ------------------------------------------------------------------- */
size_t seldq_linear_workspace_bytes(const seldq_linear_desc_t* d, int32_t pass);
int seldq_linear_fwd(const seldq_linear_desc_t* d, const float* x, const float* const* host_w,
                     const float* bias, float* y, void* workspace, size_t workspace_bytes, void* stream);
int seldq_linear_dgrad(const seldq_linear_desc_t* d, const float* gy, const float* const* host_w,
                       float* gx, void* workspace, size_t workspace_bytes, void* stream);
int seldq_linear_wgrad(const seldq_linear_desc_t* d, const float* x, const float* gy,
                       float* const* host_gw, float* gbias, int32_t accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- helpers ------------------------------------------------------------------------------ */
/* dst_bf16[i] = bf16(src[i]) (round to nearest even) */
int seldq_cast_bf16(const float* src, void* dst_bf16, size_t n, void* stream);

/* ---- STFT front end (F1) ----------------------------------------------------------------- */
/* x: (n_signals, n_samples) float32.  out: (n_batch, (1+output_phase)*n_ch, n_bins, n_frames)
 * with n_signals = n_batch*n_ch, magnitude of channel c at plane c and phase at plane n_ch+c
 * (the reference's np.concatenate on axis -3).  Periodic Hamming window, zero boundary
 * extension of nperseg/2, zero tail padding to a whole hop, scaling 1/sum(w); bin 0 dropped if
 * cut_dc, last frame dropped if cut_last.  nperseg must be 512. */
int seldq_stft_shape(int64_t n_samples, int32_t nperseg, int32_t noverlap, int32_t cut_dc,
                     int32_t cut_last, int32_t* n_bins, int32_t* n_frames);
int seldq_stft_magphase(const float* x, int32_t n_batch, int32_t n_ch, int64_t n_samples,
                        int32_t nperseg, int32_t noverlap, int32_t cut_dc, int32_t output_phase,
                        int32_t cut_last, float* out, void* stream);

/* The same front end with the steps train.py runs around it fused in:
 *   input_int16   x is int16 PCM (value / 32768) instead of float32: half the bytes read
 *   mean, inv_std the data-set normalisation of train.py:374-408, per plane (0 magnitude, 1 phase):
 *                 out = (feature - mean) * inv_std   ({0, 0} and {1, 1}: none)
 *   stats         when not NULL: sum and sum of squares of the UN-normalised features per plane are added to
 *                 stats[2 plane], stats[2 plane + 1] (doubles, zeroed by the caller) -- what np.mean / np.std of
 *                 train.py reduce over the stored array; with out == NULL the pass only gathers them (no store) */
typedef struct {
  int32_t input_int16;
  float mean[2], inv_std[2];
  double* stats;
} seldq_stft_options_t;
int seldq_stft_features(const void* x, int32_t n_batch, int32_t n_ch, int64_t n_samples, int32_t nperseg,
                        int32_t noverlap, int32_t cut_dc, int32_t output_phase, int32_t cut_last,
                        const seldq_stft_options_t* opt, float* out, void* stream);

/* TC_Block tail, model.py:210-231: activation followed by nn.MaxPool1d(pool) (stride = pool, floor mode) in one kernel
 * per direction.  act: SELDQ_ACT_RELU (relu1 / relu2 + maxpool1 / maxpool2) or SELDQ_ACT_TANH (tanh + maxpool3).
 * x: float32 (rows = N * C, t); y, gy: (rows, t / pool); gx: (rows, t).  The backward pass re-derives the arg-max
 * from x (first maximum wins, as nn.MaxPool1d) and needs y for the activation's derivative. */
#define SELDQ_ACT_RELU 0
#define SELDQ_ACT_TANH 1
int seldq_act_pool1d_fwd(const float* x, int64_t rows, int32_t t, int32_t pool, int32_t act, float* y, void* stream);
int seldq_act_pool1d_bwd(const float* x, const float* y, const float* gy, int64_t rows, int32_t t, int32_t pool,
                         int32_t act, float* gx, void* stream);

/* Evaluation path: gen_submission_list_task2 (utility_functions.py:184-210) for a batch of clips.
 * sed: float32 (clips, frames, classes * overlaps), doa: float32 (clips, frames, classes * overlaps * 3).  A cell whose
 * SED output rounds to non-zero (round half to even, as np.round) yields the row
 * [frame, class, x * max_loc, y * max_loc, z * max_loc, overlap] (6 floats); rows: (clips, frames * classes * overlaps, 6)
 * capacity, filled per clip in the reference's order (frame, then cell); counts[clip] = rows written. */
int seldq_seld_events(const float* sed, const float* doa, int32_t clips, int32_t frames, int32_t classes,
                      int32_t overlaps, float max_loc, float* rows, int32_t* counts, void* stream);

/* Rotation variants of the quaternion layers: quaternion_conv_rotation (quaternion_ops.py:174-232),
 * quaternion_transpose_conv_rotation (:235-295) and quaternion_linear_rotation (:330-388) build ONE real weight of
 * nc x nc blocks from the compact tensors -- nc = 3, or 4 with quaternion_format (block row 0 and block column 0 zero) --
 * and run a plain convolution / matrix product on it.  host_w: host array of the 4 device pointers r, i, j, k, each
 * float32 (d0, d1, taps) contiguous (convolution: (out / nc', in / nc', k...); transposed convolution and linear:
 * (in / nc', out / nc', ...)).  out: float32 (nc d0, nc d1, taps), or its (nc d1, nc d0, taps) transpose with
 * transpose_out (the linear variant runs as a 1 x 1 convolution, whose weight is (out, in)); IEEE float32 in the
 * reference's order of operations (bit-identical to oracle/algebra.py rotation_weight in float32).  The contraction is seldq_conv_* with SELDQ_ALG_REAL on that weight.
 * _bwd: g_out = gradient of `out` (same layout) -> host_gw: the gradients of r, i, j, k (overwritten). */
int seldq_rotation_weight(const float* const* host_w, int64_t d0, int64_t d1, int64_t taps, int32_t quaternion_format,
                          int32_t transpose_out, float* out, void* stream);
int seldq_rotation_weight_bwd(const float* const* host_w, const float* g_out, int64_t d0, int64_t d1, int64_t taps,
                              int32_t quaternion_format, int32_t transpose_out, float* const* host_gw, void* stream);

/* Point-wise operators on quaternion-valued tensors read as float32 (outer, 4, m) -- component c of quaternion (o, x)
 * at o * 4 m + c * m + x, the get_r / get_i / get_j / get_k slices of dimension 1 of the reference:
 * hamilton_product (quaternion_ops.py:467-507 = dual_quaternion_ops.py:374-414), q_normalize and quaternion_exp
 * (dual_quaternion_ops.py:206-246), and what their gradients need.  b is ignored by the one-operand operators. */
typedef enum {
  SELDQ_QOP_HAMILTON = 0,         /* out = a (x) b                                                          */
  SELDQ_QOP_HAMILTON_CONJ_B = 1,  /* out = a (x) conj(b): gradient of the first factor, a = grad, b = second */
  SELDQ_QOP_HAMILTON_CONJ_A = 2,  /* out = conj(a) (x) b: gradient of the second factor, a = first, b = grad */
  SELDQ_QOP_NORMALIZE = 3,        /* out = a / sqrt(|a|^2 + 1e-4)                                            */
  SELDQ_QOP_NORMALIZE_BWD = 4,    /* a = input, b = grad of the output -> grad of the input                  */
  SELDQ_QOP_EXP = 5,              /* out = e^r (cos |v|', v / |v|' sin |v|'), |v|' = |v| + 1e-4              */
  SELDQ_QOP_EXP_BWD = 6           /* a = input, b = grad of the output -> grad of the input                  */
} seldq_qpointwise_op_t;
int seldq_quaternion_pointwise(int32_t op, const float* a, const float* b, float* out, int64_t outer, int64_t m,
                               void* stream);

/* The optimiser step of train.py:502-504, :560 -- torch.optim.Adam(lr, betas = (b1, b2), eps), no weight decay, no
 * amsgrad -- over flat float32 buffers of n elements (parameters, gradients, exp_avg, exp_avg_sq; 16-byte aligned).
 * step: device float holding the number of steps taken so far; incremented by the call (so the launch can be
 * captured in a CUDA graph).  The hyper-parameters arrive as the doubles PyTorch holds them in: 1 - b1, 1 - b2 and the
 * bias corrections are formed in double, as torch/optim/adam.py forms them, before anything is rounded to float. */
int seldq_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double b1,
                    double b2, double eps, float* step, void* stream);
/* The same update over a PART of the buffers (pointers already offset, 16-byte aligned): every part of one optimiser
 * step reads the same `step`; only the call with advance != 0 -- the last one -- increments it.  Lets a trainer update
 * the parameters whose gradients are complete while the rest of the backward pass still runs. */
int seldq_adam_step_part(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double b1,
                         double b2, double eps, float* step, int32_t advance, void* stream);

/* debug: a 64 x uint64 device buffer that CTA 0 of every later tensor-core convolution launch stamps with
 * %globaltimer values at its role hand-offs (tools/fprop_trace.py); NULL switches it off. */
int seldq_debug_fprop_trace(void* dev_buf);

/* ---- multi-head self-attention (model.py:12-51) ------------------------------------------------
 *     out = softmax(q k^T / sqrt(d)) v      per (sample, head), heads = E / d channel groups of the projections
 * q, k, v: float32 (N, E, S) -- the output layout of the 1x1 projections (model.py:34-36; channel e = head * d + i);
 * out: float32 (N, S, E) -- what fc_out consumes (model.py:48-49).  bf16 operands, fp32 accumulation and softmax
 * (the tensor-core mode, gated at rel 2e-2); the (N, heads, S, S) energy / attention tensors of the reference are
 * never materialised.  d in {16, 32, 48}, S a multiple of 8 (seldq_attention_supported returns 1).
 *   saved      seldq_attention_saved_bytes(): bf16 operand copies the forward makes and the backward re-reads
 *   lse        float32 (N * heads * S): row-wise log-sum-exp (log2 domain), written by fwd, read by bwd
 *   workspace  seldq_attention_bwd_workspace_bytes(): operand copies of d_out + the row sums of d_out * out
 * seldq_attention_bwd: d_out (N, S, E) -> dq, dk, dv (N, E, S). */
typedef struct {
  int32_t batch, heads, seq, head_dim;
} seldq_attention_desc_t;
int seldq_attention_supported(const seldq_attention_desc_t* d);
size_t seldq_attention_saved_bytes(const seldq_attention_desc_t* d);
size_t seldq_attention_bwd_workspace_bytes(const seldq_attention_desc_t* d);
int seldq_attention_fwd(const seldq_attention_desc_t* d, const float* q, const float* k, const float* v, float* out,
                        float* lse, void* saved, void* stream);
int seldq_attention_bwd(const seldq_attention_desc_t* d, const void* saved, const float* out, const float* lse,
                        const float* d_out, float* dq, float* dk, float* dv, void* workspace, size_t workspace_bytes,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELDQ_H_ */

/*
 * fac_b200.h -- C ABI of the B200-native PPG -> Mel -> WaveGlow inference path.
 *
 * The reference (guanlongzhao/fac-via-ppg) has no FFI layer: its boundary is two
 * Python methods, WaveGlow.infer (src/waveglow/glow.py:252-293) and
 * Tacotron2.inference (src/common/model.py:597-610).  The drop-in Python classes
 * of this repository keep those signatures and call the entry points below
 * through ctypes.  Conventions:
 *   - every pointer is a DEVICE pointer into caller-owned (PyTorch-owned) memory;
 *     nothing is allocated or freed across the boundary;
 *   - every call takes the cudaStream_t to launch on (as void*), is asynchronous,
 *     and returns 0 on success or a non-zero code; fac_last_error() gives text;
 *   - activations are "channels-last": element (b, t, c) of a (B, T, C) tensor is
 *     at ptr[b*T*C + t*C + c], so a time column is one contiguous vector;
 *   - not thread-safe per stream; safe across streams/devices.
 *
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 */
#ifndef FAC_B200_H
#define FAC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FAC_MAX_FLOWS 16
#define FAC_MAX_LAYERS 16

/* ---- bookkeeping ------------------------------------------------------- */
int fac_version(void);
const char* fac_last_error(void);
/* Number of kernels launched by this library since load / since the last reset
 * (bench.py reports it as gpu_launches). */
long long fac_launch_count(void);
void fac_reset_launch_count(void);
/* Launches replayed from a CUDA graph captured around the entry points do not pass through the library:
 * the caller adds them here (the count taken while capturing, once per replay). */
void fac_add_launch_count(long long n);

/* ---- generic implicit-GEMM Conv1d / Linear (fp32 FFMA) ----------------- */
/* One input of a K-concatenated implicit GEMM.  Output row m (time step) reads
 * input rows m + tap*dilation - center for tap = 0..taps-1; rows outside [0, T_src)
 * are zero (Conv1d zero padding).  Element (b, row, c) is at
 * ptr[b*batch_stride + row*row_stride + c*ch_stride].  channels % 8 == 0. */
typedef struct fac_conv_src {
  const float* ptr;
  long long batch_stride, row_stride, ch_stride;
  int channels, taps, dilation, center;
  int rows;            /* T_src: number of valid rows of this input */
  int _pad;
} fac_conv_src;

enum { FAC_ACT_NONE = 0, FAC_ACT_RELU = 1, FAC_ACT_TANH = 2 };
enum {
  FAC_EPI_LINEAR = 0,   /* out = act(acc + bias) [* mask] [+ residual]                       */
  FAC_EPI_GATE = 1,     /* packed columns (2c, 2c+1) = (tanh, sigmoid) pre-activations of
                           channel c; out[:, c] = tanh(.)*sigmoid(.)  (glow.py:33-40)        */
  FAC_EPI_RES_SKIP = 2  /* columns [0, n_split): out += v (residual stream, glow.py:166);
                           columns [n_split, N): out2 (+)= v (skip sum, glow.py:167-174)     */
};

typedef struct fac_conv_epilogue {
  int kind, act;
  float* out;            /* (B, T_out, N_out) channels-last with the strides below */
  long long out_batch_stride, out_row_stride;
  const float* mask;     /* optional multiplicative mask, same addressing as out   */
  const float* residual; /* optional additive input, same addressing as out        */
  float* out2;           /* FAC_EPI_RES_SKIP: skip accumulator (same strides)       */
  int n_split;           /* FAC_EPI_RES_SKIP: width of the residual part            */
  int accumulate_out2;   /* 0: out2 = v, 1: out2 += v                               */
  const int* row_lengths; /* FAC_EPI_LINEAR, optional [B]: utterance b has row_lengths[b] valid rows; rows beyond
                             are written as exact zeros, so the next Conv1d sees the zero padding a shorter
                             utterance processed alone would see (variable-length batches)                   */
} fac_conv_epilogue;

/* Replaces torch.nn.Conv1d / torch.nn.Linear as used through ConvNorm / LinearNorm
 * (reference src/common/layers.py:40-71) and the WN convolutions
 * (src/waveglow/glow.py:156-164).  w_packed is [K_total][N_pad] row-major with
 * K index = (source, tap, channel) and N_pad = N rounded up to 128 (zero filled);
 * bias is [N_pad] or NULL.  n_phases > 1 runs that many independent GEMMs that
 * differ only by w_packed += phase*w_phase_stride and out += phase*out_phase_stride
 * (used by the transposed-conv upsampler). */
int fac_conv_gemm_f32(const fac_conv_src* srcs, int n_srcs, const float* w_packed, const float* bias,
                      int B, int T_out, int N, const fac_conv_epilogue* epi,
                      int n_phases, long long w_phase_stride, long long out_phase_stride, void* stream);

/* ---- WaveGlow reverse flow --------------------------------------------- */
/* Weights of one WaveGlow after remove_weightnorm (glow.py:295-311), repacked by
 * fac_via_ppg_b200/packing.py.  All fp32 device pointers. */
typedef struct fac_wg_flow {
  int n_half, n_rem;                 /* coupling half width, channels alive in this flow (glow.py:195-207) */
  const float* start_w;              /* [n_half][C]  (transposed start.weight)                */
  const float* start_b;              /* [C]                                                   */
  const float* end_w;                /* [2*n_half][C]                                         */
  const float* end_b;                /* [2*n_half]                                            */
  const float* w_inv;                /* [n_rem][n_rem] = inverse of convinv weight (glow.py:89-95) */
  const float* in_cond_w[FAC_MAX_LAYERS]; /* [3*C + n_cond][2*C] gate-interleaved columns     */
  const float* in_cond_b[FAC_MAX_LAYERS]; /* [2*C] = in.bias + cond.bias, gate-interleaved    */
  const float* res_skip_w[FAC_MAX_LAYERS];/* [C][N_pad]                                       */
  const float* res_skip_b[FAC_MAX_LAYERS];/* [N_pad]                                          */
} fac_wg_flow;

typedef struct fac_wg_model {
  int n_flows, n_layers, n_channels, n_group, n_mel, hop, n_early_every, n_early_size, upsample_taps;
  int kernel_size;                   /* WN in_layers kernel size (3) */
  const float* upsample_w;           /* [hop/n_group phases][taps*n_mel][n_mel*n_group (padded to 128)] */
  const float* upsample_b;           /* [n_mel*n_group (padded)] bias replicated per group slot */
  fac_wg_flow flows[FAC_MAX_FLOWS];
} fac_wg_model;

/* Scratch the caller allocates for B utterances of F frames (T_g = F*hop/n_group columns):
 * spect (B,T_g,n_mel*n_group), x / acts / skip (B,T_g,C). */
typedef struct fac_wg_workspace {
  float* spect; float* x; float* acts; float* skip;
} fac_wg_workspace;

/* glow.py:253-259: ConvTranspose1d(n_mel,n_mel,1024,stride=hop) + trim + group squeeze,
 * written directly in squeezed channels-last form.  mel_cl is (B, F, n_mel). */
int fac_waveglow_upsample_squeeze_f32(const fac_wg_model* m, const float* mel_cl, float* spect,
                                      int B, int F, void* stream);
/* glow.py:156 : x = start(audio_0).  audio is (B, T_g, n_group) with the flow's live
 * channels in the LAST n_rem slots of each column. */
int fac_wn_start_f32(const fac_wg_model* m, int flow, const float* audio, float* x, int B, int Tg, void* stream);
/* glow.py:158-174, one WN layer: gate(in_layer(x) + cond_layer(spect)) -> acts;
 * res_skip(acts) -> x += res, skip (+)= skip. */
int fac_wn_layer_f32(const fac_wg_model* m, int flow, int layer, const fac_wg_workspace* ws,
                     int B, int Tg, void* stream);
/* glow.py:175 + 278-283: end conv, affine coupling inverse, invertible 1x1 (reverse), in place. */
int fac_wn_end_coupling_f32(const fac_wg_model* m, int flow, const float* skip, float* audio,
                            int B, int Tg, void* stream);
/* glow.py:252-293 without the RNG: `audio` (B, T_g, n_group) arrives holding sigma*z in
 * every slot (the reference's three normal_() draws laid out per slot, see
 * fac_via_ppg_b200/waveglow/glow.py) and leaves holding the waveform (B, T_g*n_group). */
int fac_waveglow_infer_f32(const fac_wg_model* m, const float* mel_cl, float* audio,
                           const fac_wg_workspace* ws, int B, int F, void* stream);

/* ---- WaveGlow WN layers on tcgen05 tensor cores -------------------------- */
/* Weights of the tensor-core path, derived from the fp32 packed weights by packing.py.
 * bf16 operand matrices are [N][K] row-major (K contiguous); *_hi = bf16(w), *_lo = bf16(w - hi).
 *   w1 = in_layer|cond_layer: N = 2C gate-interleaved like in_cond_w, K = 3C + n_cond ordered
 *        tap0|tap1|tap2|cond (glow.py:159-162)
 *   w2 = [residual half of res_skip | identity]: N = C, K = 2C, so that the tensor core itself
 *        performs x <- x + res (glow.py:164-166); lo has zeros in the identity half; unused for
 *        the last layer, whose res_skip feeds the skip path only (two-launch form of a layer)
 *   w2r = residual half of res_skip alone: N = C, K = C (fused one-launch form: the residual add happens in
 *        the epilogue of the second GEMM)
 *   wc = end.weight @ (skip half of res_skip): the skip path collapsed into an 8-channel update
 *        per layer (glow.py:167-175: end() is linear in the skip sum); fp32 [8][C], rows >= 2*n_half zero
 *   out_bias = end.bias + end.weight @ sum_i (skip half of res_skip bias_i); fp32 [8]
 *   res_b = residual half of the res_skip bias; fp32 [C] */
typedef struct fac_wg_tc_flow {
  const void* w1_hi[FAC_MAX_LAYERS]; const void* w1_lo[FAC_MAX_LAYERS];
  const void* w2_hi[FAC_MAX_LAYERS]; const void* w2_lo[FAC_MAX_LAYERS];
  const void* w2r_hi[FAC_MAX_LAYERS]; const void* w2r_lo[FAC_MAX_LAYERS];
  const float* wc[FAC_MAX_LAYERS];
  const float* res_b[FAC_MAX_LAYERS];
  const float* out_bias;
} fac_wg_tc_flow;
typedef struct fac_wg_tc_weights {
  /* upsampler (glow.py:253) as hop/n_group phase GEMMs: [phases][n_mel*n_group][taps*mel_pad] bf16, K ordered
   * tap-major with the mel channels zero-padded to mel_pad (a multiple of 32) */
  const void* up_hi; const void* up_lo;
  int mel_pad, _pad;
  fac_wg_tc_flow flows[FAC_MAX_FLOWS];
} fac_wg_tc_weights;

/* Scratch of the tensor-core path for B utterances of T_g columns: the padded mel copies, the
 * bf16 hi/lo operand copies the TMA loads read ((B,T_g,channels) channels-last; the residual
 * stream x lives ONLY as its hi+lo pair) and out8 (B,T_g,8) fp32, the running end() pre-activation.
 * The *_lo buffers may be NULL when nsplit == 1.
 * x2 (optional): a second residual-stream pair.  When present (and nsplit == 2) a layer runs as ONE fused launch
 * (csrc/waveglow_fused.cu: first GEMM -> gate -> residual GEMM -> residual add, acts never leaves the SM) that
 * reads the stream from one pair and writes the other, because neighbouring time tiles still read the old
 * values for their dilated taps: layer i reads x when i is even and x2 when i is odd, and writes the other one
 * (fac_wn_start_tc always writes x).  acts_hi/acts_lo may then be NULL (a test hook when not).
 * flow_sync (optional, with x2): 4 * B * ceil(T_g / 128) bytes, one counter per 128-column time tile, for
 * fac_waveglow_flow_step_tc, which then runs a whole flow step -- start, every WN layer, end + coupling +
 * invertible 1x1 -- as ONE cooperative launch in which a tile of a layer waits for its own and its two neighbour
 * tiles of the previous layer (dilated taps reach no further) instead of for the whole grid. */
typedef struct fac_wg_tc_workspace {
  void* mel_hi; void* mel_lo;          /* (B, F, mel_pad) bf16 */
  void* spect_hi; void* spect_lo;
  void* x_hi; void* x_lo;
  void* acts_hi; void* acts_lo;
  float* out8;
  void* x2_hi; void* x2_lo;
  void* flow_sync;
} fac_wg_tc_workspace;

/* nsplit = 1: bf16 operands; nsplit = 2: split-bf16 (3 UMMAs per product, fp32-grade result). */
/* glow.py:253-259 (upsample + trim + squeeze) on the tensor cores, from mel_cl (B, F, n_mel) fp32. */
int fac_waveglow_tc_prepare_spect(const fac_wg_model* m, const fac_wg_tc_weights* w, const fac_wg_tc_workspace* ws,
                                  const float* mel_cl, int B, int F, int nsplit, void* stream);
int fac_wn_start_tc(const fac_wg_model* m, int flow, const float* audio, const fac_wg_tc_workspace* ws,
                    int B, int Tg, int nsplit, void* stream);
/* glow.py:158-174 for one layer: TMA-fed tcgen05 GEMM + gate epilogue (+ out8 update) and the residual GEMM
 * (skipped for the last layer) -- one fused launch when the workspace carries x2 (see above), else two launches
 * with x updated in place. */
int fac_wn_layer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, int layer,
                    const fac_wg_tc_workspace* ws, int B, int Tg, int nsplit, void* stream);
/* glow.py:175 + 278-283 from out8: coupling inverse and invertible 1x1 (reverse), in place on audio. */
int fac_wn_end_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, const float* out8, float* audio,
                  int B, int Tg, void* stream);
/* (The kernels of a flow step -- start, the fused layers, end -- are chained by programmatic dependent launch: each
 * may start while its predecessor in the stream is still running and waits for it only before it touches the
 * residual stream, out8 or audio.  Environment FAC_TC_PDL=0 launches them fully serialised.) */
/* One step of the reverse flow (reference src/waveglow/glow.py:272-290 for flow k: WN.forward :154-175 on audio_0,
 * affine coupling inverse :278-281, Invertible1x1Conv reverse :283) in place on `audio` (B, T_g, n_group), whose
 * live channels are the last n_rem slots of every column.  With nsplit == 2 and a workspace that carries x2 and
 * flow_sync this is ONE kernel launch (csrc/waveglow_fused.cu: start -> 8 fused layers chained by per-tile
 * dependency counters, no grid barrier -> end / coupling / W^-1, weights TMA-streamed); otherwise it is the
 * sequence fac_wn_start_tc, fac_wn_layer_tc x n_layers, fac_wn_end_tc. */
int fac_waveglow_flow_step_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, int flow, float* audio,
                              const fac_wg_tc_workspace* ws, int B, int Tg, int nsplit, void* stream);
/* Optional cycle counters of the tensor-core kernel: device buffer of 2*256*8 int64 ([G1|G2][cta][8]:
 * producer wait-empty, MMA wait-tmem, MMA wait-full, MMA total, epilogue wait, epilogue busy); NULL disables. */
void fac_tc_set_profile_buffer(long long* device_buf);
/* 0 (default): automatic; 1: one CTA per 128-column tile; 2: CTA pairs (thread-block cluster of 2,
 * tcgen05 cta_group::2, UMMA M = 256, each CTA stages half of the weight rows). */
int fac_tc_set_cta_group(int cta_group);
/* 0: always the two-launch form of a layer; 1: one fused launch per layer; 2 (default): additionally one launch per
 * flow step where the workspace allows it (flow_sync); 3: like 2 with start and end as separate kernels (three
 * launches per flow step).  Bits 4-9: L2 prefetch distance of the fused kernel's producer in K
 * steps; bits 10-13: 1 + mask of its L2 eviction hints (A/B measurements). */
int fac_tc_set_fused(int enabled);
/* Utterances per pass of fac_waveglow_infer_tc over a flow (they are independent): 0 (default) = the whole
 * batch; a group whose residual stream and gated activations fit the L2 keeps them out of HBM between the
 * GEMMs of a layer. */
int fac_tc_set_batch_group(int utterances);
/* K elements per TMA/UMMA pipeline stage: 0 (default) automatic, 32 (SWIZZLE_64B rows) or 64 (SWIZZLE_128B rows). */
int fac_tc_set_k_block(int k_block);
/* Same contract as fac_waveglow_infer_f32 (glow.py:252-293), WN layers on the tensor cores. */
int fac_waveglow_infer_tc(const fac_wg_model* m, const fac_wg_tc_weights* w, const float* mel_cl, float* audio,
                          const fac_wg_tc_workspace* ws, int B, int F, int nsplit, void* stream);

/* ---- generic Conv1d / Linear on the tcgen05 tensor cores ------------------ */
/* Same role as fac_conv_gemm_f32 (torch.nn.Conv1d / torch.nn.Linear behind ConvNorm / LinearNorm,
 * reference src/common/layers.py:40-71; used by the encoder and the postnet, model.py:178-184, 237-249)
 * with bf16 operands: the input is a channels-last (B, T, c_pad) bf16 tensor (a_hi [+ a_lo]), c_pad a
 * multiple of 64 with zero padding channels; output row t reads input rows t + tap - center, rows outside
 * [0, T) are zero (Conv1d padding).  Weights are [n_pad][taps*c_pad] bf16 (K contiguous, tap-major),
 * n_pad a multiple of 64, rows >= n_valid zero.  nsplit = 1: single 16-bit operands; nsplit = 2: split
 * operands (x = hi + lo, 3 UMMAs per product): ~2^-16 relative with bf16, ~2^-21 with IEEE half.  Epilogue: v = act(acc + bias) [* mask] [+ residual];
 * bias is [n_pad] or NULL; mask / residual / out are fp32 with n_valid columns (n_valid % 4 == 0) and the
 * given row strides; out_hi/out_lo (optional) receive the (B, T, n_pad) bf16 operand copies for the next
 * layer (padding columns exact zeros). */
typedef struct fac_tc_conv {
  const void* a_hi; const void* a_lo;
  const void* w_hi; const void* w_lo;
  const float* bias;
  const float* mask; const float* residual;
  float* out;
  void* out_hi; void* out_lo;
  long long mask_ld, res_ld, out_ld;
  int B, T, c_pad, taps, center, n_pad, n_valid, act, nsplit;
  int fp16;   /* 0: operands are bf16; 1: IEEE half (hi + lo = 22 significand bits; needs |x| < 65504) */
  /* K-chunked accumulation: the tensor core's fp32 accumulator truncates on every accumulation; with
   * k_chunk > 0 (a multiple of 64) every output tile walks the contraction in chains of <= k_chunk elements,
   * each in its own TMEM accumulator, whose partial sums meet in fp32 round-to-nearest through scratch, a
   * (B*T, n_valid) fp32 buffer.  0 = one chain. */
  int k_chunk, _pad;
  float* scratch;
  const int* row_lengths;  /* optional [B]: rows t >= row_lengths[b] of utterance b are written as exact zeros
                              (fp32 and 16-bit outputs): per-utterance Conv1d zero padding in a ragged batch */
} fac_tc_conv;
int fac_conv_gemm_tc(const fac_tc_conv* conv, void* stream);
/* (B, C, T) channel-major fp32 -> (B, T, pad) channels-last 16-bit hi [+ lo] (bf16, or IEEE half when fp16 != 0),
 * channels [C, pad) zero. */
int fac_transpose_split_16(const float* in, void* hi, void* lo, int B, int C, int T, int pad, int fp16, void* stream);
/* (n_rows, C) fp32 -> (n_rows, pad) 16-bit hi [+ lo], columns [C, pad) zero. */
int fac_pad_split_16(const float* in, void* hi, void* lo, long long n_rows, int C, int pad, int fp16, void* stream);

/* ---- pruned posteriorgram input -------------- */
/* The step before the path (reference src/common/data_utils.py:55-59 get_ppg) hands Tacotron2.inference a dense
 * (T, 5816) posteriorgram whose frames are almost entirely tail.  fac_ppg_sparsify turns the channel-major
 * (B, D, T) input into per-frame lists of the entries > threshold: idx / val are (B, T, k), ascending channel
 * order, padded with (0, 0.0); k <= 64.  *overflow (one int, zeroed by the caller) counts the frames that had
 * more than k survivors -- those lists are truncated and the caller must not use them. */
int fac_ppg_sparsify(const float* ppg, int* idx, float* val, int* overflow, int B, int D, int T, int k,
                     float threshold, void* stream);
/* First encoder prenet layer (reference src/common/model.py:124-135: bias-free Linear D -> E, ReLU, dropout mask)
 * on such lists: out[b,t,:] = relu(sum_j val[b,t,j] * w_t[idx[b,t,j], :]) * mask[b,t,:], exact fp32 FMA in list
 * order.  w_t is [D][w_ld] (the transposed weight, fac_via_ppg_b200/packing.py enc.pre0_w); mask (B,T,E) or NULL;
 * row_lengths as in fac_tc_conv; out (fp32, row stride out_ld) and/or out_hi/out_lo ((B,T,pad) IEEE-half hi/lo
 * operand copies for fac_conv_gemm_tc, channels [E, pad) zero). */
int fac_prenet0_sparse_f32(const int* idx, const float* val, const float* w_t, int w_ld, const float* mask,
                           const int* row_lengths, float* out, int out_ld, void* out_hi, void* out_lo, int B, int T,
                           int k, int D, int E, int pad, void* stream);

/* ---- PPG -> Mel (Tacotron2 variant) ------------------------------------- */
/* Recurrent part of the encoder's bidirectional LSTM (reference src/common/model.py:211-213,
 * 246-247).  xp is (B, T, 2*4H): x W_ih^T + b_ih + b_hh of the forward direction in columns
 * [0, 4H) and of the reverse direction in [4H, 8H) (gate order i, f, g, o), produced by
 * fac_conv_gemm_f32.  w_hh is [2][4H][H] (weight_hh_l0, weight_hh_l0_reverse).  out is
 * (B, T, 2H) = [forward h | reverse h], i.e. the encoder `memory`. */
int fac_lstm_bidir_f32(const float* xp, const float* w_hh, float* out, int B, int T, int H, void* stream);
/* Same for a ragged batch (the batched form of Tacotron2.inference; the reference builds input_lengths at
 * src/common/model.py:599 and packs the sequences for nn.LSTM in Encoder.forward :225-233): utterance b has
 * lengths[b] <= T valid rows, its reverse direction starts at row lengths[b]-1, and rows >= lengths[b] of `out`
 * are left untouched (the caller zero-fills them).  lengths == NULL means T for everybody. */
int fac_lstm_bidir_var_f32(const float* xp, const float* w_hh, float* out, const int* lengths, int B, int T, int H,
                           void* stream);

/* Optional cycle counters of the BiLSTM kernel: device buffer of grid*4 int64 per CTA (tensor-core gate mat-vec, cell
 * update + DSMEM hand-over, cluster barrier); NULL disables. */
void fac_lstm_set_profile_buffer(long long* device_buf);

/* Decoder weights (fp32 device pointers), packed by fac_via_ppg_b200/packing.py from the
 * reference state-dict keys decoder.* (SURVEY.md section 8a). */
typedef struct fac_taco_decoder_weights {
  const float* w_att;    /* [1200][1200] attention_rnn: cat(weight_ih (prenet 300 | context 600), weight_hh 300) */
  const float* b_att;    /* [1200] bias_ih + bias_hh                                                            */
  const float* w_dec;    /* [1200][1200] decoder_rnn: cat(weight_ih (h_att 300 | context 600), weight_hh 300)   */
  const float* b_dec;    /* [1200]                                                                              */
  const float* wq;       /* [150][300]  attention_layer.query_layer weight                                      */
  const float* w_loc;    /* [2][31][32] attention_layer.location_layer.location_conv weight, filter index last  */
  const float* w_ld_t;   /* [32][150]   location_dense weight, transposed                                       */
  const float* v;        /* [150]       attention_layer.v weight                                                */
  const float* w_pp;     /* [381][900]  rows 0..79 linear_projection, row 80 gate_layer, rows 81..380 =
                                        prenet.layers.0 @ linear_projection (no bias / nonlinearity sits between
                                        the projection and the first prenet layer, model.py:132-135, 436-438)   */
  const float* b_pp;     /* [381]       projection bias, gate bias, prenet.layers.0 @ projection bias           */
  const float* w_pre2;   /* [300][300]  decoder.prenet.layers.1 weight                                          */
} fac_taco_decoder_weights;

#define FAC_TACO_XCHG_WORDS 1800   /* 4 x 300 + 600 words per utterance and parity */
#define FAC_TACO_XCHG_HINTS 64     /* arrival counters behind the two copies */

/* Decoder state the caller allocates ZERO-FILLED (reference model.py:304-335 initialises every
 * state to zero); the kernel owns it while running. */
typedef struct fac_taco_decoder_state {
  float* c_att;   /* [B][300]    attention_cell                              */
  float* c_dec;   /* [B][300]    decoder_cell                                */
  unsigned long long* xchg; /* 2 * FAC_TACO_XCHG_WORDS * B + FAC_TACO_XCHG_HINTS words, exchange area: every vector that crosses CTAs (prenet
                     output, attention_hidden, decoder_hidden, prenet layer-0 output, attention_context) as
                     (value, version) 8-byte words, two copies by version parity; producers publish with ONE
                     store, consumers poll for the version they expect -- the kernel has no grid barrier */
  float* w_prev;  /* [B][T_in]   attention_weights                           */
  float* w_cum;   /* [B][T_in]   attention_weights_cum                       */
  int* done;      /* [8]: #utterances stopped by the gate, #stopped by max_steps, steps run, spare, [4..5] the
                     (count, version) word that carries the stop count of a step, spare, [7] != 0: a hand-over
                     timed out (the run is invalid; fac_taco_decoder_run's caller must check it) */
  int* out_len;   /* [B] number of frames of each utterance (0 while running) */
  /* optional sparse form of the alignments (NULL = off): only the <= 2*window+1 positions of the attention window
   * carry weight (everything else is masked to -inf by reference utils.py:46-78), so step t of utterance b is
   * align_win[b][t][0 .. 2*window] (zero-padded) starting at input position align_start[b][t] -- 41 floats per
   * step instead of T_in (273 MB per 60 s utterance dense) */
  float* align_win;   /* (B, max_steps, 2*window+1) pre-zeroed */
  int* align_start;   /* (B, max_steps) */
} fac_taco_decoder_state;

/* Diagnostic: runs `iters` grid-wide barriers (release-add + acquire spin, what round 1's decoder used between
 * its phases) on a full cooperative grid; time the launch to get the per-barrier latency that the dataflow decoder
 * avoids.  `zeroed_counter` is one zero-initialised uint32. */
int fac_selftest_grid_barrier(unsigned int* zeroed_counter, int iters, void* stream);

/* Optional cycle counters of the decoder kernel: device buffer of grid*32 int64 per CTA (slot meanings per role:
 * tools/decoder_cycle_breakdown.py; [10] = total); NULL disables. */
void fac_taco_set_profile_buffer(long long* device_buf);

/* The whole autoregressive loop of Decoder.inference (reference model.py:489-535 with decode
 * :387-442, Attention :100-121, window mask utils.py:46-78) in one persistent kernel.
 *   memory (B,T_in,600), pmem = memory_layer(memory) (B,T_in,150), lengths[B] (int32),
 *   drop (max_steps, 2, B, 300) uint8 in {0,1}: the always-on prenet dropout masks
 *   (model.py:132-135) of step t, layers 0/1;  outputs mel (B,max_steps,80), gate (B,max_steps),
 *   align (B,max_steps,T_in) pre-zeroed or NULL.  B <= (number of SMs - 100): one CTA per
 *   utterance runs the attention, the others hold the rows of one matrix each and exchange vectors as
 *   (value, version) words (no grid barrier; state->xchg).  Stops when every utterance's
 *   sigmoid(gate) > gate_threshold has fired (model.py:524) or at max_steps (model.py:526-528).
 *   After the kernel, state->done[7] != 0 means a hand-over timed out (the GPU was shared with a kernel
 *   that kept CTAs of this one from running, or the state was not zero-filled): the outputs are invalid. */
int fac_taco_decoder_run(const fac_taco_decoder_weights* w, const float* memory, const float* pmem,
                         const int* lengths, const unsigned char* drop, const fac_taco_decoder_state* state,
                         float* mel, float* gate, float* align, int B, int T_in, int max_steps,
                         int window, float gate_threshold, void* stream);

/* ---- Denoiser (the step after WaveGlow.infer on the CLI path) ------------- */
/* reference src/waveglow/denoiser.py:63-68: in place on an STFT spectrum laid out as rows of
 * [n_bins real | n_bins imaginary | padding] with leading dimension ld: magnitude <- max(magnitude -
 * bias_mag[bin]*strength, 0), phase kept.  The STFT and its inverse (reference src/common/stft.py:79-138,
 * a dense-DFT Conv1d / ConvTranspose1d) are fac_conv_gemm_f32 calls on the hop-reshaped signal. */
int fac_denoise_spectrum_f32(float* spec, const float* bias_mag, float strength, long long n_rows, int n_bins,
                             int ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FAC_B200_H */

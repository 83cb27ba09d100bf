/* libdesta_b200.so — C ABI of the B200-native waveform -> LLM-embedding front-end for DeSTA2.5-Audio.
 *
 * Every entry point takes plain device pointers + sizes + a CUDA stream (passed as void*), launches
 * asynchronously on that stream, allocates nothing, and returns 0 or a negative DESTA_ERR_* code
 * (message via desta_last_error()).  There is no CPU fallback.
 *
 * Each function cites the reference interface it replaces (paths relative to voidful/DeSTA2.5-Audio;
 * $TF = the pinned third-party `transformers` package holding the arithmetic).
 */
#ifndef DESTA_B200_H
#define DESTA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DESTA_OK 0
#define DESTA_ERR_ARG (-1)
#define DESTA_ERR_CUDA (-2)
#define DESTA_ERR_SHAPE (-3)

int desta_version(void);
const char* desta_last_error(void);

/* ---- (1) log-mel: replaces WhisperFeatureExtractor.__call__ -> _torch_extract_fbank_features
 *      ($TF/models/whisper/feature_extraction_whisper.py:135-164, 189-342), called at
 *      desta/models/modeling_desta25.py:1570,1583 and desta/trainer/data/simple_dataset.py:239-243. ---- */
#define DESTA_MEL_SAMPLES 480000
#define DESTA_MEL_FRAMES 3000
/* Host-side plan (Hann window, FFT twiddles, sparse slaney mel filters built in fp64, cast fp32).
 * Call with host_buf == NULL to get the size; the caller copies the filled buffer to the device. */
size_t desta_logmel_plan_bytes(int n_mels);
int desta_logmel_plan_build(int n_mels, void* host_buf, size_t bytes);
/* wave [N,480000] f32 (already right-padded/truncated; n_valid[N] optional: samples >= n_valid[i] read as 0).
 * plan = device copy of the plan buffer.  clip_max [N] 32-bit words, written by the call: the per-clip maximum of
 * log10(mel) in the library's order-preserving encoding.
 *   out_f32 != NULL: out_f32 [N,n_mels,3000] (reference layout) and, if given, out_bf16_tm [N,3000,n_mels] (time-major)
 *     receive the final values  (max(L, Lmax - 8) + 4) / 4   (two launches).
 *   out_f32 == NULL (product path, ONE launch, algorithmic HBM traffic): out_bf16_tm receives (L + 4) / 4 WITHOUT the
 *     per-clip floor.  The floor commutes with the monotone affine map and with bf16 rounding, so the consumer applies
 *     max(v, bf16((Lmax - 8 + 4) / 4)) on read: pass clip_max as `mel_clip_max` to desta_im2col_k3 / desta_encoder_fwd. */
int desta_logmel(const float* wave, const int32_t* n_valid, int N, int n_mels, const void* plan,
                 float* out_f32, void* out_bf16_tm, float* clip_max, void* stream);

/* ---- building blocks of (2)/(3): tcgen05 GEMM with fused epilogue.  C = epi(A[M,K] W[N,K]^T), bf16 in, fp32 acc.
 *      Replaces nn.Linear / Conv1d-as-GEMM calls of $TF/models/whisper/modeling_whisper.py:310-355,404-408,
 *      $TF/models/bert/modeling_bert.py:168-207,258-259,294-298,339-356, modeling_desta25.py:563-570,606. ---- */
typedef struct desta_gemm_epilogue {
  float* out_f32;  int ld_f32;          /* optional fp32 output */
  void* out_bf16;  int ld_bf16;         /* optional bf16 output */
  const float* bias;                    /* [N] or NULL */
  const float* resid; int ld_resid;     /* fp32 residual added last, or NULL */
  const float* pos;   int pos_rows;     /* fp32 [pos_rows,N] table, row (m % pos_rows) added after activation */
  float scale;        int scale_cols;   /* columns < scale_cols are multiplied by scale after the bias */
  int gelu;                             /* erf-GELU after bias/scale */
  /* LayerNorm folded into the GEMMs on either side of it (M > 128 only; the encoder's pre-LN layers,
   * $TF/models/whisper/modeling_whisper.py:393-408, lose their stand-alone LayerNorm pass):
   *   producer (fp32 output mode): ln_stats_out [M, ceil(N/128), 2] fp32 receives, per final output row and 128-column
   *     group, the partial (sum, sum of squares); out_bf16 receives the bf16 copy of the same rows.
   *   consumer (bf16 output mode): A is that bf16 copy, W holds gain[k] * W[n,k], bias holds sum_k beta[k] W[n,k] + b[n],
   *     ln_colsum[n] = sum_k of the bf16 W row; the epilogue computes rstd * (acc - mean * ln_colsum[n]) + bias[n] from
   *     ln_stats_in [M, ln_groups, 2] (mean / variance over ln_cols columns, eps = ln_eps) before scale / GELU. */
  float* ln_stats_out;
  const float* ln_stats_in; int ln_groups; int ln_cols; float ln_eps;
  const float* ln_colsum;
} desta_gemm_epilogue;
int desta_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                    const desta_gemm_epilogue* ep, void* stream);

/* Counter-based dropout (Philox-4x32-7) for the Q-Former in train mode.  keep iff rnd16 >= thr16
 * (thr16 = round(p * 65536); 0 = off); kept values are multiplied by scale = 1/(1-p).  (seed, site, element index) fully
 * determine the mask, so the backward regenerates it instead of storing it.  One Philox call covers an aligned group
 * of 8 consecutive elements of a row; element grid: [M, N] tensors -> group = row * (N/8) + col/8 (N % 8 == 0);
 * attention probabilities -> row = (b*heads + h)*Tq + t, group = row * ceil(Tk/8) + key/8.
 * Reference: BertConfig hidden_dropout_prob / attention_probs_dropout_prob = 0.1, active in train mode
 * ($TF/models/bert/modeling_bert.py:198-201 probabilities, :296, :354 dense outputs). */
typedef struct desta_dropout {
  uint32_t thr16;
  float scale;
  uint64_t seed;
  uint32_t site;       /* which dropout call of the step (layer / position) */
  float* row_keep_ratio; /* optional, desta_attention only: [batch,heads,Tq] <- (sum_k dropped, rescaled P) / (sum_k P) per query
                          * row.  With it the value-bias gradient of the attention is  sum_q ratio_q dO_q  (desta_colsum_heads)
                          * instead of a column sum over dV. */
} desta_dropout;

/* LayerNorm over the last dim (fp32 in; fp32 statistics).  out_bf16 and/or out_f32.
 * Replaces nn.LayerNorm at modeling_whisper.py:393,403 (eps 1e-5), modeling_bert.py:297,355 (eps 1e-12),
 * modeling_desta25.py:166 (proj.0). */
int desta_layernorm(const float* x, int ldx, const float* gamma, const float* beta, float eps, int M, int D,
                    void* out_bf16, int ld_bf16, float* out_f32, int ld_f32, void* stream);

/* Non-causal multi-head attention  softmax(scale * Q K^T) V, head_dim 64, bf16 in/out, fp32 softmax.
 * Q rows:  q + (b*Tq + t)*ldq + h*64 ; K/V likewise with Tk, ldk, ldv ; out + (b*Tq + t)*ldo + h*64.
 * Replaces WhisperAttention (modeling_whisper.py:284-358, scale 1 because q is pre-scaled) and the eager
 * Bert self/cross attention (modeling_bert.py:168-207, 234-284, scale head_dim^-1/2). */
int desta_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                    int batch, int heads, int Tq, int Tk, float scale, float* lse /* optional [batch,heads,Tq] */,
                    const desta_dropout* drop /* optional: dropout on the attention probabilities */, void* stream);

/* Kernel variant serving the encoder's calls of desta_attention (scale * log2 e == 1): 0 = four softmax warps, softmax shift
 * applied by the tensor cores (default); 1 = eight softmax warps with Q in tensor memory, one tile per CTA; 2 = the same,
 * persistent.  v < 0 only queries.  Returns the previous value.  All variants are tested against the same oracle; the
 * selector exists for A/B measurements (profiles/r2_attention_unit8.md).  Process-wide; not part of the reference's interface. */
int desta_attention_variant(int v);

/* im2col for the conv stem (modeling_desta25.py:563-564): Conv1d(k=3,pad=1,stride s) as GEMM rows.
 * x time-major bf16 [N, T_in, C] -> cols bf16 [N*T_out, 3*C] with column index = tap*C + c  (T_out = T_in/stride).
 * mel_clip_max: NULL, or the clip_max words of a desta_logmel call with out_f32 == NULL: clip n's values are read as
 * max(v, bf16((Lmax_n - 8 + 4) / 4)) (the log-mel floor of feature_extraction_whisper.py:157-158, applied here). */
int desta_im2col_k3(const void* x, int N, int T_in, int C, int stride, void* cols, const float* mel_clip_max, void* stream);

/* Q-Former layer mix (modeling_desta25.py:600-604): out[n,q,:] = sum_k softmax(layer_weights[q,:])[k] * x[k,n,q,:]
 * x fp32 [taps, N*Q, D] -> out fp32 [N*Q, D]. */
int desta_layer_mix(const float* x, const float* layer_weights, int taps, int N, int Q, int D, float* out, void* stream);

/* ---- (3b) connector backward (row a6: the autograd graph of modeling_desta25.py:587-606 over
 *      $TF/models/bert/modeling_bert.py:168-207,234-298,330-356; taps need no gradient, encoder frozen :1451-1453). ---- */
/* dW[Nout,Kin] += dY[Mred,Nout]^T X[Mred,Kin]  (bf16 operands read in place, fp32 accumulate; split-K with atomics). */
int desta_gemm_bf16_wgrad(const void* dY, int ldy, const void* X, int ldx, int Mred, int Nout, int Kin, float* dW, int ldw,
                          void* stream);
/* Attention backward (Tq <= 128): dq/dk/dv bf16 with the same row layout as q/k/v; lse from desta_attention. */
int desta_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, int ldo,
                        const void* dout, int lddo, const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv,
                        int lddv, int batch, int heads, int Tq, int Tk, float scale, const desta_dropout* drop, void* stream);
/* y = (keep ? x * scale : 0) (+ resid) with the mask of (seed, site) over the [M,N] grid, N % 8 == 0.
 * Forward: BertSelfOutput / BertOutput  LayerNorm(dropout(dense(x)) + residual)  (modeling_bert.py:294-298, 352-356) with
 * resid != NULL; backward of the dropped dense output with resid == NULL.  out_f32 and/or out_bf16 (out_f32 may alias x). */
int desta_dropout_add(const float* x, int ldx, const float* resid, int ld_resid, int M, int N, const desta_dropout* drop,
                      float* out_f32, int ld_f32, void* out_bf16, int ld_bf16, void* stream);
/* Test hook: materialise the keep mask (1 byte per element, rows of `cols` elements padded to groups of 8 internally)
 * of (seed, site) for a [rows, cols] grid. */
int desta_dropout_mask(const desta_dropout* drop, int rows, int cols, uint8_t* mask, void* stream);
/* LayerNorm backward from the saved fp32 input; dgamma/dbeta are accumulated (+=). */
int desta_layernorm_bwd(const float* dy, int lddy, const float* x, int ldx, const float* gamma, float eps, int M, int D,
                        float* dx_f32, int ld_f32, void* dx_bf16, int ld_bf16, float* dgamma, float* dbeta, void* stream);
/* erf-GELU on a saved bf16 pre-activation, and its backward dh = dy * gelu'(h) (all bf16, n elements). */
int desta_gelu_fwd(const void* h, void* y, size_t n, void* stream);
int desta_gelu_bwd(const void* dy, const void* h, void* dh, size_t n, void* stream);
/* out[j] += sum_m x[m,j]  (bias gradients; dtype 0 = fp32, 1 = bf16). */
int desta_colsum(const void* x, int dtype, int ld, int M, int N, float* out, void* stream);
/* out[h*64 + c] += sum_{b,t} w[(b*heads + h)*Tq + t] * x[(b*Tq + t)*ld + h*64 + c]   (x bf16; w == NULL: weights 1).
 * Value-bias gradient of an attention whose rows of P sum to w (1 without dropout: O = P (X Wv^T + 1 bv^T) carries bv as is),
 * computed from dO [batch*Tq, heads*64] - replaces the column sum over dV [batch*Tk, heads*64]. */
int desta_colsum_heads(const void* x, int ld, const float* w, int batch, int heads, int Tq, float* out, void* stream);
/* Backward of desta_layer_mix: dx [taps,N*Q,D] (written), dlw [Q,taps] (+=). */
int desta_layer_mix_bwd(const float* dm, const float* x, const float* layer_weights, int taps, int N, int Q, int D, float* dx,
                        float* dlw, void* stream);
int desta_cast_f32_bf16(const float* x, void* y, size_t n, void* stream);
/* 64-bit content hash of each row (row_bytes % 4 == 0, rows 4-byte aligned): out[r] = sum over 32-bit words w_i of
 * mix64(w_i, i) mod 2^64 - position-keyed, order-independent sum, hence deterministic.  Scope row n2: the ASR pass of
 * generate() (modeling_desta25.py:1582-1601) encodes a SUBSET of the clips the perception call encodes again; the host
 * keys the encoder state per clip with this hash (and confirms candidates bit-for-bit) so no clip is encoded twice. */
int desta_row_hash(const void* rows, int64_t n_rows, int64_t row_bytes, int64_t row_stride_bytes, uint64_t* out, void* stream);
/* ---- (4) splice: replaces the gathers + per-audio scatter loop of
 *      DeSTA25AudioModel._prepare_inputs_for_llm (modeling_desta25.py:940-1050; gathers :975-982,1009; loop :1014-1045).
 *      out[b,s,:] = table[input_ids[b,s]] for text rows; rows [start, start+Q) of prompt b <- bf16_rne(audio_feats[i]);
 *      rows [start+Q, start+Q+T_i) <- table[trans_ids[trans_off[i] + j]].  Later clips win on overlap (loop order). ---- */
typedef struct desta_splice_desc {
  int32_t prompt;      /* text_batch_idx */
  int32_t start;       /* audio_start_position */
  int32_t n_trans;     /* T_i */
  int32_t trans_off;   /* offset into trans_ids */
} desta_splice_desc;
/* feat_dtype: 0 = fp32, 1 = bf16.  Returns DESTA_ERR_SHAPE if a span leaves [0,S) (the reference raises at :1041); descs are
 * validated on the host copy `descs_host`.  err_word: optional int64[2] the kernel can write (device memory, or pinned host
 * memory under UVA): token ids outside [0, vocab) make the reference's embedding lookup raise (:981, :1009); the kernel
 * reads a clamped row instead (memory safety) and reports it as err_word[0] += offending rows, err_word[1] = one such id. */
int desta_splice(const int64_t* input_ids, const void* table, int64_t vocab, const desta_splice_desc* descs_dev,
                 const desta_splice_desc* descs_host, int n_desc, const int64_t* trans_ids, const void* audio_feats,
                 int feat_dtype, int Q, void* out, int B, int S, int d, int64_t* err_word, void* stream);
/* Gradient of the splice w.r.t. the audio features: dfeats[i] = grad_out[prompt_i, start_i : start_i + Q] (bf16 -> fp32). */
int desta_splice_bwd(const void* grad_out, const desta_splice_desc* descs_dev, int n_desc, int Q, int S, int d, float* dfeats,
                     void* stream);

/* ---- (0) PCM decode + resample to 16 kHz (scope row n3): replaces the sample conversion, channel average and
 *      librosa.core.resample call of AudioSegment.__init__ / .from_file (desta/utils/audio.py:147-166, 221-243, 305-315;
 *      called with target_sr=16000, channel_selector="average" at modeling_desta25.py:1553-1557 and
 *      trainer/data/simple_dataset.py:145-149), fused with the right-pad / truncate to 30 s of the feature extractor.
 *      The host only parses the container header and hands over the raw interleaved PCM bytes.
 *      out[c, n] = sum_k x_c[k] g(n M/L - k) for n < n_out_c, 0 for n_out_c <= n < out_len, where x_c = channel mean of the
 *      float32-converted frames (soundfile scaling: s16 / 2^15, s24 / 2^23, s32 / 2^31, u8 (v-128)/2^7, floats as they are)
 *      and g is a Kaiser-windowed sinc with the band edges of the resampler the reference uses (librosa's default
 *      res_type "soxr_hq": pass band to 0.913 of the lower Nyquist, stop band from 1.0, ~120 dB).  sr_in == sr_out gives
 *      L = M = P = 1 and the call reduces to conversion + down-mix + padding, bit-exact with the reference. ---- */
#define DESTA_PCM_U8 0
#define DESTA_PCM_S16 1
#define DESTA_PCM_S24 2
#define DESTA_PCM_S32 3
#define DESTA_PCM_F32 4
#define DESTA_PCM_F64 5
typedef struct desta_resample_plan {
  int32_t L, M;          /* sr_out / sr_in reduced: output n sits at input time n M / L */
  int32_t P;             /* taps per output sample (even, or 1 for the identity plan) */
  int32_t reserved;      /* flags; bit 0: force the general (any-ratio) kernel - the tests use it to cover both */
} desta_resample_plan;
typedef struct desta_pcm_clip {
  const void* pcm;       /* device pointer to interleaved frames, little-endian, `channels` samples per frame */
  int64_t n_frames;
  int64_t n_out;         /* ceil(n_frames L / M) as librosa fixes the length (audio.py:164); the kernel clamps it to out_len */
  int64_t out_row;       /* row of `out` this clip is written to (clips of one batch may be grouped by format) */
} desta_pcm_clip;
/* Fills *plan; returns the byte size of the coefficient table (L rows of P floats, row r = output residue n mod L). */
size_t desta_resample_plan_bytes(int sr_in, int sr_out, desta_resample_plan* plan);
int desta_resample_plan_build(int sr_in, int sr_out, float* host_taps, size_t bytes);
/* clips_dev[n_clips] (device array) share one format / channel count / rate; out [rows, out_stride] f32, row = out_row. */
int desta_decode_resample(const desta_pcm_clip* clips_dev, int n_clips, int fmt, int channels, const desta_resample_plan* plan,
                          const float* taps_dev, float* out, int64_t out_stride, int64_t out_len, void* stream);

/* ---- (2)/(3) stage-level entry points: the whole encoder / connector forward as ONE call (native schedule over the
 *      building blocks above; nothing is allocated - the caller passes a workspace of desta_*_workspace_bytes()).
 *      desta_encoder_fwd replaces WhisperPerception.forward_whisper's conv stem + layer loop with taps
 *      (modeling_desta25.py:563-598); desta_connector_fwd replaces the Q-Former use at :587-606.  The connector
 *      backward (training) is scheduled by the host over the (3b) entry points. ---- */
#define DESTA_MAX_TAPS 8
typedef struct desta_dims {
  int n_mels, d_model, n_heads, ffn, n_layers;
  int n_taps, tap_layers[DESTA_MAX_TAPS];     /* 0-based encoder layer ids, ascending (modeling_desta25.py:132-146) */
  int qf_layers, qf_ffn, prompt_size, d_llm, max_pos;
} desta_dims;
typedef struct desta_enc_layer {              /* bf16 GEMM weights [out,in] row-major, fp32 bias / LayerNorm */
  const float *ln1_w, *ln1_b;
  const void* qkv_w; const float* qkv_b;      /* [3d,d] = cat(q,k,v); bias with zeros for k (k_proj has none) */
  const void* o_w; const float* o_b;
  const float *ln2_w, *ln2_b;
  const void* fc1_w; const float* fc1_b;
  const void* fc2_w; const float* fc2_b;
  /* optional LayerNorm-folded operands (all six set -> desta_encoder_fwd runs without LayerNorm launches):
   * *_w_ln = bf16(gain[k] * W[n,k]), *_s[n] = sum_k of that bf16 row (fp32), *_c[n] = sum_k beta[k] W[n,k] + bias[n] */
  const void* qkv_w_ln; const float *qkv_s, *qkv_c;
  const void* fc1_w_ln; const float *fc1_s, *fc1_c;
} desta_enc_layer;
typedef struct desta_encoder_weights {
  const void* conv1_w; const float* conv1_b;  /* [d, 3*n_mels] (tap-major im2col order) */
  const void* conv2_w; const float* conv2_b;  /* [d, 3*d] */
  const float* pos;                           /* [max_pos, d] */
  const desta_enc_layer* layers;              /* [n_layers] (host array) */
} desta_encoder_weights;
size_t desta_encoder_workspace_bytes(const desta_dims* dims, int N);
/* mel_tm: bf16 time-major log-mel [N, 2*max_pos, n_mels]; mel_clip_max: NULL if mel_tm is final, else the clip_max words of
 * the desta_logmel call that produced it without the floor (see desta_logmel / desta_im2col_k3).
 * taps: bf16 [n_taps, N*max_pos, d].  h: fp32 [N*max_pos, d] residual stream after the last layer (input of the final
 * LayerNorm for the ASR pass). */
int desta_encoder_fwd(const desta_dims* dims, const desta_encoder_weights* w, const void* mel_tm, const float* mel_clip_max, int N,
                      void* workspace, size_t workspace_bytes, void* taps, float* h, void* stream);

typedef struct desta_qf_layer {
  const void* s_qkv_w; const float* s_qkv_b; const void* s_o_w; const float* s_o_b; const float *s_ln_w, *s_ln_b;
  const void* c_q_w; const float* c_q_b; const void* c_o_w; const float* c_o_b; const float *c_ln_w, *c_ln_b;
  const void* f1_w; const float* f1_b; const void* f2_w; const float* f2_b; const float *f_ln_w, *f_ln_b;
} desta_qf_layer;
typedef struct desta_connector_weights {
  const float* layer_weights;                 /* [prompt_size, n_taps] */
  const float* prompts;                       /* [n_taps * prompt_size, d] */
  const void* kv_w; const float* kv_b;        /* cross-attention K/V of all layers: [qf_layers*2*d, d], [qf_layers*2*d] */
  const float *proj_ln_w, *proj_ln_b;
  const void* proj_w; const float* proj_b;    /* [d_llm, d] */
  const desta_qf_layer* layers;               /* [qf_layers] (host array) */
} desta_connector_weights;
size_t desta_connector_workspace_bytes(const desta_dims* dims, int N);
/* taps as written by desta_encoder_fwd -> out fp32 [N*prompt_size, d_llm] (eval mode: no dropout). */
int desta_connector_fwd(const desta_dims* dims, const desta_connector_weights* w, const void* taps, int N, void* workspace,
                        size_t workspace_bytes, float* out, void* stream);

/* ---- (3b) connector training as two stage-level calls: the forward that keeps the activations the backward needs,
 *      and the backward that fills ONE fp32 gradient arena (row a6: autograd of modeling_desta25.py:587-606 over
 *      $TF/models/bert/modeling_bert.py; only the connector is trainable, :1456-1463, so taps get no gradient).
 *      Arena layout (every tensor padded to a multiple of 4 floats): layer_weights [Q,taps], prompts [taps*Q,d],
 *      kv_w [L*2d,d], kv_b [L*2d], proj_ln_w [d], proj_ln_b [d], proj_w [d_llm,d], proj_b [d_llm]; then per Q-Former
 *      layer: s_qkv_w [3d,d], s_qkv_b, s_o_w [d,d], s_o_b, s_ln_w, s_ln_b, c_q_w [d,d], c_q_b, c_o_w [d,d], c_o_b, c_ln_w,
 *      c_ln_b, f1_w [ffn,d], f1_b, f2_w [d,ffn], f2_b, f_ln_w, f_ln_b  (= the fused layouts of desta_connector_weights).
 *      Dropout: p_hidden / p_attn = 0 in eval mode; site = layer * 8 + {0 self probs, 1 self out, 2 cross probs,
 *      3 cross out, 4 FFN out}; the backward regenerates the masks from the same seed. ---- */
typedef struct desta_dropout_plan { float p_hidden, p_attn; uint64_t seed; } desta_dropout_plan;
typedef struct desta_qf_layer_t {            /* transposed bf16 weights [in,out] for the data-gradient GEMMs */
  const void *s_qkv, *s_o, *c_q, *c_o, *f1, *f2;
} desta_qf_layer_t;
typedef struct desta_connector_weights_t { const void* proj; const desta_qf_layer_t* layers; } desta_connector_weights_t;
size_t desta_connector_train_workspace_bytes(const desta_dims* dims, int N);
size_t desta_connector_grad_floats(const desta_dims* dims);
int desta_connector_train_fwd(const desta_dims* dims, const desta_connector_weights* w, const void* taps, int N, void* workspace,
                              size_t workspace_bytes, const desta_dropout_plan* plan, float* out, void* stream);
/* grad_arena must be zeroed by the caller (gradients are accumulated).  on_ready(user, first_float, n_floats) - optional -
 * is called on the host right after the launches that complete a contiguous arena range (one Q-Former layer at a time,
 * last layer first; the leading non-layer range at the end), so that a data-parallel all-reduce can be enqueued behind
 * them while the rest of the backward is still being issued. */
typedef void (*desta_grad_ready_fn)(void* user, size_t first_float, size_t n_floats);
int desta_connector_bwd(const desta_dims* dims, const desta_connector_weights* w, const desta_connector_weights_t* wt,
                        const void* taps, int N, void* workspace, size_t workspace_bytes, const desta_dropout_plan* plan,
                        const float* grad_out, float* grad_arena, size_t arena_floats, desta_grad_ready_fn on_ready,
                        void* user, void* stream);

#ifdef __cplusplus
}
#endif
#endif

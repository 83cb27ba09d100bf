// Stage-level entry points: the encoder forward with taps and the Q-Former connector forward as one C call each.
// Host-side C++ that only sequences the kernels of this library on the caller's stream over a caller-owned workspace
// (ref: desta/models/modeling_desta25.py:563-598 conv stem + layer loop with taps, :587-606 Q-Former + mix + proj;
// $TF/models/whisper/modeling_whisper.py:380-414; $TF/models/bert/modeling_bert.py:379-421).
// Arithmetic contract: bf16 GEMM operands, fp32 accumulation, fp32 residual stream, fp32
// LayerNorm / softmax statistics, taps handed to the connector in bf16.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "desta_b200.h"

namespace desta {

struct Carver {
  uint8_t* base; size_t off;
  void* take(size_t bytes) { void* p = base ? base + off : nullptr; off += (bytes + 1023) & ~(size_t)1023; return p; }
};

struct EncWs { __nv_bfloat16 *cols1, *c1, *cols2, *y, *qkv, *ao, *ff; float* stats; size_t bytes; };
static EncWs carve_encoder(const desta_dims* d, int N, void* ws) {
  Carver c{(uint8_t*)ws, 0};
  const size_t M = (size_t)N * d->max_pos, D = d->d_model;
  EncWs w;
  w.cols1 = (__nv_bfloat16*)c.take(2 * M * 3 * d->n_mels * 2);
  w.c1 = (__nv_bfloat16*)c.take(2 * M * D * 2);
  w.cols2 = (__nv_bfloat16*)c.take(M * 3 * D * 2);
  w.y = (__nv_bfloat16*)c.take(M * D * 2);
  w.qkv = (__nv_bfloat16*)c.take(M * 3 * D * 2);
  w.ao = (__nv_bfloat16*)c.take(M * D * 2);
  w.ff = (__nv_bfloat16*)c.take(M * (size_t)d->ffn * 2);
  w.stats = (float*)c.take(M * (size_t)((D + 127) / 128) * 2 * 4);     // LayerNorm hand-over: per row and 128-column group (sum, sumsq)
  w.bytes = c.off;
  return w;
}

struct ConWs {
  __nv_bfloat16 *kv, *x_bf, *q_qkv, *q_ao, *a_bf, *q_q, *c_bf, *q_ff, *mix_bf;
  float *x_f32, *q_tmp, *a_f32, *c_f32, *mix;
  size_t bytes;
};
static ConWs carve_connector(const desta_dims* d, int N, void* ws) {
  Carver c{(uint8_t*)ws, 0};
  const size_t D = d->d_model, K = d->n_taps, R = K * N * d->prompt_size, NQ = (size_t)N * d->prompt_size;
  ConWs w;
  w.kv = (__nv_bfloat16*)c.take(K * N * d->max_pos * (size_t)d->qf_layers * 2 * D * 2);
  w.x_f32 = (float*)c.take(R * D * 4);   w.x_bf = (__nv_bfloat16*)c.take(R * D * 2);
  w.q_qkv = (__nv_bfloat16*)c.take(R * 3 * D * 2);
  w.q_ao = (__nv_bfloat16*)c.take(R * D * 2);
  w.q_tmp = (float*)c.take(R * D * 4);
  w.a_f32 = (float*)c.take(R * D * 4);   w.a_bf = (__nv_bfloat16*)c.take(R * D * 2);
  w.q_q = (__nv_bfloat16*)c.take(R * D * 2);
  w.c_f32 = (float*)c.take(R * D * 4);   w.c_bf = (__nv_bfloat16*)c.take(R * D * 2);
  w.q_ff = (__nv_bfloat16*)c.take(R * (size_t)d->qf_ffn * 2);
  w.mix = (float*)c.take(NQ * D * 4);    w.mix_bf = (__nv_bfloat16*)c.take(NQ * D * 2);
  w.bytes = c.off;
  return w;
}

static int check_dims(const desta_dims* d, const char* who) {
  if (!d || d->n_heads <= 0 || d->d_model != d->n_heads * 64 || d->n_taps < 1 || d->n_taps > DESTA_MAX_TAPS || d->n_layers < 1 ||
      d->max_pos < 1 || d->prompt_size < 1 || d->prompt_size > 128) {
    set_error("%s: unsupported dims (head_dim must be 64, 1..%d taps, prompt_size <= 128)", who, DESTA_MAX_TAPS);
    return DESTA_ERR_ARG;
  }
  return DESTA_OK;
}

// C = epi(A W^T): thin wrapper filling the epilogue struct
// LayerNorm hand-over of a GEMM (desta_gemm_epilogue): producer side = stats_out, consumer side = the other four
struct LnLink { float* stats_out; const float* stats_in; const float* colsum; int cols; float eps; };
static int gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out_bf16, int ld_bf16,
                float* out_f32, int ld_f32, const float* resid, int ld_resid, const float* pos, int pos_rows, float scale,
                int scale_cols, int gelu, void* stream, const LnLink* ln = nullptr) {
  desta_gemm_epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32; ep.out_bf16 = out_bf16; ep.ld_bf16 = ld_bf16; ep.bias = bias;
  ep.resid = resid; ep.ld_resid = ld_resid; ep.pos = pos; ep.pos_rows = pos_rows; ep.scale = scale; ep.scale_cols = scale_cols;
  ep.gelu = gelu;
  if (ln) {
    ep.ln_stats_out = ln->stats_out;
    ep.ln_stats_in = ln->stats_in; ep.ln_colsum = ln->colsum; ep.ln_cols = ln->cols; ep.ln_groups = (ln->cols + 127) / 128; ep.ln_eps = ln->eps;
  }
  return desta_gemm_bf16(A, lda, W, ldw, M, N, K, &ep, stream);
}

__global__ void __launch_bounds__(256)
broadcast_prompts_kernel(const float* __restrict__ prompts, int taps, int N, int QD, float* __restrict__ x_f32,
                         __nv_bfloat16* __restrict__ x_bf) {
  // x[k, n, :, :] = layer_prompts[k] for every clip n (modeling_desta25.py:590 `.expand(batch, -1, -1)`)
  const size_t total = (size_t)taps * N * (QD / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % (QD / 4), kn = i / (QD / 4), k = kn / N;
    const float4 p = __ldg(reinterpret_cast<const float4*>(prompts + k * QD) + v);
    reinterpret_cast<float4*>(x_f32 + kn * QD)[v] = p;
    reinterpret_cast<uint2*>(x_bf + kn * QD)[v] = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
  }
}


}  // namespace desta

using namespace desta;

#define RUN(call) do { int _rc = (call); if (_rc != DESTA_OK) return _rc; } while (0)

extern "C" size_t desta_encoder_workspace_bytes(const desta_dims* dims, int N) {
  if (check_dims(dims, "desta_encoder_workspace_bytes") || N <= 0) return 0;
  return carve_encoder(dims, N, nullptr).bytes;
}

extern "C" int desta_encoder_fwd(const desta_dims* dims, const desta_encoder_weights* w, const void* mel_tm, const float* mel_clip_max,
                                 int N, void* workspace, size_t workspace_bytes, void* taps, float* h, void* stream) {
  RUN(check_dims(dims, "desta_encoder_fwd"));
  if (!w || !w->layers || !mel_tm || !workspace || !taps || !h) { set_error("desta_encoder_fwd: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  const EncWs ws = carve_encoder(dims, N, workspace);
  if (workspace_bytes < ws.bytes) { set_error("desta_encoder_fwd: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes); return DESTA_ERR_ARG; }
  const int d = dims->d_model, T = dims->max_pos, M = N * T, H = dims->n_heads, F = dims->ffn;
  // q = (y Wq^T + bq) * head_dim^-1/2 (modeling_whisper.py:310).  The softmax works with base-2 exponentials, so log2(e)
  // is folded into the same epilogue scale and the attention kernel is called with scale = ln 2: its scores are then
  // already log2-domain exponents (UNIT mode of attn_fwd_kernel: the max shift is applied by the tensor cores).
  const float q_scale = 1.4426950408889634f / sqrtf((float)(d / H));
  const float attn_scale = 0.6931471805599453f;
  // conv stem as im2col + GEMM: gelu(conv1), gelu(conv2 stride 2) + positions -> fp32 residual stream (:563-570)
  RUN(desta_im2col_k3(mel_tm, N, 2 * T, dims->n_mels, 1, ws.cols1, mel_clip_max, stream));
  RUN(gemm(ws.cols1, 3 * dims->n_mels, w->conv1_w, 3 * dims->n_mels, 2 * M, d, 3 * dims->n_mels, w->conv1_b, ws.c1, d, nullptr, 0,
           nullptr, 0, nullptr, 0, 1.f, 0, 1, stream));
  RUN(desta_im2col_k3(ws.c1, N, 2 * T, d, 2, ws.cols2, nullptr, stream));
  // LayerNorm folded into the GEMMs around it when the pack carries the gain-folded weights: every residual GEMM
  // (conv2, out_proj, fc2) also writes the bf16 copy of the stream (`hb`; on tap layers that copy IS the tap) and the
  // per-row partial sums, the QKV / fc1 GEMMs read `hb` and normalise in their epilogue.  64 launches and 64 fp32 reads
  // of the [M, d] stream fewer per forward.
  bool fold = M > 128;
  for (int i = 0; i < dims->n_layers; ++i) {
    const desta_enc_layer& L = w->layers[i];
    fold = fold && L.qkv_w_ln && L.qkv_s && L.qkv_c && L.fc1_w_ln && L.fc1_s && L.fc1_c;
  }
  if (fold) {
    const LnLink produce = {ws.stats, nullptr, nullptr, 0, 0.f};
    const __nv_bfloat16* hb = ws.y;
    RUN(gemm(ws.cols2, 3 * d, w->conv2_w, 3 * d, M, d, 3 * d, w->conv2_b, ws.y, d, h, d, nullptr, 0, w->pos, T, 1.f, 0, 1, stream, &produce));
    int tap = 0;
    for (int i = 0; i < dims->n_layers; ++i) {
      const desta_enc_layer& L = w->layers[i];
      const LnLink use_qkv = {nullptr, ws.stats, L.qkv_s, d, 1e-5f}, use_fc1 = {nullptr, ws.stats, L.fc1_s, d, 1e-5f};
      RUN(gemm(hb, d, L.qkv_w_ln, d, M, 3 * d, d, L.qkv_c, ws.qkv, 3 * d, nullptr, 0, nullptr, 0, nullptr, 0, q_scale, d, 0, stream, &use_qkv));
      RUN(desta_attention(ws.qkv, 3 * d, ws.qkv + d, 3 * d, ws.qkv + 2 * d, 3 * d, ws.ao, d, N, H, T, T, attn_scale, nullptr, nullptr, stream));
      RUN(gemm(ws.ao, d, L.o_w, d, M, d, d, L.o_b, ws.y, d, h, d, h, d, nullptr, 0, 1.f, 0, 0, stream, &produce));
      RUN(gemm(ws.y, d, L.fc1_w_ln, d, M, F, d, L.fc1_c, ws.ff, F, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 1, stream, &use_fc1));
      __nv_bfloat16* copy = ws.y;
      if (tap < dims->n_taps && dims->tap_layers[tap] == i) { copy = (__nv_bfloat16*)taps + (size_t)tap * M * d; ++tap; }
      RUN(gemm(ws.ff, F, L.fc2_w, F, M, d, F, L.fc2_b, copy, d, h, d, h, d, nullptr, 0, 1.f, 0, 0, stream, &produce));
      hb = copy;
    }
    if (tap != dims->n_taps) { set_error("desta_encoder_fwd: tap_layers must be ascending layer ids < n_layers"); return DESTA_ERR_ARG; }
    return DESTA_OK;
  }
  RUN(gemm(ws.cols2, 3 * d, w->conv2_w, 3 * d, M, d, 3 * d, w->conv2_b, nullptr, 0, h, d, nullptr, 0, w->pos, T, 1.f, 0, 1, stream));
  int tap = 0;
  for (int i = 0; i < dims->n_layers; ++i) {
    const desta_enc_layer& L = w->layers[i];
    RUN(desta_layernorm(h, d, L.ln1_w, L.ln1_b, 1e-5f, M, d, ws.y, d, nullptr, 0, stream));
    RUN(gemm(ws.y, d, L.qkv_w, d, M, 3 * d, d, L.qkv_b, ws.qkv, 3 * d, nullptr, 0, nullptr, 0, nullptr, 0, q_scale, d, 0, stream));
    RUN(desta_attention(ws.qkv, 3 * d, ws.qkv + d, 3 * d, ws.qkv + 2 * d, 3 * d, ws.ao, d, N, H, T, T, attn_scale, nullptr, nullptr, stream));
    RUN(gemm(ws.ao, d, L.o_w, d, M, d, d, L.o_b, nullptr, 0, h, d, h, d, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_layernorm(h, d, L.ln2_w, L.ln2_b, 1e-5f, M, d, ws.y, d, nullptr, 0, stream));
    RUN(gemm(ws.y, d, L.fc1_w, d, M, F, d, L.fc1_b, ws.ff, F, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 1, stream));
    void* tap_out = nullptr;
    if (tap < dims->n_taps && dims->tap_layers[tap] == i) { tap_out = (__nv_bfloat16*)taps + (size_t)tap * M * d; ++tap; }
    RUN(gemm(ws.ff, F, L.fc2_w, F, M, d, F, L.fc2_b, tap_out, d, h, d, h, d, nullptr, 0, 1.f, 0, 0, stream));
  }
  if (tap != dims->n_taps) { set_error("desta_encoder_fwd: tap_layers must be ascending layer ids < n_layers"); return DESTA_ERR_ARG; }
  return DESTA_OK;
}

extern "C" size_t desta_connector_workspace_bytes(const desta_dims* dims, int N) {
  if (check_dims(dims, "desta_connector_workspace_bytes") || N <= 0) return 0;
  return carve_connector(dims, N, nullptr).bytes;
}

extern "C" int desta_connector_fwd(const desta_dims* dims, const desta_connector_weights* w, const void* taps, int N, void* workspace,
                                   size_t workspace_bytes, float* out, void* stream) {
  RUN(check_dims(dims, "desta_connector_fwd"));
  if (!w || !w->layers || !taps || !workspace || !out) { set_error("desta_connector_fwd: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  const ConWs ws = carve_connector(dims, N, workspace);
  if (workspace_bytes < ws.bytes) { set_error("desta_connector_fwd: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes); return DESTA_ERR_ARG; }
  const int d = dims->d_model, T = dims->max_pos, Q = dims->prompt_size, K = dims->n_taps, H = dims->n_heads, F = dims->qf_ffn;
  const int B = K * N, R = B * Q, ldkv = dims->qf_layers * 2 * d;
  const float scale = 1.0f / sqrtf((float)(d / H));
  if ((size_t)K * N * T > 0x7fffffffull) { set_error("desta_connector_fwd: too many tap rows for one GEMM"); return DESTA_ERR_ARG; }
  // cross-attention K/V of every Q-Former layer for every tap: one GEMM over [K*N*1500, d]
  RUN(gemm(taps, d, w->kv_w, d, K * N * T, ldkv, d, w->kv_b, ws.kv, ldkv, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
  {
    const size_t total = (size_t)B * (Q * d / 4);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    broadcast_prompts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w->prompts, K, N, Q * d, ws.x_f32, ws.x_bf);
    RUN(check_launch("broadcast_prompts_kernel"));
  }
  for (int l = 0; l < dims->qf_layers; ++l) {
    const desta_qf_layer& L = w->layers[l];
    RUN(gemm(ws.x_bf, d, L.s_qkv_w, d, R, 3 * d, d, L.s_qkv_b, ws.q_qkv, 3 * d, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_attention(ws.q_qkv, 3 * d, ws.q_qkv + d, 3 * d, ws.q_qkv + 2 * d, 3 * d, ws.q_ao, d, B, H, Q, Q, scale, nullptr, nullptr, stream));
    RUN(gemm(ws.q_ao, d, L.s_o_w, d, R, d, d, L.s_o_b, nullptr, 0, ws.q_tmp, d, ws.x_f32, d, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_layernorm(ws.q_tmp, d, L.s_ln_w, L.s_ln_b, 1e-12f, R, d, ws.a_bf, d, ws.a_f32, d, stream));
    RUN(gemm(ws.a_bf, d, L.c_q_w, d, R, d, d, L.c_q_b, ws.q_q, d, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
    const __nv_bfloat16* kl = ws.kv + (size_t)l * 2 * d;
    RUN(desta_attention(ws.q_q, d, kl, ldkv, kl + d, ldkv, ws.q_ao, d, B, H, Q, T, scale, nullptr, nullptr, stream));
    RUN(gemm(ws.q_ao, d, L.c_o_w, d, R, d, d, L.c_o_b, nullptr, 0, ws.q_tmp, d, ws.a_f32, d, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_layernorm(ws.q_tmp, d, L.c_ln_w, L.c_ln_b, 1e-12f, R, d, ws.c_bf, d, ws.c_f32, d, stream));
    RUN(gemm(ws.c_bf, d, L.f1_w, d, R, F, d, L.f1_b, ws.q_ff, F, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 1, stream));
    RUN(gemm(ws.q_ff, F, L.f2_w, F, R, d, F, L.f2_b, nullptr, 0, ws.q_tmp, d, ws.c_f32, d, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_layernorm(ws.q_tmp, d, L.f_ln_w, L.f_ln_b, 1e-12f, R, d, ws.x_bf, d, ws.x_f32, d, stream));
  }
  RUN(desta_layer_mix(ws.x_f32, w->layer_weights, K, N, Q, d, ws.mix, stream));
  RUN(desta_layernorm(ws.mix, d, w->proj_ln_w, w->proj_ln_b, 1e-5f, N * Q, d, ws.mix_bf, d, nullptr, 0, stream));
  RUN(gemm(ws.mix_bf, d, w->proj_w, d, N * Q, dims->d_llm, d, w->proj_b, nullptr, 0, out, dims->d_llm, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
  return DESTA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Connector training: forward with saved activations + hand-scheduled backward (row a6).  Same launches, same order as
// the host-side per-op schedule in desta_b200/autograd.py (kept as the reference for this schedule; tested bit-identical).
// ---------------------------------------------------------------------------------------------------------------------
namespace desta {

struct TrainLayer {
  __nv_bfloat16 *x_bf, *qkv, *ao_s, *a_bf, *q_c, *ao_c, *c_bf, *h1, *f;
  float *x_f32, *lse_s, *t1, *a_f32, *lse_c, *rr_c, *t2, *c_f32, *t3;
};
struct TrainWs {
  __nv_bfloat16 *kv, *x_out_bf, *mix_bf, *dkv, *g_bf_a, *g_bf_b, *g_qkv, *g_ff, *g_h1, *g_out_bf;
  float *x_out, *mix, *g_f32_a, *g_f32_b, *g_f32_c, *g_mix, *g_mixn, *g_x;
  TrainLayer layer[32];
  size_t bytes;
};
static TrainWs carve_train(const desta_dims* d, int N, void* ws) {
  Carver c{(uint8_t*)ws, 0};
  const size_t D = d->d_model, K = d->n_taps, Q = d->prompt_size, T = d->max_pos, H = d->n_heads, F = d->qf_ffn, L = d->qf_layers;
  const size_t R = K * N * Q, B = K * N, NQ = (size_t)N * Q;
  TrainWs w;
  w.kv = (__nv_bfloat16*)c.take(K * N * T * L * 2 * D * 2);
  for (size_t l = 0; l < L; ++l) {
    TrainLayer& S = w.layer[l];
    S.x_bf = (__nv_bfloat16*)c.take(R * D * 2);   S.x_f32 = (float*)c.take(R * D * 4);
    S.qkv = (__nv_bfloat16*)c.take(R * 3 * D * 2); S.lse_s = (float*)c.take(B * H * Q * 4);
    S.ao_s = (__nv_bfloat16*)c.take(R * D * 2);    S.t1 = (float*)c.take(R * D * 4);
    S.a_bf = (__nv_bfloat16*)c.take(R * D * 2);    S.a_f32 = (float*)c.take(R * D * 4);
    S.q_c = (__nv_bfloat16*)c.take(R * D * 2);     S.lse_c = (float*)c.take(B * H * Q * 4);
    S.rr_c = (float*)c.take(B * H * Q * 4);         // cross attention: dropped / plain row sums of P (value-bias gradient)
    S.ao_c = (__nv_bfloat16*)c.take(R * D * 2);    S.t2 = (float*)c.take(R * D * 4);
    S.c_bf = (__nv_bfloat16*)c.take(R * D * 2);    S.c_f32 = (float*)c.take(R * D * 4);
    S.h1 = (__nv_bfloat16*)c.take(R * F * 2);      S.f = (__nv_bfloat16*)c.take(R * F * 2);
    S.t3 = (float*)c.take(R * D * 4);
  }
  w.x_out = (float*)c.take(R * D * 4);    w.x_out_bf = (__nv_bfloat16*)c.take(R * D * 2);
  w.mix = (float*)c.take(NQ * D * 4);     w.mix_bf = (__nv_bfloat16*)c.take(NQ * D * 2);
  w.dkv = (__nv_bfloat16*)c.take(K * N * T * L * 2 * D * 2);
  w.g_f32_a = (float*)c.take(R * D * 4);  w.g_f32_b = (float*)c.take(R * D * 4);  w.g_f32_c = (float*)c.take(R * D * 4);
  w.g_bf_a = (__nv_bfloat16*)c.take(R * D * 2); w.g_bf_b = (__nv_bfloat16*)c.take(R * D * 2);
  w.g_qkv = (__nv_bfloat16*)c.take(R * 3 * D * 2);
  w.g_ff = (__nv_bfloat16*)c.take(R * F * 2);   w.g_h1 = (__nv_bfloat16*)c.take(R * F * 2);
  w.g_out_bf = (__nv_bfloat16*)c.take(NQ * (size_t)d->d_llm * 2);
  w.g_mix = (float*)c.take(NQ * D * 4);   w.g_mixn = (float*)c.take(NQ * D * 4);
  w.g_x = (float*)c.take(R * D * 4);
  w.bytes = c.off;
  return w;
}

// gradient arena (floats), same order / padding as desta_b200/autograd.py TrainState
struct GradLayer { float *s_qkv_w, *s_qkv_b, *s_o_w, *s_o_b, *s_ln_w, *s_ln_b, *c_q_w, *c_q_b, *c_o_w, *c_o_b, *c_ln_w, *c_ln_b,
                         *f1_w, *f1_b, *f2_w, *f2_b, *f_ln_w, *f_ln_b; size_t first, count; };
struct GradArena { float *layer_weights, *prompts, *kv_w, *kv_b, *proj_ln_w, *proj_ln_b, *proj_w, *proj_b; size_t global_count;
                   GradLayer layer[32]; size_t floats; };
static GradArena carve_grads(const desta_dims* d, float* base) {
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) / 4 * 4; return p; };
  const size_t D = d->d_model, K = d->n_taps, Q = d->prompt_size, F = d->qf_ffn, L = d->qf_layers, DL = d->d_llm;
  GradArena g;
  g.layer_weights = take(Q * K); g.prompts = take(K * Q * D); g.kv_w = take(L * 2 * D * D); g.kv_b = take(L * 2 * D);
  g.proj_ln_w = take(D); g.proj_ln_b = take(D); g.proj_w = take(DL * D); g.proj_b = take(DL);
  g.global_count = off;
  for (size_t l = 0; l < L; ++l) {
    GradLayer& x = g.layer[l];
    x.first = off;
    x.s_qkv_w = take(3 * D * D); x.s_qkv_b = take(3 * D); x.s_o_w = take(D * D); x.s_o_b = take(D); x.s_ln_w = take(D); x.s_ln_b = take(D);
    x.c_q_w = take(D * D); x.c_q_b = take(D); x.c_o_w = take(D * D); x.c_o_b = take(D); x.c_ln_w = take(D); x.c_ln_b = take(D);
    x.f1_w = take(F * D); x.f1_b = take(F); x.f2_w = take(D * F); x.f2_b = take(D); x.f_ln_w = take(D); x.f_ln_b = take(D);
    x.count = off - x.first;
  }
  g.floats = off;
  return g;
}

static bool site(const desta_dropout_plan* plan, int layer, int which, desta_dropout* out) {
  if (!plan) return false;
  const float p = (which == 0 || which == 2) ? plan->p_attn : plan->p_hidden;
  long thr = lroundf(p * 65536.0f);
  if (thr <= 0) return false;
  if (thr > 65535) thr = 65535;
  out->thr16 = (uint32_t)thr;
  out->scale = (float)(1.0 / (1.0 - (double)thr / 65536.0));
  out->seed = plan->seed;
  out->site = (uint32_t)(layer * 8 + which);
  out->row_keep_ratio = nullptr;
  return true;
}

// t = dropout(a W^T + b) + resid  (BertSelfOutput / BertOutput before their LayerNorm)
static int dense_res(const void* a, int lda, const void* W, int M, int N, int K, const float* bias, float* t, const float* resid,
                     const desta_dropout_plan* plan, int layer, int which, void* stream) {
  desta_dropout dr;
  if (!site(plan, layer, which, &dr)) return gemm(a, lda, W, K, M, N, K, bias, nullptr, 0, t, N, resid, N, nullptr, 0, 1.f, 0, 0, stream);
  int rc = gemm(a, lda, W, K, M, N, K, bias, nullptr, 0, t, N, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream);
  if (rc) return rc;
  return desta_dropout_add(t, N, resid, N, M, N, &dr, t, N, nullptr, 0, stream);
}

}  // namespace desta

extern "C" size_t desta_connector_train_workspace_bytes(const desta_dims* dims, int N) {
  if (check_dims(dims, "desta_connector_train_workspace_bytes") || N <= 0 || dims->qf_layers > 32) return 0;
  return carve_train(dims, N, nullptr).bytes;
}

extern "C" size_t desta_connector_grad_floats(const desta_dims* dims) {
  if (check_dims(dims, "desta_connector_grad_floats") || dims->qf_layers > 32) return 0;
  return carve_grads(dims, nullptr).floats;
}

extern "C" int desta_connector_train_fwd(const desta_dims* dims, const desta_connector_weights* w, const void* taps, int N,
                                         void* workspace, size_t workspace_bytes, const desta_dropout_plan* plan, float* out,
                                         void* stream) {
  RUN(check_dims(dims, "desta_connector_train_fwd"));
  if (!w || !w->layers || !taps || !workspace || !out || dims->qf_layers > 32) { set_error("desta_connector_train_fwd: bad argument"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  TrainWs ws = carve_train(dims, N, workspace);
  if (workspace_bytes < ws.bytes) { set_error("desta_connector_train_fwd: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes); return DESTA_ERR_ARG; }
  const int d = dims->d_model, T = dims->max_pos, Q = dims->prompt_size, K = dims->n_taps, H = dims->n_heads, F = dims->qf_ffn;
  const int B = K * N, R = B * Q, ldkv = dims->qf_layers * 2 * d, L = dims->qf_layers;
  const float scale = 1.0f / sqrtf((float)(d / H));
  RUN(gemm(taps, d, w->kv_w, d, K * N * T, ldkv, d, w->kv_b, ws.kv, ldkv, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
  {
    const size_t total = (size_t)B * (Q * d / 4);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    broadcast_prompts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w->prompts, K, N, Q * d, ws.layer[0].x_f32, ws.layer[0].x_bf);
    RUN(check_launch("broadcast_prompts_kernel"));
  }
  for (int l = 0; l < L; ++l) {
    const desta_qf_layer& P = w->layers[l];
    TrainLayer& S = ws.layer[l];
    desta_dropout dr;
    RUN(gemm(S.x_bf, d, P.s_qkv_w, d, R, 3 * d, d, P.s_qkv_b, S.qkv, 3 * d, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
    RUN(desta_attention(S.qkv, 3 * d, S.qkv + d, 3 * d, S.qkv + 2 * d, 3 * d, S.ao_s, d, B, H, Q, Q, scale, S.lse_s,
                        site(plan, l, 0, &dr) ? &dr : nullptr, stream));
    RUN(dense_res(S.ao_s, d, P.s_o_w, R, d, d, P.s_o_b, S.t1, S.x_f32, plan, l, 1, stream));
    RUN(desta_layernorm(S.t1, d, P.s_ln_w, P.s_ln_b, 1e-12f, R, d, S.a_bf, d, S.a_f32, d, stream));
    RUN(gemm(S.a_bf, d, P.c_q_w, d, R, d, d, P.c_q_b, S.q_c, d, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
    const __nv_bfloat16* kl = ws.kv + (size_t)l * 2 * d;
    const bool cross_drop = site(plan, l, 2, &dr);
    if (cross_drop) dr.row_keep_ratio = S.rr_c;
    RUN(desta_attention(S.q_c, d, kl, ldkv, kl + d, ldkv, S.ao_c, d, B, H, Q, T, scale, S.lse_c, cross_drop ? &dr : nullptr, stream));
    RUN(dense_res(S.ao_c, d, P.c_o_w, R, d, d, P.c_o_b, S.t2, S.a_f32, plan, l, 3, stream));
    RUN(desta_layernorm(S.t2, d, P.c_ln_w, P.c_ln_b, 1e-12f, R, d, S.c_bf, d, S.c_f32, d, stream));
    RUN(gemm(S.c_bf, d, P.f1_w, d, R, F, d, P.f1_b, S.h1, F, nullptr, 0, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));   // pre-activation kept
    RUN(desta_gelu_fwd(S.h1, S.f, (size_t)R * F, stream));
    RUN(dense_res(S.f, F, P.f2_w, R, d, F, P.f2_b, S.t3, S.c_f32, plan, l, 4, stream));
    __nv_bfloat16* nx_bf = l + 1 < L ? ws.layer[l + 1].x_bf : ws.x_out_bf;
    float* nx_f32 = l + 1 < L ? ws.layer[l + 1].x_f32 : ws.x_out;
    RUN(desta_layernorm(S.t3, d, P.f_ln_w, P.f_ln_b, 1e-12f, R, d, nx_bf, d, nx_f32, d, stream));
  }
  RUN(desta_layer_mix(ws.x_out, w->layer_weights, K, N, Q, d, ws.mix, stream));
  RUN(desta_layernorm(ws.mix, d, w->proj_ln_w, w->proj_ln_b, 1e-5f, N * Q, d, ws.mix_bf, d, nullptr, 0, stream));
  RUN(gemm(ws.mix_bf, d, w->proj_w, d, N * Q, dims->d_llm, d, w->proj_b, nullptr, 0, out, dims->d_llm, nullptr, 0, nullptr, 0, 1.f, 0, 0, stream));
  return DESTA_OK;
}

extern "C" int desta_connector_bwd(const desta_dims* dims, const desta_connector_weights* w, const desta_connector_weights_t* wt,
                                   const void* taps, int N, void* workspace, size_t workspace_bytes, const desta_dropout_plan* plan,
                                   const float* grad_out, float* grad_arena, size_t arena_floats, desta_grad_ready_fn on_ready,
                                   void* user, void* stream) {
  RUN(check_dims(dims, "desta_connector_bwd"));
  if (!w || !w->layers || !wt || !wt->layers || !taps || !workspace || !grad_out || !grad_arena || dims->qf_layers > 32) {
    set_error("desta_connector_bwd: bad argument");
    return DESTA_ERR_ARG;
  }
  if (N <= 0) return DESTA_OK;
  TrainWs ws = carve_train(dims, N, workspace);
  if (workspace_bytes < ws.bytes) { set_error("desta_connector_bwd: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes); return DESTA_ERR_ARG; }
  GradArena G = carve_grads(dims, grad_arena);
  if (arena_floats < G.floats) { set_error("desta_connector_bwd: gradient arena too small (%zu < %zu floats)", arena_floats, G.floats); return DESTA_ERR_ARG; }
  const int d = dims->d_model, T = dims->max_pos, Q = dims->prompt_size, K = dims->n_taps, H = dims->n_heads, F = dims->qf_ffn;
  const int B = K * N, R = B * Q, ldkv = dims->qf_layers * 2 * d, L = dims->qf_layers, DL = dims->d_llm, NQ = N * Q;
  const float scale = 1.0f / sqrtf((float)(d / H));
  // masks g_f32_a into (g_f32_c fp32, g_bf_a bf16) for a dropped dense branch; returns the fp32 gradient of its bias
  auto drop_grad = [&](int l, int which, const float** gd) -> int {
    desta_dropout dr;
    *gd = ws.g_f32_a;
    if (!site(plan, l, which, &dr)) return DESTA_OK;
    *gd = ws.g_f32_c;
    return desta_dropout_add(ws.g_f32_a, d, nullptr, 0, R, d, &dr, ws.g_f32_c, d, ws.g_bf_a, d, stream);
  };
  auto dgrad = [&](const void* a, int lda, const void* Wt, int M, int Nn, int Kk, void* out_bf16, float* out_f32, const float* resid) {
    return gemm(a, lda, Wt, Kk, M, Nn, Kk, nullptr, out_bf16, Nn, out_f32, Nn, resid, Nn, nullptr, 0, 1.f, 0, 0, stream);
  };
  // ---- proj: out = LN(mix) Wp^T + bp ----
  RUN(desta_cast_f32_bf16(grad_out, ws.g_out_bf, (size_t)NQ * DL, stream));
  RUN(desta_gemm_bf16_wgrad(ws.g_out_bf, DL, ws.mix_bf, d, NQ, DL, d, G.proj_w, d, stream));
  RUN(desta_colsum(grad_out, 0, DL, NQ, DL, G.proj_b, stream));
  RUN(dgrad(ws.g_out_bf, DL, wt->proj, NQ, d, DL, nullptr, ws.g_mixn, nullptr));
  RUN(desta_layernorm_bwd(ws.g_mixn, d, ws.mix, d, w->proj_ln_w, 1e-5f, NQ, d, ws.g_mix, d, nullptr, 0, G.proj_ln_w, G.proj_ln_b, stream));
  // ---- layer mix ----
  RUN(desta_layer_mix_bwd(ws.g_mix, ws.x_out, w->layer_weights, K, N, Q, d, ws.g_x, G.layer_weights, stream));
  const float* dx = ws.g_x;
  for (int l = L - 1; l >= 0; --l) {
    const desta_qf_layer& P = w->layers[l];
    const desta_qf_layer_t& PT = wt->layers[l];
    TrainLayer& S = ws.layer[l];
    GradLayer& g = G.layer[l];
    const float* gd;
    desta_dropout dr;
    // x' = LN(t3)
    RUN(desta_layernorm_bwd(dx, d, S.t3, d, P.f_ln_w, 1e-12f, R, d, ws.g_f32_a, d, ws.g_bf_a, d, g.f_ln_w, g.f_ln_b, stream));
    // t3 = dropout(f W2^T + b2) + c
    RUN(drop_grad(l, 4, &gd));
    RUN(desta_gemm_bf16_wgrad(ws.g_bf_a, d, S.f, F, R, d, F, g.f2_w, F, stream));
    RUN(desta_colsum(gd, 0, d, R, d, g.f2_b, stream));
    RUN(dgrad(ws.g_bf_a, d, PT.f2, R, F, d, ws.g_ff, nullptr, nullptr));                       // df
    RUN(desta_gelu_bwd(ws.g_ff, S.h1, ws.g_h1, (size_t)R * F, stream));                          // dh1
    // h1 = c W1^T + b1
    RUN(desta_gemm_bf16_wgrad(ws.g_h1, F, S.c_bf, d, R, F, d, g.f1_w, d, stream));
    RUN(desta_colsum(ws.g_h1, 1, F, R, F, g.f1_b, stream));
    RUN(dgrad(ws.g_h1, F, PT.f1, R, d, F, nullptr, ws.g_f32_b, ws.g_f32_a));                    // dc = dt3 + dh1 W1
    // c = LN(t2)
    RUN(desta_layernorm_bwd(ws.g_f32_b, d, S.t2, d, P.c_ln_w, 1e-12f, R, d, ws.g_f32_a, d, ws.g_bf_a, d, g.c_ln_w, g.c_ln_b, stream));
    // t2 = dropout(ao_c Wco^T + b) + a
    RUN(drop_grad(l, 3, &gd));
    RUN(desta_gemm_bf16_wgrad(ws.g_bf_a, d, S.ao_c, d, R, d, d, g.c_o_w, d, stream));
    RUN(desta_colsum(gd, 0, d, R, d, g.c_o_b, stream));
    RUN(dgrad(ws.g_bf_a, d, PT.c_o, R, d, d, ws.g_bf_b, nullptr, nullptr));                     // d(ao_c)
    const __nv_bfloat16* kl = ws.kv + (size_t)l * 2 * d;
    __nv_bfloat16* dkl = ws.dkv + (size_t)l * 2 * d;
    // value bias of the K/V projection: O = P~ (X Wv^T + 1 bv^T)  ->  d bv = sum_q (row sum of P~)_q dO_q, a reduction over dO
    // (row sums: 1 without dropout, rr_c with it) instead of a column sum over this layer's 490 MB of dV
    RUN(desta_colsum_heads(ws.g_bf_b, d, site(plan, l, 2, &dr) ? S.rr_c : nullptr, B, H, Q, G.kv_b + (size_t)l * 2 * d + d, stream));
    RUN(desta_attention_bwd(S.q_c, d, kl, ldkv, kl + d, ldkv, S.ao_c, d, ws.g_bf_b, d, S.lse_c, ws.g_qkv, 3 * d, dkl, ldkv, dkl + d, ldkv,
                            B, H, Q, T, scale, site(plan, l, 2, &dr) ? &dr : nullptr, stream));
    // q_c = a Wq^T + bq   (dq_c = first d columns of g_qkv)
    RUN(desta_gemm_bf16_wgrad(ws.g_qkv, 3 * d, S.a_bf, d, R, d, d, g.c_q_w, d, stream));
    RUN(desta_colsum(ws.g_qkv, 1, 3 * d, R, d, g.c_q_b, stream));
    RUN(dgrad(ws.g_qkv, 3 * d, PT.c_q, R, d, d, nullptr, ws.g_f32_b, ws.g_f32_a));              // da = dt2 + dq_c Wq
    // a = LN(t1)
    RUN(desta_layernorm_bwd(ws.g_f32_b, d, S.t1, d, P.s_ln_w, 1e-12f, R, d, ws.g_f32_a, d, ws.g_bf_a, d, g.s_ln_w, g.s_ln_b, stream));
    // t1 = dropout(ao_s Wso^T + b) + x
    RUN(drop_grad(l, 1, &gd));
    RUN(desta_gemm_bf16_wgrad(ws.g_bf_a, d, S.ao_s, d, R, d, d, g.s_o_w, d, stream));
    RUN(desta_colsum(gd, 0, d, R, d, g.s_o_b, stream));
    RUN(dgrad(ws.g_bf_a, d, PT.s_o, R, d, d, ws.g_bf_b, nullptr, nullptr));                     // d(ao_s)
    RUN(desta_attention_bwd(S.qkv, 3 * d, S.qkv + d, 3 * d, S.qkv + 2 * d, 3 * d, S.ao_s, d, ws.g_bf_b, d, S.lse_s, ws.g_qkv, 3 * d,
                            ws.g_qkv + d, 3 * d, ws.g_qkv + 2 * d, 3 * d, B, H, Q, Q, scale, site(plan, l, 0, &dr) ? &dr : nullptr, stream));
    // qkv = x Wqkv^T + b
    RUN(desta_gemm_bf16_wgrad(ws.g_qkv, 3 * d, S.x_bf, d, R, 3 * d, d, g.s_qkv_w, d, stream));
    RUN(desta_colsum(ws.g_qkv, 1, 3 * d, R, 3 * d, g.s_qkv_b, stream));
    RUN(dgrad(ws.g_qkv, 3 * d, PT.s_qkv, R, d, 3 * d, nullptr, ws.g_x, ws.g_f32_a));            // dx = dt1 + dqkv Wqkv
    dx = ws.g_x;
    if (on_ready) on_ready(user, g.first, g.count);
  }
  // ---- learned prompts: x0[k, n] = layer_prompts[k] for every clip n ----
  for (int k = 0; k < K; ++k)
    RUN(desta_colsum(dx + (size_t)k * N * Q * d, 0, Q * d, N, Q * d, G.prompts + (size_t)k * Q * d, stream));
  // ---- cross-attention K/V projections of all layers: one weight-gradient GEMM over all taps' rows ----
  RUN(desta_gemm_bf16_wgrad(ws.dkv, ldkv, taps, d, K * N * T, ldkv, d, G.kv_w, d, stream));
  // bias gradients of those projections: no pass over dK / dV - a key bias shifts every score of a row alike, the softmax
  // ignores it and its gradient is identically zero (the arena is zeroed); the value-bias gradient was taken from dO per layer
  // (desta_colsum_heads in the loop above)
  if (on_ready) on_ready(user, 0, G.global_count);
  return DESTA_OK;
}

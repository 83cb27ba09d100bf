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

struct EncWs { __nv_bfloat16 *cols1, *c1, *cols2, *y, *qkv, *ao, *ff; size_t bytes; };
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
static int gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out_bf16, int ld_bf16,
                float* out_f32, int ld_f32, const float* resid, int ld_resid, const float* pos, int pos_rows, float scale,
                int scale_cols, int gelu, void* stream) {
  desta_gemm_epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32; ep.out_bf16 = out_bf16; ep.ld_bf16 = ld_bf16; ep.bias = bias;
  ep.resid = resid; ep.ld_resid = ld_resid; ep.pos = pos; ep.pos_rows = pos_rows; ep.scale = scale; ep.scale_cols = scale_cols;
  ep.gelu = gelu;
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

extern "C" int desta_encoder_fwd(const desta_dims* dims, const desta_encoder_weights* w, const void* mel_tm, int N, void* workspace,
                                 size_t workspace_bytes, void* taps, float* h, void* stream) {
  RUN(check_dims(dims, "desta_encoder_fwd"));
  if (!w || !w->layers || !mel_tm || !workspace || !taps || !h) { set_error("desta_encoder_fwd: null pointer"); return DESTA_ERR_ARG; }
  if (N <= 0) return DESTA_OK;
  const EncWs ws = carve_encoder(dims, N, workspace);
  if (workspace_bytes < ws.bytes) { set_error("desta_encoder_fwd: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes); return DESTA_ERR_ARG; }
  const int d = dims->d_model, T = dims->max_pos, M = N * T, H = dims->n_heads, F = dims->ffn;
  const float q_scale = 1.0f / sqrtf((float)(d / H));
  // conv stem as im2col + GEMM: gelu(conv1), gelu(conv2 stride 2) + positions -> fp32 residual stream (:563-570)
  RUN(desta_im2col_k3(mel_tm, N, 2 * T, dims->n_mels, 1, ws.cols1, stream));
  RUN(gemm(ws.cols1, 3 * dims->n_mels, w->conv1_w, 3 * dims->n_mels, 2 * M, d, 3 * dims->n_mels, w->conv1_b, ws.c1, d, nullptr, 0,
           nullptr, 0, nullptr, 0, 1.f, 0, 1, stream));
  RUN(desta_im2col_k3(ws.c1, N, 2 * T, d, 2, ws.cols2, stream));
  RUN(gemm(ws.cols2, 3 * d, w->conv2_w, 3 * d, M, d, 3 * d, w->conv2_b, nullptr, 0, h, d, nullptr, 0, w->pos, T, 1.f, 0, 1, stream));
  int tap = 0;
  for (int i = 0; i < dims->n_layers; ++i) {
    const desta_enc_layer& L = w->layers[i];
    RUN(desta_layernorm(h, d, L.ln1_w, L.ln1_b, 1e-5f, M, d, ws.y, d, nullptr, 0, stream));
    RUN(gemm(ws.y, d, L.qkv_w, d, M, 3 * d, d, L.qkv_b, ws.qkv, 3 * d, nullptr, 0, nullptr, 0, nullptr, 0, q_scale, d, 0, stream));
    RUN(desta_attention(ws.qkv, 3 * d, ws.qkv + d, 3 * d, ws.qkv + 2 * d, 3 * d, ws.ao, d, N, H, T, T, 1.0f, nullptr, nullptr, stream));
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

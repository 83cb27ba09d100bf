"""Whisper encoder forward with per-layer taps — fp32 torch restatement.  TEST INFRASTRUCTURE ONLY.

Follows ref: desta/models/modeling_desta25.py:544-608 (forward_whisper: conv stem :563-570,
layer loop :577-598, taps after target layers :587; the final encoder.layer_norm is NOT applied)
and transformers/models/whisper/modeling_whisper.py:380-414 (pre-LN layer), :284-358 (attention:
q scaled by head_dim**-0.5 after bias, k_proj bias-free, no mask), :215-238 (eager attention)."""
import torch
import torch.nn.functional as F

from .weights import ENC


def conv_stem(W, mel):
    h = F.gelu(F.conv1d(mel, W[ENC + "conv1.weight"], W[ENC + "conv1.bias"], padding=1))
    h = F.gelu(F.conv1d(h, W[ENC + "conv2.weight"], W[ENC + "conv2.bias"], stride=2, padding=1))
    h = h.permute(0, 2, 1)
    return h + W[ENC + "embed_positions.weight"][: h.shape[1]]


def encoder_layer(W, i, h, n_heads):
    p = f"{ENC}layers.{i}."
    B, T, d = h.shape
    hd = d // n_heads
    y = F.layer_norm(h, (d,), W[p + "self_attn_layer_norm.weight"], W[p + "self_attn_layer_norm.bias"], 1e-5)
    q = F.linear(y, W[p + "self_attn.q_proj.weight"], W[p + "self_attn.q_proj.bias"]) * hd ** -0.5
    k = F.linear(y, W[p + "self_attn.k_proj.weight"])
    v = F.linear(y, W[p + "self_attn.v_proj.weight"], W[p + "self_attn.v_proj.bias"])
    q, k, v = (t.view(B, T, n_heads, hd).transpose(1, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, scale=1.0)
    o = o.transpose(1, 2).reshape(B, T, d)
    h = h + F.linear(o, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
    y = F.layer_norm(h, (d,), W[p + "final_layer_norm.weight"], W[p + "final_layer_norm.bias"], 1e-5)
    y = F.gelu(F.linear(y, W[p + "fc1.weight"], W[p + "fc1.bias"]))
    return h + F.linear(y, W[p + "fc2.weight"], W[p + "fc2.bias"])


def encoder_taps(W, cfg, mel):
    """mel [N, n_mels, 3000] fp32 -> list of taps [N, 1500, d] (hidden after each tap layer)."""
    if mel.shape[-1] != 2 * cfg.max_pos:
        raise ValueError(f"Whisper expects the mel input features to be of length {2 * cfg.max_pos}, "
                         f"but found {mel.shape[-1]}.")  # modeling_desta25.py:557-560
    h = conv_stem(W, mel)
    taps = []
    for i in range(cfg.n_layers):
        h = encoder_layer(W, i, h, cfg.n_heads)
        if i in cfg.tap_layers:
            taps.append(h)
    return taps

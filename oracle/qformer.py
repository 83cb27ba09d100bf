"""Q-Former connector forward — fp32 torch restatement (autograd gives the backward oracle).
TEST INFRASTRUCTURE ONLY.

Follows ref: desta/models/modeling_desta25.py:587-606 (inline use in forward_whisper), :148-168 (params)
and transformers/models/bert/modeling_bert.py:379-421 (BertLayer, post-LN, is_decoder + cross-attn),
:168-207 / :234-284 (self / cross attention, eager, no mask, scale head_dim**-0.5),
:287-298, :345-356 (dense + LayerNorm(x + residual), eps 1e-12), :330-342 (erf GELU FFN).
Dropout (BertConfig default 0.1, train mode only): `drop(site, x)` is called at the five nn.Dropout sites of a
BertLayer in the reference's order (:198-201 self probabilities, :296 self output, :279 cross probabilities, :296 cross
output, :354 FFN output) with site = (tap, layer, which).  `drop=None` is eval mode; `torch_dropout` replays torch's own
RNG stream (pins the site order against the reference); `mask_dropout` applies explicit keep masks (GPU parity)."""
import torch
import torch.nn.functional as F

from .weights import CON

SELF_PROBS, SELF_OUT, CROSS_PROBS, CROSS_OUT, FFN_OUT = range(5)


def torch_dropout(p_hidden=0.1, p_attn=0.1):
    def drop(site, x):
        return F.dropout(x, p_attn if site[2] in (SELF_PROBS, CROSS_PROBS) else p_hidden, training=True)
    return drop


def mask_dropout(masks, scales):
    """masks[(tap, layer, which)] = bool keep mask shaped like the dropped tensor; scales[which] = 1 / (1 - p)."""
    def drop(site, x):
        return x * masks[site].to(x.dtype) * scales[site[2]]
    return drop


def _mha(W, p, x, kv, n_heads, drop=None, site_probs=None, site_out=None):
    B, Tq, d = x.shape
    hd = d // n_heads
    q = F.linear(x, W[p + "self.query.weight"], W[p + "self.query.bias"])
    k = F.linear(kv, W[p + "self.key.weight"], W[p + "self.key.bias"])
    v = F.linear(kv, W[p + "self.value.weight"], W[p + "self.value.bias"])
    q = q.view(B, Tq, n_heads, hd).transpose(1, 2)
    k = k.view(B, -1, n_heads, hd).transpose(1, 2)
    v = v.view(B, -1, n_heads, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5
    a = torch.softmax(s, dim=-1)
    if drop is not None:
        a = drop(site_probs, a)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, Tq, d)
    o = F.linear(o, W[p + "output.dense.weight"], W[p + "output.dense.bias"])
    if drop is not None:
        o = drop(site_out, o)
    return F.layer_norm(o + x, (d,), W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], 1e-12)


def qformer_layer(W, l, x, tap, n_heads, drop=None, k=0):
    p = f"{CON}qformer.layer.{l}."
    d = x.shape[-1]
    a = _mha(W, p + "attention.", x, x, n_heads, drop, (k, l, SELF_PROBS), (k, l, SELF_OUT))
    c = _mha(W, p + "crossattention.", a, tap, n_heads, drop, (k, l, CROSS_PROBS), (k, l, CROSS_OUT))
    y = F.gelu(F.linear(c, W[p + "intermediate.dense.weight"], W[p + "intermediate.dense.bias"]))
    y = F.linear(y, W[p + "output.dense.weight"], W[p + "output.dense.bias"])
    if drop is not None:
        y = drop((k, l, FFN_OUT), y)
    return F.layer_norm(y + c, (d,), W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], 1e-12)


def connector(W, cfg, taps, drop=None):
    """taps: list of [N, 1500, d] -> audio features [N, prompt_size, d_llm]."""
    N = taps[0].shape[0]
    outs = []
    for k, tap in enumerate(taps):
        x = W[f"{CON}layer_prompts.{k}"].expand(N, -1, -1)
        for l in range(cfg.qf_layers):
            x = qformer_layer(W, l, x, tap, cfg.n_heads, drop, k)
        outs.append(x[:, : cfg.prompt_size])
    stack = torch.stack(outs, dim=0).permute(1, 2, 0, 3)             # [N, Q, taps, d]
    w = torch.softmax(W[CON + "layer_weights"], dim=-1).unsqueeze(-1)  # [Q, taps, 1]
    m = (stack * w).sum(dim=2)
    d = m.shape[-1]
    m = F.layer_norm(m, (d,), W[CON + "proj.0.weight"], W[CON + "proj.0.bias"], 1e-5)
    return F.linear(m, W[CON + "proj.1.weight"], W[CON + "proj.1.bias"])

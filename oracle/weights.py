"""Deterministic random weights keyed by the reference's checkpoint names.  TEST INFRASTRUCTURE ONLY.

Names follow the HF state-dict layout the reference saves/loads
(ref: desta/models/modeling_desta25.py:1284-1292 state_dict, :148-168 connector params;
SURVEY.md Appendix A.3/A.4).  Biases and LayerNorm affine terms are non-trivial on purpose so that
every epilogue term is exercised by parity tests (HF default init would zero them)."""
import torch

ENC = "perception.whisper.model.encoder."
CON = "perception.connector."


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_weights(cfg, seed: int = 0, w_std: float = 0.02, with_table: bool = True):
    g = _gen(seed)
    W = {}

    def lin(name, out_f, in_f, bias=True, std=w_std):
        W[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
        if bias:
            W[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def ln(name, d):
        W[name + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        W[name + ".bias"] = 0.05 * torch.randn(d, generator=g)

    d, f = cfg.d_model, cfg.ffn
    W[ENC + "conv1.weight"] = torch.randn(d, cfg.n_mels, 3, generator=g) * (1.0 / (cfg.n_mels * 3) ** 0.5)
    W[ENC + "conv1.bias"] = torch.randn(d, generator=g) * 0.02
    W[ENC + "conv2.weight"] = torch.randn(d, d, 3, generator=g) * (1.0 / (d * 3) ** 0.5)
    W[ENC + "conv2.bias"] = torch.randn(d, generator=g) * 0.02
    W[ENC + "embed_positions.weight"] = torch.randn(cfg.max_pos, d, generator=g) * 0.1
    for i in range(cfg.n_layers):
        p = f"{ENC}layers.{i}."
        lin(p + "self_attn.q_proj", d, d)
        lin(p + "self_attn.k_proj", d, d, bias=False)
        lin(p + "self_attn.v_proj", d, d)
        lin(p + "self_attn.out_proj", d, d)
        ln(p + "self_attn_layer_norm", d)
        lin(p + "fc1", f, d)
        lin(p + "fc2", d, f)
        ln(p + "final_layer_norm", d)
    W[CON + "layer_weights"] = 0.5 * torch.randn(cfg.prompt_size, len(cfg.tap_layers), generator=g)
    for k in range(len(cfg.tap_layers)):
        W[f"{CON}layer_prompts.{k}"] = torch.randn(1, cfg.prompt_size, d, generator=g)
    for l in range(cfg.qf_layers):
        p = f"{CON}qformer.layer.{l}."
        for att in ("attention", "crossattention"):
            lin(p + att + ".self.query", d, d)
            lin(p + att + ".self.key", d, d)
            lin(p + att + ".self.value", d, d)
            lin(p + att + ".output.dense", d, d)
            ln(p + att + ".output.LayerNorm", d)
        lin(p + "intermediate.dense", cfg.qf_ffn, d)
        lin(p + "output.dense", d, cfg.qf_ffn)
        ln(p + "output.LayerNorm", d)
    ln(CON + "proj.0", d)
    lin(CON + "proj.1", cfg.d_llm, d)
    if with_table:
        W["llm_model.model.embed_tokens.weight"] = (torch.randn(cfg.vocab, cfg.d_llm, generator=g) * 0.02).to(torch.bfloat16)
    return W

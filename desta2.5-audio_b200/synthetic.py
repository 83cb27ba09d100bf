"""Random-init containers of the named architectures, created directly on the device (no checkpoints, no network):
used by bench.py and the GPU tests.  HF modules are instantiated on the meta device and materialised on the GPU."""
import types

import torch
import torch.nn as nn

from .connector import QformerConnector
from .perception import WhisperPerceptionB200


class _WhisperShell(nn.Module):
    """`.model.encoder` with the HF parameter names; the decoder is irrelevant to this path and left out."""

    def __init__(self, encoder):
        super().__init__()
        self.model = nn.Module()
        self.model.encoder = encoder


@torch.no_grad()
def _init(module, seed):
    g = torch.Generator(device=next(module.parameters()).device)
    g.manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() == 1:
            if name.endswith("weight"):      # LayerNorm gain
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=p.device))
        elif "layer_prompts" in name:
            p.copy_(torch.randn(p.shape, generator=g, device=p.device))
        elif "layer_weights" in name:
            p.zero_()
        elif "embed_positions" in name:
            p.copy_(0.1 * torch.randn(p.shape, generator=g, device=p.device))
        elif "conv" in name:
            fan_in = p.shape[1] * p.shape[2]
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * fan_in ** -0.5)
        else:
            p.copy_(0.02 * torch.randn(p.shape, generator=g, device=p.device))


def build_perception(n_mels, d_model, n_heads, ffn, n_layers, tap_layers, qf_layers, prompt_size, d_llm, device, seed=0,
                     max_pos=1500):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    wcfg = WhisperConfig(num_mel_bins=n_mels, d_model=d_model, encoder_layers=n_layers, encoder_attention_heads=n_heads,
                         encoder_ffn_dim=ffn, max_source_positions=max_pos, decoder_layers=1, decoder_attention_heads=n_heads,
                         decoder_ffn_dim=64, vocab_size=64, max_target_positions=8, pad_token_id=0, bos_token_id=1,
                         eos_token_id=2, decoder_start_token_id=1)
    with torch.device("meta"):
        enc = WhisperEncoder(wcfg)
        con = QformerConnector(d_model, n_heads, d_llm, prompt_size, qf_layers, len(tap_layers))
    enc = enc.to_empty(device=device)
    con = con.to_empty(device=device)
    _init(enc, seed)
    _init(con, seed + 1)
    for p in enc.parameters():
        p.requires_grad_(False)
    return WhisperPerceptionB200(_WhisperShell(enc), con, tap_layers, prompt_size).eval()

"""Parameter container mirroring the reference's QformerConnector (same attribute names, shapes and init), so that
`state_dict()` / `model.safetensors` keys stay `perception.connector.{layer_weights, layer_prompts.N, qformer.layer.L.*,
proj.{0,1}.*}` (ref: desta/models/modeling_desta25.py:126-168, checkpoint layout SURVEY.md §3.4).

Only the container is shared with the reference; the arithmetic runs in desta_b200.engine on libdesta_b200.so.
"""
import torch
import torch.nn as nn

TAP_LAYERS = {  # modeling_desta25.py:132-143
    "openai/whisper-medium": [5, 11, 17, 23],
    "openai/whisper-small": [2, 5, 8, 11],
    "openai/whisper-tiny": [0, 1, 2, 3],
    "openai/whisper-large-v3": [7, 15, 23, 31],
    "openai/whisper-large-v3-turbo": [7, 15, 23, 31],
}


def tap_layers_for(encoder_model_id: str):
    if encoder_model_id not in TAP_LAYERS:
        raise NotImplementedError(f"model_id {encoder_model_id} not implemented")   # same error as the reference
    return list(TAP_LAYERS[encoder_model_id])


class QformerConnector(nn.Module):
    def __init__(self, d_model: int, n_heads: int, d_llm: int, prompt_size: int = 64, qformer_num_hidden_layers: int = 2,
                 n_taps: int = 4):
        super().__init__()
        from transformers import BertConfig
        from transformers.models.bert.modeling_bert import BertEncoder
        self.layer_prompts = nn.ParameterList([nn.Parameter(torch.randn(1, prompt_size, d_model)) for _ in range(n_taps)])
        self.layer_weights = nn.Parameter(torch.zeros(prompt_size, n_taps, dtype=torch.float))
        cfg = BertConfig()
        cfg.num_hidden_layers = qformer_num_hidden_layers
        cfg.num_attention_heads = n_heads
        cfg.hidden_size = d_model
        cfg.add_cross_attention = True
        cfg.is_decoder = True
        cfg._attn_implementation = "eager"
        self.qformer = BertEncoder(cfg)          # parameter container only: its forward is never called here
        self.proj = nn.Sequential(nn.LayerNorm(d_model), nn.Linear(d_model, d_llm))

    def forward(self, *a, **k):
        raise RuntimeError("desta_b200.QformerConnector is a parameter container; call WhisperPerceptionB200")

"""Shape configs for the path (explicit: there is no network / HF hub here).

ref: SURVEY.md §8 flagship shapes; examples/train/config/desta25_debug.yaml:36-46 (debug);
desta/models/modeling_desta25.py:132-146 (tap layer ids per encoder)."""
from dataclasses import dataclass, field
from typing import List


@dataclass
class PathConfig:
    name: str
    n_mels: int
    d_model: int
    n_heads: int
    ffn: int
    n_layers: int
    tap_layers: List[int]
    qf_layers: int
    qf_ffn: int = 3072          # BertConfig default intermediate_size, never overridden (:156-162)
    prompt_size: int = 64
    d_llm: int = 4096
    vocab: int = 128256
    max_pos: int = 1500
    encoder_model_id: str = ""

    @property
    def head_dim(self):
        return self.d_model // self.n_heads


CONFIGS = {
    # 4-layer toy with Whisper's structure: used for committed golden vectors (small weights)
    "micro": PathConfig("micro", 80, 128, 2, 256, 4, [0, 1, 2, 3], 2, qf_ffn=3072, prompt_size=64, d_llm=256, vocab=512,
                        encoder_model_id="openai/whisper-tiny"),
    # the flagship's exact layer shapes (d 1280, 20 heads, ffn 5120, 128 mels, 1500 positions, d_llm 4096) with only two
    # encoder layers / one Q-Former layer so that the CPU oracle finishes in seconds: pins every flagship GEMM / attention
    # shape against the oracle, which the 32-layer model cannot do within the test budget
    "l3_slim": PathConfig("l3_slim", 128, 1280, 20, 5120, 2, [0, 1], 1, d_llm=4096, vocab=1024,
                          encoder_model_id="openai/whisper-large-v3"),
    # openai/whisper-tiny + Qformer 2L -> Qwen3-0.6B dims (desta25_debug.yaml)
    "debug": PathConfig("debug", 80, 384, 6, 1536, 4, [0, 1, 2, 3], 2, d_llm=1024, vocab=151936,
                        encoder_model_id="openai/whisper-tiny"),
    # whisper-large-v3 + Qformer 6L -> Llama-3.1-8B
    "large_v3_llama": PathConfig("large_v3_llama", 128, 1280, 20, 5120, 32, [7, 15, 23, 31], 6, d_llm=4096,
                                 vocab=128256, encoder_model_id="openai/whisper-large-v3"),
    # whisper-large-v3-turbo (same encoder) + Qformer 6L -> Llama-3.1-8B
    "turbo_llama": PathConfig("turbo_llama", 128, 1280, 20, 5120, 32, [7, 15, 23, 31], 6, d_llm=4096,
                              vocab=128256, encoder_model_id="openai/whisper-large-v3-turbo"),
    # Qwen3-4B backbone (docs/train_qwen3.md)
    "large_v3_qwen3_4b": PathConfig("large_v3_qwen3_4b", 128, 1280, 20, 5120, 32, [7, 15, 23, 31], 6, d_llm=2560,
                                    vocab=151936, encoder_model_id="openai/whisper-large-v3"),
}

"""Import the reference's own model code in the BUILD CONTAINER (no network, no HF cache).
TEST INFRASTRUCTURE ONLY — used by oracle/make_golden.py; never on the GPU box (/root/reference is absent there).

Three shims: (1) stub librosa/soundfile AFTER importing transformers
(ref: desta/utils/audio.py:41-43,56 import them at module import); (2) from_pretrained of AutoConfig /
WhisperForConditionalGeneration / AutoModelForCausalLM build random-init models from explicit configs
(ref: modeling_desta25.py:505-506, 673-674, 713-718 would hit the hub); (3) WhisperEncoderLayer.forward
returns a 1-tuple as under transformers 4.x, which the reference indexes with [0] (:585)."""
import sys
import types

REF_ROOT = "/root/reference"


def load_reference(cfg):
    import torch
    import transformers  # noqa: F401  (must be imported before stubbing)
    import transformers.audio_utils  # noqa: F401
    from transformers import (AutoConfig, AutoModelForCausalLM, AutoProcessor, AutoTokenizer,  # noqa: F401
                              LlamaConfig, LlamaForCausalLM, WhisperConfig, WhisperForConditionalGeneration)
    from transformers.models.whisper import modeling_whisper

    for name in ("librosa", "soundfile"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.available_formats = lambda: {}
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    wcfg = WhisperConfig(
        num_mel_bins=cfg.n_mels, d_model=cfg.d_model, encoder_layers=cfg.n_layers,
        encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn, decoder_layers=1,
        decoder_attention_heads=cfg.n_heads, decoder_ffn_dim=64, max_source_positions=cfg.max_pos,
        vocab_size=64, max_target_positions=8, pad_token_id=0, bos_token_id=1, eos_token_id=2,
        decoder_start_token_id=1)
    lcfg = LlamaConfig(hidden_size=cfg.d_llm, intermediate_size=64, num_hidden_layers=1,
                       num_attention_heads=2, num_key_value_heads=2, vocab_size=cfg.vocab,
                       max_position_embeddings=1024)

    def auto_config(model_id, *a, **k):
        return wcfg if "whisper" in model_id else lcfg

    AutoConfig.from_pretrained = staticmethod(auto_config)
    WhisperForConditionalGeneration.from_pretrained = classmethod(
        lambda cls, *a, **k: WhisperForConditionalGeneration(wcfg))
    AutoModelForCausalLM.from_pretrained = staticmethod(
        lambda *a, **k: LlamaForCausalLM(lcfg).to(torch.bfloat16))

    if not getattr(modeling_whisper.WhisperEncoderLayer, "_tuple_shim", False):
        orig = modeling_whisper.WhisperEncoderLayer.forward

        def fwd(self, *a, **k):
            out = orig(self, *a, **k)
            return out if isinstance(out, tuple) else (out,)

        modeling_whisper.WhisperEncoderLayer.forward = fwd
        modeling_whisper.WhisperEncoderLayer._tuple_shim = True

    from desta.models import modeling_desta25 as ref
    return ref

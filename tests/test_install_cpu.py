"""install() against the REAL reference model class (build container only: skipped where /root/reference is absent)."""
import os

import numpy as np
import pytest
import torch

from oracle import refshim

REF = (refshim.ref_root() or "/nonexistent") + "/desta/models/modeling_desta25.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this machine")
def test_install_keeps_reference_checkpoint_layout():
    import desta_b200
    from desta_b200.perception import WhisperPerceptionB200
    from oracle.configs import CONFIGS
    from oracle.refshim import load_reference
    cfg = CONFIGS["micro"]
    ref = load_reference(cfg)
    rcfg = ref.DeSTA25Config(llm_model_id="llama-micro", encoder_model_id=cfg.encoder_model_id,
                             qformer_num_hidden_layers=cfg.qf_layers, prompt_size=cfg.prompt_size)
    model = ref.DeSTA25AudioModel(rcfg)
    keys_before = sorted(model.state_dict().keys())
    trainable_before = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    desta_b200.install(model)
    assert isinstance(model.perception, WhisperPerceptionB200)
    assert sorted(model.state_dict().keys()) == keys_before                    # trainable-only state_dict (:1284-1292)
    assert all(k.startswith("perception.connector.") for k in keys_before)
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == trainable_before
    assert model.perception.dims.tap_layers == [0, 1, 2, 3] and model.perception.dims.prompt_size == cfg.prompt_size
    # the product path has no CPU fallback: calling it without CUDA must fail loudly, not silently run elsewhere
    with pytest.raises((AssertionError, RuntimeError)):
        model.perception(input_features=torch.zeros(1, cfg.n_mels, 3000))


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this machine")
def test_install_patch_audio_rebinds_the_name_generate_uses():
    """install(model, patch_audio=True) swaps the module-level `AudioSegment` that generate() calls
    (modeling_desta25.py:13, :1553-1557) for the GPU decode + resample mirror; the default install leaves it alone."""
    import sys

    import desta_b200
    from desta_b200 import audio
    from oracle.configs import CONFIGS
    from oracle.refshim import load_reference
    cfg = CONFIGS["micro"]
    ref = load_reference(cfg)
    rcfg = ref.DeSTA25Config(llm_model_id="llama-micro", encoder_model_id=cfg.encoder_model_id,
                             qformer_num_hidden_layers=cfg.qf_layers, prompt_size=cfg.prompt_size)
    mod = sys.modules[ref.DeSTA25AudioModel.__module__]
    original = mod.AudioSegment
    try:
        model = ref.DeSTA25AudioModel(rcfg)
        desta_b200.install(model)
        assert mod.AudioSegment is original
        model2 = ref.DeSTA25AudioModel(rcfg)
        desta_b200.install(model2, patch_audio=True)
        assert mod.AudioSegment is audio.AudioSegment
        # same call shape as the reference's: classmethod from_file(path, target_sr=..., channel_selector=...) -> .samples
        assert callable(mod.AudioSegment.from_file) and hasattr(mod.AudioSegment, "samples")
    finally:
        mod.AudioSegment = original


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this machine")
def test_install_wraps_the_processor_generate_creates_lazily():
    """The documented order is from_pretrained -> install -> generate, and the reference creates `self.processor` inside
    generate() (`_setup_generation`, modeling_desta25.py:1467-1482, called at :1514-1515): install() must catch it."""
    import desta_b200
    from desta_b200.processor import WaveformProcessor
    from oracle.configs import CONFIGS
    cfg = CONFIGS["micro"]
    ref = refshim.load_reference(cfg)
    rcfg = ref.DeSTA25Config(llm_model_id="llama-micro", encoder_model_id=cfg.encoder_model_id,
                             qformer_num_hidden_layers=cfg.qf_layers, prompt_size=cfg.prompt_size)
    model = ref.DeSTA25AudioModel(rcfg)
    refshim.offline_generation(model, cfg.n_mels)
    assert not hasattr(model, "processor") and not hasattr(model, "tokenizer")
    desta_b200.install(model)
    desta_b200.install(model)                                  # idempotent: no double wrapping
    model._setup_generation()                                  # what generate() does on its first call
    assert isinstance(model.processor, WaveformProcessor) and not isinstance(model.processor._wrapped, WaveformProcessor)
    feats = model.processor([np.zeros(1600, np.float32), np.ones(500000, np.float32)], sampling_rate=16000,
                            return_tensors="pt").input_features
    assert feats.shape == (2, 480000) and feats.dtype == torch.float32 and model.processor._wrapped.calls == 0
    assert model.processor.batch_decode([[5, 6, 7]], skip_special_tokens=True) == model.processor._wrapped.batch_decode([[5, 6, 7]])
    # a model that already has a processor (install after the first generate) is wrapped on the spot
    model2 = ref.DeSTA25AudioModel(rcfg)
    refshim.offline_generation(model2, cfg.n_mels)
    model2._setup_generation()
    desta_b200.install(model2)
    assert isinstance(model2.processor, WaveformProcessor)

"""Scope row n2: the ASR pass of generate() (ref: desta/models/modeling_desta25.py:1582-1601,
`self.perception.whisper.generate(input_features=...)`) re-encodes the clips the perception call encodes again.
`WhisperPerceptionB200.encoder_outputs` yields the HF-compatible encoder output (final LayerNorm applied) from the same
kernels, `wrap_whisper_generate` feeds it to HF's generate, and the next perception call on the same clips skips its
encoder run.  Tolerance as for the taps: max|delta|/max|ref| <= 2e-2 and cosine >= 0.999."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import logmel as o_mel  # noqa: E402
from oracle.configs import CONFIGS  # noqa: E402
from oracle.synth import synth_waves  # noqa: E402
from oracle.weights import make_weights  # noqa: E402
from test_path_gpu import KINDS, _cos, _rel, build_perception  # noqa: E402


@pytest.mark.parametrize("name", ["micro", "debug"])
def test_encoder_outputs_match_hf_encoder_and_asr_shares_the_run(name):
    cfg = CONFIGS[name]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)[:3]
    mel = torch.from_numpy(o_mel.logmel(waves, cfg.n_mels))
    p = build_perception(cfg, W)
    with torch.no_grad():                      # non-trivial final LayerNorm (HF default init is the identity affine)
        p.whisper.model.encoder.layer_norm.weight.normal_(1.0, 0.1)
        p.whisper.model.encoder.layer_norm.bias.normal_(0.0, 0.05)
    p._enc_pack = None
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    with torch.no_grad():
        ref = p.whisper.model.encoder(mel.cuda()).last_hidden_state          # HF's own encoder (fp32, eager)
        ours = p.encoder_outputs(wave).last_hidden_state
        ours_from_mel = p.encoder_outputs(mel.cuda()).last_hidden_state      # reference-style log-mel input
    assert ours.shape == ref.shape == (3, cfg.max_pos, cfg.d_model)
    for got in (ours, ours_from_mel):
        rel, cos = _rel(got, ref), _cos(got, ref)
        assert rel <= 2e-2 and cos >= 0.999, f"encoder output: rel {rel:.4f} cos {cos:.6f}"

    # the ASR pass through the wrapped generate: same signature as the reference's call site, waveform features
    with torch.no_grad():
        feats_plain, _ = p(input_features=wave)
        feats_plain = feats_plain.clone()
        tok_ref = p.whisper.generate(input_features=mel.cuda(), attention_mask=None, max_new_tokens=6)
    p.wrap_whisper_generate()
    p.wrap_whisper_generate()                                                # idempotent
    skipped0 = getattr(p, "encoder_runs_skipped", 0)
    with torch.no_grad():
        tok = p.whisper.generate(input_features=wave, attention_mask=None, max_new_tokens=6)
        assert tok.shape == tok_ref.shape
        feats_shared, _ = p(input_features=wave)                             # same clips: encoder state is reused
        assert getattr(p, "encoder_runs_skipped", 0) == skipped0 + 1
        assert torch.equal(feats_shared, feats_plain)
        feats_again, _ = p(input_features=wave)                              # the share is one-shot
        assert getattr(p, "encoder_runs_skipped", 0) == skipped0 + 1
        assert torch.equal(feats_again, feats_plain)
        # the same clips in another order after an ASR pass: matched PER CLIP (content hash + bit-for-bit confirmation)
        p.whisper.generate(input_features=wave, attention_mask=None, max_new_tokens=2)
        other = wave.flip(0).contiguous()
        feats_other, _ = p(input_features=other)
        assert getattr(p, "encoder_runs_skipped", 0) == skipped0 + 2
        assert torch.equal(feats_other.flip(0), feats_plain)
        # genuinely different clips: nothing is reused
        p.whisper.generate(input_features=wave, attention_mask=None, max_new_tokens=2)
        reused0 = getattr(p, "encoder_clips_reused", 0)
        noise = (wave + 0.001).contiguous()
        p(input_features=noise)
        assert getattr(p, "encoder_clips_reused", 0) == reused0 and getattr(p, "encoder_runs_skipped", 0) == skipped0 + 2


def test_asr_subset_is_not_encoded_twice():
    """generate() runs the ASR pass on the SUBSET of clips that are speech without a transcript and the perception call on
    ALL clips (modeling_desta25.py:1562-1601): only the clips outside the subset go through the encoder again, and the
    features are bit-identical to encoding the whole batch."""
    from desta_b200 import ops
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)[:5]
    p = build_perception(cfg, W)
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    with torch.no_grad():
        feats_plain, _ = p(input_features=wave)
        feats_plain = feats_plain.clone()
        taps_plain = [t.clone() for t in p.taps(wave)]
    p.wrap_whisper_generate()
    subset = [3, 1]                                                          # ASR batch: clips 3 and 1, in that order
    with torch.no_grad():
        p.whisper.generate(input_features=wave[subset].contiguous(), attention_mask=None, max_new_tokens=3)
        c0, r0 = ops.STATS.get("encoder_clips", 0), getattr(p, "encoder_clips_reused", 0)
        feats, _ = p(input_features=wave)
        assert ops.STATS.get("encoder_clips", 0) - c0 == 3, "only the 3 clips outside the ASR subset may be encoded"
        assert getattr(p, "encoder_clips_reused", 0) - r0 == 2
        assert torch.equal(feats, feats_plain)
        # and through the taps accessor (cache is one-shot: this call encodes everything again)
        for a, b in zip(p.taps(wave), taps_plain):
            assert torch.equal(a, b)
        # log-mel input (the reference's own processor output) is matched the same way
        mel = torch.from_numpy(o_mel.logmel(waves, cfg.n_mels)).cuda()
        feats_mel, _ = p(input_features=mel)
        p.whisper.generate(input_features=mel[[0, 4]].contiguous(), attention_mask=None, max_new_tokens=3)
        c0 = ops.STATS.get("encoder_clips", 0)
        feats_mel2, _ = p(input_features=mel)
        assert ops.STATS.get("encoder_clips", 0) - c0 == 3 and torch.equal(feats_mel2, feats_mel)

"""End-to-end parity of the device path (log-mel -> encoder taps -> Q-Former connector -> splice) against the oracle
and the committed golden vectors produced by the reference's own code.  Tolerances are north_star's:
taps / connector outputs in bf16: max|delta|/max|ref| <= 2e-2 and cosine >= 0.999."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder as o_enc, logmel as o_mel, qformer as o_qf  # noqa: E402
from oracle.configs import CONFIGS  # noqa: E402
from oracle.synth import synth_waves  # noqa: E402
from oracle.weights import make_weights  # noqa: E402

KINDS = ("noise_full", "speechy_12s", "short_3p5s", "silence", "long_31s", "tone_clip")


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def _cos(a, b):
    a, b = a.float().cpu().flatten(), b.float().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm())).item()


def build_perception(cfg, W, device="cuda"):
    """Stand-in for the reference model's containers (HF Whisper module + connector mirror) loaded with oracle weights."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from desta_b200.connector import QformerConnector
    from desta_b200.perception import WhisperPerceptionB200
    wcfg = WhisperConfig(num_mel_bins=cfg.n_mels, d_model=cfg.d_model, encoder_layers=cfg.n_layers,
                         encoder_attention_heads=cfg.n_heads, encoder_ffn_dim=cfg.ffn, decoder_layers=1,
                         decoder_attention_heads=cfg.n_heads, decoder_ffn_dim=64, max_source_positions=cfg.max_pos,
                         vocab_size=64, max_target_positions=8, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                         decoder_start_token_id=1)
    whisper = WhisperForConditionalGeneration(wcfg)
    con = QformerConnector(cfg.d_model, cfg.n_heads, cfg.d_llm, cfg.prompt_size, cfg.qf_layers, len(cfg.tap_layers))
    p = WhisperPerceptionB200(whisper, con, cfg.tap_layers, cfg.prompt_size)
    sd = {k[len("perception."):]: v for k, v in W.items() if k.startswith("perception.")}
    missing, unexpected = p.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not [m for m in missing if ("encoder." in m and "encoder.layer_norm" not in m) or "connector" in m], missing
    return p.to(device).eval()


@pytest.mark.parametrize("name,n_clips", [("micro", 3), ("debug", 2)])
def test_taps_and_connector_vs_oracle(name, n_clips, golden_dir):
    cfg = CONFIGS[name]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)[:n_clips]
    mel = torch.from_numpy(o_mel.logmel(waves, cfg.n_mels))
    with torch.no_grad():
        taps_ref = o_enc.encoder_taps(W, cfg, mel)
        feats_ref = o_qf.connector(W, cfg, taps_ref)
    p = build_perception(cfg, W)
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    taps = [t.clone() for t in p.taps(wave)]
    for k, (t, r) in enumerate(zip(taps, taps_ref)):
        rel, cos = _rel(t, r), _cos(t, r)
        assert rel <= 2e-2 and cos >= 0.999, f"tap {k}: rel {rel:.4f} cos {cos:.6f}"
    with torch.no_grad():
        feats, lens = p(input_features=wave)
    assert lens == [cfg.prompt_size] * n_clips and feats.shape == (n_clips, cfg.prompt_size, cfg.d_llm)
    rel, cos = _rel(feats, feats_ref), _cos(feats, feats_ref)
    assert rel <= 2e-2 and cos >= 0.999, f"connector out: rel {rel:.4f} cos {cos:.6f}"
    # reference-style input (log-mel computed by the caller) takes the same path
    with torch.no_grad():
        feats2, _ = p(input_features=mel)
    assert _rel(feats2, feats_ref) <= 2e-2
    if name == "micro":
        gold = np.load(os.path.join(golden_dir, "micro_path.npz"))
        for k in range(4):
            g = torch.from_numpy(gold[f"tap{k}"])
            assert _rel(taps[k][:, ::25], g) <= 2e-2 and _cos(taps[k][:, ::25], g) >= 0.999
        assert _rel(feats, torch.from_numpy(gold["feats"])) <= 2e-2
    with pytest.raises(ValueError), torch.no_grad():
        p(input_features=mel[..., :2000])

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


@pytest.mark.parametrize("name", ["micro", "debug"])
def test_native_schedule_equals_per_op_schedule(name):
    """desta_encoder_fwd / desta_connector_fwd (one C call per stage) launch the same kernels in the same order as the
    per-op schedule: taps, residual stream and connector output are bit-identical."""
    from desta_b200 import engine
    cfg = CONFIGS[name]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)[:3]
    p = build_perception(cfg, W)
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    enc, con = p._packs()
    ws = p._workspace(3)
    mel_tm, mel_floor = p.mel_time_major(wave)
    assert mel_floor is not None                       # waveform input: the floor is left to the conv1 im2col
    out_native = torch.empty(3 * cfg.prompt_size, cfg.d_llm, device="cuda")
    out_ops = torch.empty_like(out_native)
    engine.encoder_forward(enc, p.dims, ws, mel_tm, mel_floor=mel_floor)
    engine.connector_forward(con, p.dims, ws, out_native)
    taps_native, h_native = ws.taps.clone(), ws.h.clone()
    ws.taps.zero_(); ws.h.zero_()
    engine.encoder_forward_ops(enc, p.dims, ws, mel_tm, mel_floor)
    engine.connector_forward_ops(con, p.dims, ws, out_ops)
    assert torch.equal(ws.taps, taps_native) and torch.equal(ws.h, h_native) and torch.equal(out_ops, out_native)
    assert out_native.abs().max().item() > 0
    # argument validation of the stage entry points
    import ctypes
    from desta_b200._lib import DestaError, lib
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    rc = lib().desta_encoder_fwd(ctypes.byref(ws._c_dims), ctypes.byref(enc.c_weights), ctypes.c_void_p(mel_tm.data_ptr()),
                                 ctypes.c_void_p(mel_floor.data_ptr()), 3, ctypes.c_void_p(small.data_ptr()), 16, ctypes.c_void_p(ws.taps.data_ptr()),
                                 ctypes.c_void_p(ws.h.data_ptr()), None)
    assert rc == -1 and b"workspace too small" in lib().desta_last_error()


@pytest.mark.parametrize("name", ["micro", "debug"])
def test_cuda_graph_replay_equals_eager(name):
    """capture_graph(): the whole waveform -> audio-feature path replayed as one CUDA graph gives bit-identical features,
    keeps doing so for new inputs, and steps aside for other batch sizes, train mode and updated connector weights."""
    cfg = CONFIGS[name]
    W = make_weights(cfg, seed=11, with_table=False)
    waves = synth_waves(seed=7, kinds=KINDS)
    p = build_perception(cfg, W)
    wave_a = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves[:3]])).cuda()
    wave_b = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves[3:6]])).cuda()
    with torch.no_grad():
        eager_a = p.forward_features(wave_a).clone()
        eager_b = p.forward_features(wave_b).clone()
        eager_2 = p.forward_features(wave_a[:2]).clone()
    p.capture_graph(3)
    from desta_b200 import ops
    l0 = ops.STATS["launches"]
    with torch.no_grad():
        got_a = p.forward_features(wave_a)
        got_b = p.forward_features(wave_b)
        got_a2, _ = p(input_features=wave_a)
    assert ops.STATS["launches"] == l0, "replays must not launch kernels one by one"
    assert torch.equal(got_a, eager_a), "graph replay differs from eager (first input)"
    assert torch.equal(got_b, eager_b), "graph replay differs from eager (second input)"
    assert torch.equal(got_a2, eager_a), "module call did not replay"
    with torch.no_grad():
        assert torch.equal(p.forward_features(wave_a[:2]), eager_2), "other batch size must take the eager path"
    assert ops.STATS["launches"] > l0
    # the eager call replaced the module's resident workspace; the graph must still own the one it was captured with
    junk = [torch.full((1 << 20,), float(i), device="cuda") for i in range(64)]
    with torch.no_grad():
        assert torch.equal(p.forward_features(wave_b), eager_b), "graph buffers were recycled under the graph"
    del junk
    with torch.no_grad():                                                  # weights change -> the stale graph is not used
        p.connector.layer_weights[:, 0] += 1.5                             # one tap up (a uniform shift cancels in the softmax)
        l1 = ops.STATS["launches"]
        changed = p.forward_features(wave_a)
    assert ops.STATS["launches"] > l1, "stale graph replayed after a weight update"
    assert not torch.equal(changed, eager_a)

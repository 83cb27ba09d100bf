"""CPU suite: the oracle against the committed golden vectors (generated from the reference's own code by
oracle/make_golden.py), host integer logic, and the C-ABI surface of libdesta_b200.so."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import encoder as o_enc, logmel as o_mel, qformer as o_qf, splice as o_sp
from oracle.configs import CONFIGS
from oracle.synth import synth_waves
from oracle.weights import make_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KINDS = ("noise_full", "speechy_12s", "short_3p5s", "silence", "long_31s", "tone_clip")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "micro_path.npz"))


def test_start_positions_kat(gold):
    res, starts = o_sp.prepare_audio_context_and_start_positions(list(gold["kat_tokens"]), 999, [3, 3], [2, 0], 777)
    assert res == list(gold["kat_result"]) and starts == list(gold["kat_starts"]) == [2, 9] and len(res) == 13
    assert o_sp.left_pad_offset(starts, [0, 0, 0, 1, 1, 1, 1]) == [5, 12]


def test_start_positions_consumes_lists():
    a, t = [64, 64, 64], [1, 2, 3]
    o_sp.prepare_audio_context_and_start_positions([5, 9, 5], 9, a, t, 0)
    assert a == [64, 64] and t == [2, 3]     # pop(0) across calls, like the reference


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_oracle_vs_golden(n_mels, golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    mel = o_mel.logmel(synth_waves(seed=7, kinds=KINDS), n_mels)
    assert mel.shape == (6, n_mels, 3000) and mel.dtype == np.float32
    assert np.abs(mel[:, ::7, ::5] - g[f"mel{n_mels}_strided"]).max() < 1e-4
    assert np.abs(mel[:, :, :32] - g[f"mel{n_mels}_head"]).max() < 1e-4
    assert np.abs(mel[:, :, -32:] - g[f"mel{n_mels}_tail"]).max() < 1e-4
    assert np.abs(mel.astype(np.float64).sum(axis=(1, 2)) - g[f"mel{n_mels}_sum"]).max() < 0.5
    assert np.all(mel[3] == -1.5)


def test_path_oracle_vs_golden(gold):
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11)
    assert abs(float(sum(v.double().sum() for v in W.values())) - float(gold["weight_checksum"][0])) < 1e-6
    mel = torch.from_numpy(o_mel.logmel(synth_waves(seed=7, kinds=KINDS)[:3], cfg.n_mels))
    with torch.no_grad():
        taps = o_enc.encoder_taps(W, cfg, mel)
        feats = o_qf.connector(W, cfg, taps)
    for i, t in enumerate(taps):
        assert (t[:, ::25] - torch.from_numpy(gold[f"tap{i}"])).abs().max().item() < 1e-4
    assert (feats - torch.from_numpy(gold["feats"])).abs().max().item() < 1e-4
    with pytest.raises(ValueError):
        o_enc.encoder_taps(W, cfg, mel[..., :2000])
    # backward oracle (autograd through the restatement) vs the reference's autograd
    Wg = {k: (v.clone().requires_grad_(True) if k.startswith("perception.connector.") else v) for k, v in W.items()}
    out = o_qf.connector(Wg, cfg, [t.detach() for t in taps])
    (out * torch.from_numpy(gold["gout"])).sum().backward()
    names = list(gold["grad_names"])
    norms = np.array([Wg[k].grad.double().norm().item() for k in names])
    assert np.abs(norms - gold["grad_norms"]).max() / gold["grad_norms"].max() < 1e-4
    assert (Wg["perception.connector.layer_weights"].grad - torch.from_numpy(gold["grad_layer_weights"])).abs().max() < 1e-4
    gq = Wg["perception.connector.qformer.layer.0.crossattention.self.query.weight"].grad[::8, ::8]
    assert (gq - torch.from_numpy(gold["grad_q0"])).abs().max().item() < 1e-4 * max(1.0, np.abs(gold["grad_q0"]).max())


def test_train_mode_dropout_sites_vs_golden(gold):
    """The reference's train-mode output (BertConfig dropout 0.1, torch.manual_seed(123)) is reproduced by replaying
    torch's RNG stream through the restatement's five dropout sites per layer: pins their placement and order."""
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11)
    mel = torch.from_numpy(o_mel.logmel(synth_waves(seed=7, kinds=KINDS)[:3], cfg.n_mels))
    with torch.no_grad():
        taps = o_enc.encoder_taps(W, cfg, mel)
        torch.manual_seed(123)
        feats = o_qf.connector(W, cfg, taps, drop=o_qf.torch_dropout(0.1, 0.1))
    ref = torch.from_numpy(gold["feats_train_seed123"])
    assert (feats - ref).abs().max().item() < 1e-4
    assert (ref - torch.from_numpy(gold["feats"])).abs().max().item() > 1e-2      # dropout really was active


def test_splice_oracle_vs_golden(gold):
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11)
    bits = W["llm_model.model.embed_tokens.weight"].view(torch.int16).numpy().view(np.uint16)
    trans = [gold["trans0"].reshape(-1), gold["trans1"].reshape(-1), gold["trans2"].reshape(-1)]
    out = o_sp.splice(gold["input_ids"], bits, gold["feats"], trans, [tuple(p) for p in gold["positions"]])
    assert np.array_equal(out, gold["embeds_bits"])


def test_bf16_rounding_matches_torch():
    x = torch.randn(10000) * 100
    x[:6] = torch.tensor([float("inf"), -0.0, 1e-40, 3.3895313892515355e38, 1.00390625, 1.01171875])
    want = x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(o_sp.f32_to_bf16_bits(x.numpy()), want)


def test_c_abi_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from desta_b200 import _lib
    header = open(os.path.join(ROOT, "include", "desta_b200.h")).read()
    declared = set(re.findall(r"\b(desta_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/desta_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert _lib.lib().desta_version() >= 100
    assert _lib.lib().desta_logmel_plan_bytes(128) > 0 and _lib.lib().desta_logmel_plan_bytes(0) == 0


def test_host_mel_plan_matches_oracle_filters():
    from desta_b200 import _lib
    L = _lib.lib()
    for n_mels in (80, 128):
        n = L.desta_logmel_plan_bytes(n_mels)
        h = np.zeros(n // 4, dtype=np.int32)
        assert L.desta_logmel_plan_build(n_mels, h.ctypes.data_as(ctypes.c_void_p), n) == 0
        # plan layout (csrc/logmel.cu): header[32] | hann/2 [400] | tw1 [20][16] c64 | tw2 [11][16] c64 | meta [nm] int2 | taps float4
        nm, total4 = int(h[0]), int(h[1])
        off = 32 + 400 + 640 + 352
        meta = h[off:off + 2 * nm].reshape(nm, 2)
        lo, n4, woff = meta[:, 0] & 0xFFFF, meta[:, 0] >> 16, meta[:, 1]
        w = h[off + 2 * nm:].view(np.float32)
        assert w.size == 4 * total4 == 4 * int(n4.sum())
        rec = np.zeros((201 + 4, nm), np.float32)
        for m in range(nm):
            rec[lo[m]:lo[m] + 4 * n4[m], m] = w[4 * woff[m]:4 * (woff[m] + n4[m])]      # taps zero-padded to groups of four
        assert np.all(rec[201:] == 0)
        assert np.array_equal(rec[:201], o_mel.mel_filter_bank(n_mels).astype(np.float32))
        hann = h[32:432].view(np.float32)
        assert np.abs(2.0 * hann - torch.hann_window(400).numpy()).max() < 1e-6      # halved: the real-FFT untangling drops its 1/2
        split = h[3:14]                                                              # mel range of each of the 10 warps
        assert split[0] == 0 and split[10] == nm and np.all(np.diff(split) >= 0) and np.all(split % 2 == 0)
        tw1 = h[32 + 400:32 + 1040].view(np.float32).reshape(20, 16, 2)
        k1, t = 7, 9
        assert abs(tw1[k1, t, 0] - np.cos(2 * np.pi * k1 * t / 200)) < 1e-6 and abs(tw1[k1, t, 1] + np.sin(2 * np.pi * k1 * t / 200)) < 1e-6

"""Generate tests/golden/*.npz from the REFERENCE's own code (run in the build container only).
TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

For each piece of the path it runs the reference (modeling_desta25.py through the shims of oracle/refshim.py,
driving transformers' WhisperFeatureExtractor / WhisperEncoderLayer / BertEncoder), checks the oracle
restatement against it, and stores the reference outputs (sub-sampled where large) as fixtures."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import encoder as o_enc, logmel as o_mel, qformer as o_qf, splice as o_sp  # noqa: E402
from oracle.configs import CONFIGS  # noqa: E402
from oracle.refshim import load_reference  # noqa: E402
from oracle.weights import make_weights  # noqa: E402
from oracle.synth import synth_waves  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    cfg = CONFIGS["micro"]
    ref = load_reference(cfg)
    os.makedirs(GOLD, exist_ok=True)

    # ---------------- start positions (a8) ----------------
    toks = [1, 10, 999, 11, 12, 999, 13]
    res, starts = ref._prepare_audio_context_and_start_positions(toks, 999, [3, 3], [2, 0], 777)
    ores, ostarts = o_sp.prepare_audio_context_and_start_positions(toks, 999, [3, 3], [2, 0], 777)
    assert (res, starts) == (ores, ostarts) and starts == [2, 9] and len(res) == 13
    kat = {"tokens": np.array(toks), "result": np.array(res), "starts": np.array(starts)}

    # ---------------- log-mel (a1) ----------------
    from transformers import WhisperFeatureExtractor
    waves = synth_waves(seed=7, kinds=("noise_full", "speechy_12s", "short_3p5s", "silence", "long_31s", "tone_clip"))
    mel_fix = {}
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels)
        ref_mel = fe(waves, sampling_rate=16000, return_tensors="pt").input_features.numpy()
        omel = o_mel.logmel(waves, n_mels)
        err = np.abs(ref_mel - omel).max()
        print(f"logmel n_mels={n_mels}: oracle vs HF max abs err {err:.3e}")
        assert err < 1e-4, err
        ferr = np.abs(fe.mel_filters - o_mel.mel_filter_bank(n_mels)).max()
        assert ferr < 1e-12, ferr
        # store reference output: all frames of a strided mel subset + full first/last 32 frames
        mel_fix[f"mel{n_mels}_strided"] = ref_mel[:, ::7, ::5].copy()
        mel_fix[f"mel{n_mels}_head"] = ref_mel[:, :, :32].copy()
        mel_fix[f"mel{n_mels}_tail"] = ref_mel[:, :, -32:].copy()
        mel_fix[f"mel{n_mels}_sum"] = ref_mel.astype(np.float64).sum(axis=(1, 2))
    np.savez_compressed(os.path.join(GOLD, "logmel.npz"), **mel_fix)

    # ---------------- model (a2-a7) ----------------
    rcfg = ref.DeSTA25Config(llm_model_id="llama-micro", encoder_model_id=cfg.encoder_model_id,
                             qformer_num_hidden_layers=cfg.qf_layers, prompt_size=cfg.prompt_size)
    rcfg.target_layer_ids = None
    model = ref.DeSTA25AudioModel(rcfg).eval()
    W = make_weights(cfg, seed=11)
    sd = model.state_dict(keep_vars=True) if False else dict(model.named_parameters())
    missing = [k for k in W if k not in sd]
    assert not missing, missing[:5]
    with torch.no_grad():
        for k, v in W.items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
            sd[k].copy_(v.to(sd[k].dtype))
    assert model.perception.connector.config.target_layer_ids == cfg.tap_layers

    N = 3
    mel = torch.from_numpy(o_mel.logmel(waves[:N], cfg.n_mels))
    taps_ref = []
    hooks = [layer.register_forward_hook(lambda m, i, o: taps_ref.append(o[0] if isinstance(o, tuple) else o))
             for layer in model.perception.whisper.model.encoder.layers]
    feats_ref, lens = model.perception(input_features=mel)
    for h in hooks:
        h.remove()
    assert lens == [cfg.prompt_size] * N
    taps_o = o_enc.encoder_taps(W, cfg, mel)
    feats_o = o_qf.connector(W, cfg, taps_o)
    for a, b in zip(taps_ref, taps_o):
        e = (a - b).abs().max().item()
        print(f"tap: oracle vs reference max abs err {e:.3e} (max |ref| {a.abs().max():.3f})")
        assert e < 2e-4, e
    e = (feats_ref - feats_o).abs().max().item()
    print(f"connector out: oracle vs reference max abs err {e:.3e} (max |ref| {feats_ref.abs().max():.3f})")
    assert e < 2e-4, e

    # connector backward oracle vs reference autograd (dropout off: eval mode)
    torch.set_grad_enabled(True)
    gout = torch.randn(feats_ref.shape, generator=torch.Generator().manual_seed(5))
    model.zero_grad()
    f2, _ = model.perception(input_features=mel)
    (f2 * gout).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    Wg = {k: (v.clone().requires_grad_(True) if k.startswith("perception.connector.") else v) for k, v in W.items()}
    fo = o_qf.connector(Wg, cfg, [t.detach() for t in taps_o])
    (fo * gout).sum().backward()
    worst = 0.0
    for k, g in ref_grads.items():
        go = Wg[k].grad
        rel = (g - go).abs().max().item() / (g.abs().max().item() + 1e-12)
        worst = max(worst, rel)
    print(f"connector grads ({len(ref_grads)} tensors): worst max-rel err {worst:.3e}")
    assert worst < 1e-3 and len(ref_grads) == sum(k.startswith("perception.connector.") for k in W)
    torch.set_grad_enabled(False)

    # train mode: the Q-Former's five dropout sites per layer (BertConfig p = 0.1).  Replaying torch's RNG stream through
    # the restatement pins the site order / placement against the reference; the reference's output is stored.
    model.train()
    torch.manual_seed(123)
    feats_train_ref, _ = model.perception(input_features=mel)
    model.eval()
    torch.manual_seed(123)
    feats_train_o = o_qf.connector(W, cfg, taps_o, drop=o_qf.torch_dropout(0.1, 0.1))
    e = (feats_train_ref - feats_train_o).abs().max().item()
    print(f"connector out, train mode (dropout 0.1, seed 123): oracle vs reference max abs err {e:.3e}; "
          f"train vs eval differ by {(feats_train_ref - feats_ref).abs().max().item():.3f}")
    assert e < 2e-4, e

    # splice (a7): 2 prompts, 3 clips, transcript lengths 5/0/2
    B, S = 2, 160
    g = torch.Generator().manual_seed(3)
    input_ids = torch.randint(0, cfg.vocab, (B, S), generator=g)
    trans = [torch.randint(0, cfg.vocab, (1, t), generator=g) for t in (5, 0, 2)]
    positions = [(0, 4), (1, 7), (1, 7 + 64 + 3)]
    attn = torch.ones(B, S, dtype=torch.long)
    emb_ref = model._prepare_inputs_for_llm(input_ids, attn, mel, trans, positions)
    table_bits = bits(W["llm_model.model.embed_tokens.weight"])
    emb_o = o_sp.splice(input_ids.numpy(), table_bits, feats_ref.float().numpy(),
                        [t.numpy().reshape(-1) for t in trans], positions)
    assert emb_ref.dtype == torch.bfloat16
    assert np.array_equal(bits(emb_ref), emb_o), "splice oracle is not bit-exact vs reference"
    print("splice: oracle bit-exact vs reference")

    np.savez_compressed(
        os.path.join(GOLD, "micro_path.npz"),
        kat_tokens=kat["tokens"], kat_result=kat["result"], kat_starts=kat["starts"],
        **{f'tap{i}': taps_ref[i][:, ::25, :].numpy() for i in range(4)},
        feats=feats_ref.numpy(), gout=gout.numpy(), feats_train_seed123=feats_train_ref.numpy(),
        grad_names=np.array(sorted(ref_grads)),
        grad_norms=np.array([ref_grads[k].double().norm().item() for k in sorted(ref_grads)]),
        grad_layer_weights=ref_grads["perception.connector.layer_weights"].numpy(),
        grad_proj_bias=ref_grads["perception.connector.proj.1.bias"].numpy(),
        grad_q0=ref_grads["perception.connector.qformer.layer.0.crossattention.self.query.weight"][::8, ::8].numpy(),
        grad_k1=ref_grads["perception.connector.qformer.layer.1.crossattention.self.key.weight"][::8, ::8].numpy(),
        input_ids=input_ids.numpy(), trans0=trans[0].numpy(), trans1=trans[1].numpy(), trans2=trans[2].numpy(),
        positions=np.array(positions), embeds_bits=bits(emb_ref),
        weight_checksum=np.array([float(sum(v.double().sum() for v in W.values()))]),
    )
    print("wrote", GOLD)


if __name__ == "__main__":
    main()

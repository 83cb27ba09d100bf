"""Drop-in boundary on the GPU: `desta_b200.install(model)` on a stand-in with the reference model's attribute layout
(`.llm_model.model.embed_tokens`, `.perception.{whisper,connector,config}`, `.processor`, `.config.connector_mode`),
prompt assembly from token lists (host integers) and `_prepare_inputs_for_llm` against the oracle, Qwen3-style
multi-clip prompts (BASELINE configs[3])."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder as o_enc, logmel as o_mel, qformer as o_qf, splice as o_sp  # noqa: E402
from oracle.configs import CONFIGS  # noqa: E402
from oracle.synth import synth_waves  # noqa: E402
from oracle.weights import make_weights  # noqa: E402
from test_path_gpu import KINDS, build_perception  # noqa: E402


def _bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def _fake_reference_model(cfg, W):
    """Attribute layout of DeSTA25AudioModel (ref: modeling_desta25.py:698-757) without the LLM decoder."""
    ours = build_perception(cfg, W)                 # supplies HF whisper module + connector container with the weights
    ref_perception = types.SimpleNamespace(
        whisper=ours.whisper, connector=ours.connector,
        config=types.SimpleNamespace(prompt_size=cfg.prompt_size))
    ref_perception.connector.config = types.SimpleNamespace(target_layer_ids=list(cfg.tap_layers))
    embed = torch.nn.Embedding(cfg.vocab, cfg.d_llm).to(torch.bfloat16).cuda()
    with torch.no_grad():
        embed.weight.copy_(W["llm_model.model.embed_tokens.weight"])
    model = types.SimpleNamespace(
        config=types.SimpleNamespace(connector_mode="qformer_1", prompt_size=cfg.prompt_size),
        perception=ref_perception,
        llm_model=types.SimpleNamespace(model=types.SimpleNamespace(embed_tokens=embed)),
        processor=types.SimpleNamespace(batch_decode=lambda *a, **k: ["decoded"]))
    return model


def test_install_and_multi_clip_prompts_vs_oracle():
    import desta_b200
    from desta_b200.processor import WaveformProcessor
    from desta_b200.splice import apply_left_pad_offset, prepare_audio_context_and_start_positions
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11)
    model = desta_b200.install(_fake_reference_model(cfg, W))
    assert isinstance(model.processor, WaveformProcessor) and model.processor.batch_decode() == ["decoded"]
    with pytest.raises(ValueError):
        model.processor([np.zeros(100, np.float32)], sampling_rate=8000)

    # ---- prompt assembly (host integers), 2 prompts: 3 clips + 1 clip, transcripts 0 / 3 / 17 / 5 tokens ----
    LOC, PH, PAD = 500, 501, 0
    rng = np.random.default_rng(0)
    waves = synth_waves(seed=7, kinds=KINDS)[:4]
    trans = [rng.integers(1, 400, n) for n in (0, 3, 17, 5)]
    prompts = [[1, 7, LOC, 8, 9, LOC, 10, LOC, 11], [1, LOC, 12, 13]]
    clip_of_prompt = [[0, 1, 2], [3]]
    seqs, starts = [], []
    for toks, clips in zip(prompts, clip_of_prompt):
        a_sizes, t_sizes = [cfg.prompt_size] * len(clips), [len(trans[c]) for c in clips]
        a2, t2 = list(a_sizes), list(t_sizes)
        res, st = prepare_audio_context_and_start_positions(toks, LOC, a_sizes, t_sizes, PH)
        ores, ost = o_sp.prepare_audio_context_and_start_positions(toks, LOC, a2, t2, PH)
        assert (res, st) == (ores, ost) and a_sizes == [] and t_sizes == []
        seqs.append(res)
        starts.append(st)
    S = max(len(s) for s in seqs)
    ids = np.full((2, S), PAD, dtype=np.int64)
    mask = np.zeros((2, S), dtype=np.int64)
    positions = []
    for b, (s, st) in enumerate(zip(seqs, starts)):
        ids[b, S - len(s):] = s                       # left padding (tokenizer padding_side="left")
        mask[b, S - len(s):] = 1
        st = apply_left_pad_offset(st, mask[b])
        assert st == o_sp.left_pad_offset(starts[b], mask[b])
        positions += [(b, p) for p in st]

    # ---- the public call of the reference, now on libdesta_b200.so ----
    proc = model.processor(waves, sampling_rate=16000, return_tensors="pt")
    feats_in = proc.input_features.cuda()
    assert feats_in.shape == (4, 480000)
    trans_t = [torch.from_numpy(t).view(1, -1) for t in trans]
    with torch.no_grad():
        emb = model._prepare_inputs_for_llm(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda(), feats_in, trans_t, positions)
        feats, _ = model.perception(input_features=feats_in)
    assert emb.dtype == torch.bfloat16 and emb.shape == (2, S, cfg.d_llm)
    # exact given our own connector output: text / transcript rows bit-exact, audio rows = RNE(bf16) of the features
    table_bits = _bits(W["llm_model.model.embed_tokens.weight"])
    want = o_sp.splice(ids, table_bits, feats.float().cpu().numpy(), trans, positions)
    assert np.array_equal(_bits(emb), want)
    # and the features themselves against the oracle (bf16 tolerance)
    mel = torch.from_numpy(o_mel.logmel(waves, cfg.n_mels))
    with torch.no_grad():
        ref = o_qf.connector(W, cfg, o_enc.encoder_taps(W, cfg, mel))
    rel = ((feats.float().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert rel <= 2e-2
    # N == 0 fast path and length mismatch
    with torch.no_grad():
        plain = model._prepare_inputs_for_llm(torch.from_numpy(ids).cuda(), None, feats_in[:0], [], [])
    assert np.array_equal(_bits(plain), table_bits[ids])
    with pytest.raises(AssertionError), torch.no_grad():
        model._prepare_inputs_for_llm(torch.from_numpy(ids).cuda(), None, feats_in, trans_t[:3], positions)


def test_gradient_reaches_connector_through_the_splice():
    import desta_b200
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11)
    model = desta_b200.install(_fake_reference_model(cfg, W))
    model.perception.train()
    for p in model.perception.whisper.parameters():
        p.requires_grad_(False)
    waves = synth_waves(seed=7, kinds=KINDS)[:2]
    wave = torch.from_numpy(np.stack([o_mel.pad_or_truncate(w) for w in waves])).cuda()
    ids = torch.randint(0, cfg.vocab, (2, 100)).cuda()
    emb = model._prepare_inputs_for_llm(ids, None, wave, [torch.zeros(1, 0, dtype=torch.long)] * 2, [(0, 3), (1, 30)])
    assert emb.requires_grad
    g = torch.zeros_like(emb)
    g[0, 3:3 + 64] = 1.0            # only clip 0's audio rows receive gradient
    emb.backward(g)
    gw = model.perception.connector.proj[1].bias.grad
    assert gw is not None and torch.allclose(gw, torch.full_like(gw, 64.0), rtol=2e-2)


def test_stager_double_buffer_feeds_the_path():
    """Scope row n1: batches staged through pinned slots + copy stream give the same features as direct transfers."""
    from desta_b200.staging import WaveformStager, stage_batches
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11, with_table=False)
    p = build_perception(cfg, W)
    waves = synth_waves(seed=7, kinds=KINDS)
    batches = [waves[0:2], [w.tolist() for w in waves[2:4]], waves[4:6], waves[1:3]]
    st = WaveformStager(2, "cuda", slots=2)
    got = []
    with torch.no_grad():
        for wave, lens in stage_batches(st, batches):
            assert wave.is_cuda and wave.shape == (2, 480000) and lens.dtype == torch.int32
            feats, _ = p(input_features=wave)
            got.append(feats.clone())
        for k, b in enumerate(batches):
            direct = torch.from_numpy(np.stack([o_mel.pad_or_truncate(np.asarray(w, np.float32)) for w in b])).cuda()
            want, _ = p(input_features=direct)
            assert torch.equal(got[k], want), k
    assert st.h2d_bytes == 4 * 2 * (480000 * 4 + 4)


def test_processor_prefetch_hands_over_the_device_copy():
    """Scope row n1 behind the reference's own call: `processor(clips).input_features.to(device)` with
    enable_pinned_staging(device=...) returns the device buffer whose H2D copy started at packing time (loader thread,
    copy stream); features are bit-identical to a direct transfer over more batches than slots (device buffers are reused),
    a host tensor modified in place takes the regular copy, and views / other targets behave like plain tensors."""
    import threading
    from desta_b200.processor import StagedWaveform, WaveformProcessor
    cfg = CONFIGS["micro"]
    W = make_weights(cfg, seed=11, with_table=False)
    p = build_perception(cfg, W)
    waves = synth_waves(seed=7, kinds=KINDS)
    batches = [waves[0:2], waves[2:4], waves[4:6], waves[1:3], waves[3:5], [waves[5], waves[0]], waves[2:4]]
    proc = WaveformProcessor()
    proc.enable_pinned_staging(2, slots=3, device="cuda")
    got, nxt = [], {}

    def pack(k):
        nxt["f"] = proc(list(batches[k]), sampling_rate=16000, return_tensors="pt").input_features

    pack(0)
    with torch.no_grad():
        for k in range(len(batches)):
            feats = nxt["f"]
            loader = threading.Thread(target=pack, args=(k + 1,)) if k + 1 < len(batches) else None
            if loader:
                loader.start()                                   # next batch packed + prefetched while this one computes
            assert isinstance(feats, StagedWaveform) and feats.is_pinned() and not feats.is_cuda
            wave = feats.to("cuda", non_blocking=True)
            assert wave.is_cuda and wave.shape == (2, 480000) and wave.data_ptr() in [s["dev"].data_ptr() for s in proc._slots]
            out, _ = p(input_features=wave)
            got.append(out.clone())
            if loader:
                loader.join()
        for k, b in enumerate(batches):
            direct = torch.from_numpy(np.stack([o_mel.pad_or_truncate(np.asarray(w, np.float32)) for w in b])).cuda()
            want, _ = p(input_features=direct)
            assert torch.equal(got[k], want), k
    # plain-tensor behaviour everywhere else
    f = proc(list(batches[0]), sampling_rate=16000, return_tensors="pt").input_features
    assert f.to(torch.float64).dtype == torch.float64 and not f.to(torch.float64).is_cuda
    assert f[:1].to("cuda").data_ptr() not in [s["dev"].data_ptr() for s in proc._slots]       # a view is copied normally
    f.mul_(0.5)                                                  # modified in place after staging: the prefetched copy is stale
    d = f.to("cuda")
    assert d.data_ptr() not in [s["dev"].data_ptr() for s in proc._slots]
    assert torch.equal(d.cpu(), f.as_subclass(torch.Tensor))

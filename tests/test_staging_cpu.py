"""Host logic of the collate-side staging (scope row n1): pad / truncate into a reusable slot equals the oracle's
restatement of WhisperFeatureExtractor's padding ($TF/models/whisper/feature_extraction_whisper.py:296-320)."""
import numpy as np
import pytest
import torch

from oracle import logmel as o_mel
from oracle.synth import synth_waves

KINDS = ("noise_full", "speechy_12s", "short_3p5s", "silence", "long_31s", "tone_clip")


def test_pad_or_truncate_in_place_matches_oracle_and_clears_stale_data():
    from desta_b200.processor import N_SAMPLES, pad_or_truncate_batch
    waves = synth_waves(seed=7, kinds=KINDS)
    want = np.stack([o_mel.pad_or_truncate(w) for w in waves])
    got, lens = pad_or_truncate_batch(waves)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert list(lens) == [min(len(w), N_SAMPLES) for w in waves]
    slot = np.full((8, N_SAMPLES), 7.0, dtype=np.float32)            # a reused slot full of stale samples
    ln = np.full(8, -1, dtype=np.int32)
    got2, lens2 = pad_or_truncate_batch([w.tolist() if i == 2 else w for i, w in enumerate(waves)], out=slot, lengths=ln)
    assert got2.base is slot and np.array_equal(got2, want) and np.array_equal(lens2, lens)
    assert np.all(slot[6:] == 7.0)                                   # rows beyond the batch are untouched
    empty, el = pad_or_truncate_batch([])
    assert empty.shape == (0, N_SAMPLES) and el.shape == (0,)
    with pytest.raises(ValueError):
        pad_or_truncate_batch([np.zeros((2, 100), np.float32)])


def test_processor_shim_contract():
    from desta_b200.processor import WaveformProcessor
    waves = synth_waves(seed=7, kinds=KINDS)[:2]
    proc = WaveformProcessor()
    feats = proc(waves, sampling_rate=16000, return_tensors="pt").input_features
    assert feats.shape == (2, 480000) and feats.dtype == torch.float32
    assert np.array_equal(feats[1].numpy(), o_mel.pad_or_truncate(waves[1]))
    with pytest.raises(ValueError, match="sampling rate"):            # same error as WhisperFeatureExtractor.__call__
        proc(waves, sampling_rate=8000)
    with pytest.raises(AttributeError):
        proc.batch_decode                                              # nothing wrapped: delegation has no target


def test_stager_host_side_without_cuda():
    from desta_b200.staging import WaveformStager
    waves = synth_waves(seed=7, kinds=KINDS)
    st = WaveformStager(4, "cpu", slots=2, pin=False)
    assert st.fill(0, waves[:3]) == 3 and st.fill(1, waves[3:6]) == 3
    assert np.array_equal(st.host[0][:3].numpy(), np.stack([o_mel.pad_or_truncate(w) for w in waves[:3]]))
    assert st.host_len[1][:3].tolist() == [min(len(w), 480000) for w in waves[3:6]]
    with pytest.raises(ValueError):
        st.fill(0, waves[:5])


def test_pcm_stager_host_side_without_cuda(tmp_path):
    """PcmStager.fill only touches the host: header walk + payload straight into the slot, 16-byte aligned clip starts."""
    import wavutil as W
    from desta_b200.staging import PcmStager
    rng = np.random.default_rng(0)
    paths, payloads = [], []
    for i, (kind, ch, sr, n) in enumerate([(W.S16, 1, 16000, 1001), (W.S24, 2, 44100, 333), (W.F32, 1, 22050, 10)]):
        x = np.clip(0.3 * rng.standard_normal((n, ch)), -1, 0.999)
        data = W.wav_bytes(x, sr, kind, extra_chunks=(i == 1))
        p = tmp_path / f"h{i}.wav"
        p.write_bytes(data)
        paths.append(str(p))
        payloads.append(W.encode(x, kind))
    st = PcmStager(4, "cpu", slot_bytes=1 << 16, pin=False)
    assert st.fill(0, paths) == 3
    for (fmt, ch, sr, off, frames), raw in zip(st.clips[0], payloads):
        assert off % 16 == 0 and bytes(st._host_np[0][off:off + len(raw)]) == raw and frames * ch * (1, 2, 3, 4, 4, 8)[fmt] == len(raw)
    assert [c[2] for c in st.clips[0]] == [16000, 44100, 22050]
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFFxxxxWAVE")
    with pytest.raises(Exception, match="could not be decoded"):
        st.fill(1, [str(bad)])


def test_device_wave_is_only_trusted_by_identity():
    """The device copy rides on the exact object AudioSegment.samples returns; anything derived from it is plain host data
    again (the processor then uploads the edited samples instead of silently reusing the stale device tensor)."""
    from desta_b200.audio import DeviceWave
    from desta_b200.processor import _device_batch
    marker = object()
    w = DeviceWave(np.linspace(-1, 1, 1000, dtype=np.float32), marker)
    assert w.device_tensor is marker
    for derived in (w * 0.5, w[::-1], np.clip(w, -0.1, 0.1), w[:500], w.copy(), w + 0, np.asarray(w) * 1.0, -w):
        assert getattr(derived, "device_tensor", None) is None
        assert _device_batch([derived]) is None
    with np.testing.assert_raises(ValueError):
        w *= 0.5                                                     # the carrier is read-only: no silent drift
    assert w.tolist()[0] == -1.0 and w.device_tensor is marker


def test_staged_waveform_is_a_plain_tensor_without_a_device_slot():
    """processor.StagedWaveform (the pinned batch a prefetching processor hands out) must behave like the host tensor it
    wraps wherever the hand-over does not apply: dtype / cpu targets, views, arithmetic, numpy interop - and `.to()` of a
    wrapper whose slot has no device buffer is the regular copy (the CUDA hand-over itself is covered by
    tests/test_dropin_gpu.py::test_processor_prefetch_hands_over_the_device_copy)."""
    import numpy as np
    import torch
    from desta_b200.processor import StagedWaveform
    host = torch.arange(24, dtype=torch.float32).reshape(3, 8)
    t = StagedWaveform.wrap(host, dict(dev=None), None)
    assert isinstance(t, torch.Tensor) and t.shape == (3, 8) and t.data_ptr() == host.data_ptr()
    d = t.to(torch.float64)
    assert type(d) is torch.Tensor and d.dtype == torch.float64 and torch.equal(d, host.double())
    c = t.to("cpu")
    assert type(c) is torch.Tensor and torch.equal(c, host)
    v = t[:2]
    assert getattr(v, "_slot", None) is None and torch.equal(v.to("cpu"), host[:2])       # views carry no hand-over
    assert torch.equal((t * 2).as_subclass(torch.Tensor), host * 2)
    assert np.array_equal(t.numpy(), host.numpy())
    before = t._staged_version
    t.add_(1.0)                                       # in-place edit after staging: the version counter moves
    assert t._version != before

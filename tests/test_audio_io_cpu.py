"""CPU checks of the decode + resample row (n3): the oracle against what its specification implies (the resampler of the
reference, soxr "HQ" via librosa, is absent from this image: see oracle/audio_io.py), the decode rules against stdlib /
scipy readers of the same bytes, and the host-side pieces of the product (header walk, filter table, error behaviour)."""
import ctypes
import io
import math
import os
import wave

import numpy as np
import pytest
from scipy import signal
from scipy.io import wavfile

import __graft_entry__
from oracle import audio_io as A
import wavutil as W

RATES = (8000, 11025, 22050, 24000, 32000, 44100, 48000)


@pytest.mark.parametrize("sr", RATES)
def test_oracle_resampler_meets_its_specification(sr):
    n = sr // 4
    t = np.arange(n) / sr
    low = min(sr, 16000) / 2
    f = 0.9 * low                                  # inside the pass band (edge 0.913)
    x = np.sin(2 * np.pi * f * t) + 0.5 * np.sin(2 * np.pi * 440 * t + 0.3)
    y = A.resample(x, sr, 16000)
    assert y.dtype == np.float32 and len(y) == math.ceil(n * 16000 / sr)      # librosa's fix_length
    to = np.arange(len(y)) / 16000
    ref = np.sin(2 * np.pi * f * to) + 0.5 * np.sin(2 * np.pi * 440 * to + 0.3)
    edge = int(A.n_taps(sr, 16000) * 16000 / sr) + 2
    assert np.abs(y - ref)[edge:-edge].max() < 2e-6              # band-limited input -> the continuous signal re-sampled
    if 1.02 * low < sr / 2:                                      # first stop-band frequency: > 110 dB down
        ys = A.resample(np.sin(2 * np.pi * 1.02 * low * t), sr, 16000)
        assert np.abs(ys)[edge:-edge].max() < 3e-6
    g = math.gcd(sr, 16000)
    L, M = 16000 // g, sr // g
    half = (A.n_taps(sr, 16000) // 2) * L
    proto = A.kernel(np.arange(-half, half + 1) / L, sr, 16000) / L
    ysp = signal.resample_poly(x, L, M, window=proto)            # independent polyphase implementation, same prototype
    assert np.abs(ysp[:len(y)] - y).max() < 1e-6


def test_oracle_identity_and_edges():
    x = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert np.array_equal(A.resample(x, 16000, 16000), x)
    assert len(A.resample(np.zeros(0), 22050, 16000)) == 0
    assert len(A.resample(np.ones(1), 48000, 16000)) == 1
    y = A.resample(np.ones(4000), 32000, 16000)                  # DC gain 1 away from the (zero-extended) edges
    assert abs(y[1000] - 1.0) < 1e-6 and 0.45 < y[0] < 0.8        # a step edge sits half-way up (+ half the centre tap)


@pytest.mark.parametrize("kind", [W.U8, W.S16, W.S24, W.S32, W.F32, W.F64])
@pytest.mark.parametrize("ch", [1, 2, 3])
def test_oracle_decode_matches_library_readers(kind, ch):
    rng = np.random.default_rng(kind * 10 + ch)
    x = np.clip(0.4 * rng.standard_normal((777, ch)), -1, 0.999)
    data = W.wav_bytes(x, 22050, kind, extra_chunks=(ch == 2))
    k, c, sr, off, n = A.parse_wav(data)
    assert (k, c, sr, n) == (kind, ch, 22050, 777)
    got = A.decode_frames(data[off:off + n * ch * (1, 2, 3, 4, 4, 8)[kind]], kind, ch)
    assert got.dtype == np.float32 and got.shape == (777, ch)
    if kind == W.S16:                                            # stdlib reader of the same bytes, soundfile's int16 scaling
        with wave.open(io.BytesIO(data)) as w:
            ref = np.frombuffer(w.readframes(777), "<i2").reshape(-1, ch).astype(np.float32) / 32768
        assert np.array_equal(got, ref)
    if kind != W.S24:
        sr2, raw = wavfile.read(io.BytesIO(data))
        raw = raw.reshape(777, ch)
        scale = {W.U8: 128.0, W.S16: 32768.0, W.S32: 2147483648.0}.get(kind)
        ref = raw.astype(np.float32) if scale is None else ((raw.astype(np.float32) - (128 if kind == W.U8 else 0)) / np.float32(scale))
        assert sr2 == 22050 and np.array_equal(got, ref.astype(np.float32))
    mono = A.downmix_average(got)
    assert mono.dtype == np.float32 and np.array_equal(mono, np.mean(got, axis=-1) if ch > 1 else got[:, 0])


def test_extensible_header_and_truncated_data_chunk():
    x = 0.3 * np.random.default_rng(1).standard_normal((100, 2))
    data = W.wav_bytes(x, 44100, W.S24, extensible=True)
    assert A.parse_wav(data)[:3] == (A.FMT_S24, 2, 44100)
    k, c, sr, off, n = A.parse_wav(data[:-60])                   # file cut short: as many whole frames as are present
    assert n == 90


def test_product_header_walk_matches_oracle_and_rejects_other_containers(tmp_path):
    from desta_b200 import audio
    for kind, ext in ((W.S16, False), (W.F32, False), (W.S32, True), (W.F64, True)):
        data = W.wav_bytes(0.1 * np.ones((50, 2)), 24000, kind, extensible=ext, extra_chunks=True)
        p = tmp_path / f"a{kind}{int(ext)}.wav"
        p.write_bytes(data)
        with open(p, "rb") as f:
            info = audio.read_wav_info(f, len(data))
        assert (info.fmt, info.channels, info.sample_rate, info.data_offset, info.n_frames) == A.parse_wav(data)
    bad = tmp_path / "x.flac"
    bad.write_bytes(b"fLaC" + bytes(100))
    with pytest.raises(Exception, match="could not be decoded"):
        audio.AudioSegment.from_file(str(bad), target_sr=16000, channel_selector="average")
    with pytest.raises(Exception, match="could not be decoded"):
        audio.AudioSegment.from_file(str(tmp_path / "missing.wav"), target_sr=16000)


@pytest.mark.parametrize("sr", RATES + (16000,))
def test_host_filter_table_matches_oracle_kernel(sr):
    __graft_entry__.build()
    from desta_b200 import _lib
    L = _lib.lib()
    plan = _lib.ResamplePlan()
    nbytes = L.desta_resample_plan_bytes(sr, 16000, ctypes.byref(plan))
    g = math.gcd(sr, 16000)
    assert (plan.L, plan.M) == (16000 // g, sr // g) and nbytes == plan.L * plan.P * 4
    host = np.zeros(nbytes // 4, np.float32)
    assert L.desta_resample_plan_build(sr, 16000, host.ctypes.data_as(ctypes.c_void_p), nbytes) == 0
    if sr == 16000:
        assert plan.P == 1 and host[0] == 1.0
        return
    assert plan.P == A.n_taps(sr, 16000)
    r, j, H = np.arange(plan.L), np.arange(plan.P), plan.P // 2 - 1
    ref = A.kernel((((r * plan.M) % plan.L) / plan.L)[:, None] + (H - j)[None, :], sr, 16000)
    assert np.abs(ref.reshape(-1) - host).max() < 1e-7
    assert L.desta_resample_plan_bytes(0, 16000, None) == 0      # bad rates are refused


REF_ASSETS = "/root/reference/assets/audios"


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the build container")
def test_reference_assets_parse():
    seen = set()
    for name in sorted(os.listdir(REF_ASSETS)):
        data = open(os.path.join(REF_ASSETS, name), "rb").read()
        kind, ch, sr, off, n = A.parse_wav(data)
        seen.add((kind, sr))
        y, _ = A.load(data[:off + min(n, 4000) * ch * (1, 2, 3, 4, 4, 8)[kind]])
        assert y.dtype == np.float32 and np.isfinite(y).all() and np.abs(y).max() <= 1.0 + 1e-3
    assert (A.FMT_F32, 16000) in seen and (A.FMT_S16, 22050) in seen and (A.FMT_S16, 32000) in seen


@pytest.mark.parametrize("sr", [8000, 22050, 24000, 32000, 44100, 48000])
def test_oracle_agrees_with_an_independent_resampler_on_band_limited_input(sr):
    """torchaudio's Kaiser-windowed sinc resampler (its documented "kaiser_best" parameters: an independent third-party
    design with another window, width and roll-off) and the oracle must produce the same samples wherever the input has no
    energy above 0.8 of the lower Nyquist frequency: both are then ideal band-limited interpolators.  Same output length
    (ceil(n * 16000 / sr)), same time alignment, values within 2e-5."""
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(sr)
    n = sr // 2
    t = np.arange(n) / sr
    low = min(sr, 16000) / 2
    x = np.zeros(n)
    for _ in range(40):
        x += rng.uniform(0.01, 0.05) * np.sin(2 * np.pi * rng.uniform(50, 0.8 * low) * t + rng.uniform(0, 6.28))
    y = A.resample(x, sr, 16000)
    yt = ta.functional.resample(torch.from_numpy(x).float(), sr, 16000, lowpass_filter_width=64, rolloff=0.9475937167399596,
                                resampling_method="sinc_interp_kaiser", beta=14.769656459379492).numpy()
    assert len(yt) == len(y)
    assert np.abs(y - yt)[300:-300].max() < 2e-5


SOXR_FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_soxr.npz")


def passband_deviation(y, ref, sr):
    """max |y - ref| after removing everything above 0.85 x the lower Nyquist from the difference (the soxr_hq pass band ends
    at 0.913): inside the pass band every realisation of the specification must agree; the transition band is reported apart."""
    n = min(len(y), len(ref))
    d = np.fft.rfft((np.asarray(y[:n], np.float64) - np.asarray(ref[:n], np.float64)) * np.hanning(n))
    f = np.fft.rfftfreq(n, 1.0 / 16000)
    d_pass = d.copy()
    d_pass[f > 0.85 * min(sr, 16000) / 2] = 0
    return np.abs(np.fft.irfft(d_pass, n))[200:-200].max(), np.abs(np.fft.irfft(d - d_pass, n))[200:-200].max()


@pytest.mark.skipif(not os.path.exists(SOXR_FIXTURES), reason="tests/golden/resample_soxr.npz not generated yet "
                    "(oracle/make_soxr_fixtures.py needs librosa + python-soxr, absent from the build image): n3 resampler parity unpinned")
@pytest.mark.parametrize("sr", RATES)
def test_oracle_vs_soxr_fixtures(sr):
    """Pins the resampler restatement against the reference's own resampler (librosa soxr_hq) once the fixtures exist."""
    from oracle.make_soxr_fixtures import fixture_input
    gold = np.load(SOXR_FIXTURES)[f"y_{sr}"]
    y = A.resample(fixture_input(sr).astype(np.float64), sr, 16000)
    assert len(y) == len(gold)                                   # librosa fix_length: ceil(n * 16000 / sr)
    in_band, transition = passband_deviation(y, gold, sr)
    assert in_band < 2e-5, f"{sr} Hz: pass-band deviation from soxr_hq {in_band:.2e} (transition band {transition:.2e})"

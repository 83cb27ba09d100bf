"""Decode + resample oracle (numpy, float64 filter arithmetic).  TEST INFRASTRUCTURE ONLY.

Restates what the reference does between a file on disk and the 16 kHz mono float32 waveform
(ref: desta/utils/audio.py:305-315 soundfile read as float32, :147 + :221-243 sample conversion, :155 + :60-119
channel_selector="average", :161-166 librosa.core.resample(samples, orig_sr, target_sr); callers
desta/models/modeling_desta25.py:1553-1557, desta/trainer/data/simple_dataset.py:145-149).

PARITY UNPINNED for the resampling step.  The arithmetic lives in third-party code absent from this image:
librosa (unpinned in the reference's setup.py) dispatching its default res_type="soxr_hq" to python-soxr (libsoxr
"HQ": linear phase, pass band to 0.913 of the lower Nyquist frequency, stop band from 1.0, 20-bit ~ 120 dB), then
librosa.util.fix_length to ceil(n * target / orig) samples.  Neither package is installed, and the reference holds no
golden vectors for it, so this file restates the *published specification* - an ideal band-limited resampler with those
band edges, realised as a Kaiser-windowed sinc - and the tests pin it against what any implementation of that
specification must reproduce: pass-band sinusoids map to the analytically resampled sinusoid, stop-band sinusoids vanish,
scipy.signal.resample_poly with the same prototype agrees, torchaudio's independent Kaiser-sinc resampler ("kaiser_best"
parameters) agrees to 1e-5 on band-limited input (same length, same alignment), the identity rate is exact.  Inside the
transition band (7.3-8 kHz for a 16 kHz target) a different filter realisation may differ from soxr by design.

The decode and down-mix steps (everything a 16 kHz file goes through) ARE pinned: stdlib `wave` / scipy read the same
bytes and the scaling rules below are libsndfile's documented float conversion, which tests/golden fixtures already use.
"""
import math
import struct

import numpy as np

PASS, STOP, ATTEN_DB = 0.913, 1.0, 120.0
FMT_U8, FMT_S16, FMT_S24, FMT_S32, FMT_F32, FMT_F64 = range(6)


def parse_wav(data):
    """RIFF/WAVE header -> (fmt code, channels, sample_rate, byte offset of the frames, n_frames)."""
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt = 12, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = pos + 8
        if tag == b"fmt ":
            if len(data) < body + 16:
                raise ValueError("truncated fmt chunk")
            code, ch, sr, _, align, bits = struct.unpack("<HHIIHH", data[body:body + 16])
            if code == 0xFFFE:
                if len(data) < body + 26 or size < 26:
                    raise ValueError("truncated extensible fmt chunk")
                code = struct.unpack("<H", data[body + 24:body + 26])[0]
            fmt = (code, ch, sr, bits)
        elif tag == b"data":
            if fmt is None:
                raise ValueError("data chunk before fmt chunk")
            code, ch, sr, bits = fmt
            kind = {(1, 8): FMT_U8, (1, 16): FMT_S16, (1, 24): FMT_S24, (1, 32): FMT_S32, (3, 32): FMT_F32, (3, 64): FMT_F64}[(code, bits)]
            size = min(size, len(data) - body)
            return kind, ch, sr, body, size // (ch * bits // 8)
        pos = body + size + (size & 1)
    raise ValueError("no data chunk")


def decode_frames(raw, kind, channels):
    """Interleaved little-endian bytes -> float32 [n, channels] with libsndfile's float scaling (audio.py:307-315)."""
    if kind == FMT_U8:
        x = (np.frombuffer(raw, np.uint8).astype(np.float32) - np.float32(128)) / np.float32(128)
    elif kind == FMT_S16:
        x = np.frombuffer(raw, "<i2").astype(np.float32) / np.float32(32768)
    elif kind == FMT_S24:
        b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
        v = (b[:, 0] << 8 | b[:, 1] << 16 | b[:, 2] << 24) >> 8
        x = v.astype(np.float32) / np.float32(8388608)
    elif kind == FMT_S32:
        x = np.frombuffer(raw, "<i4").astype(np.float32) / np.float32(2147483648)
    elif kind == FMT_F32:
        x = np.frombuffer(raw, "<f4").astype(np.float32)
    else:
        x = np.frombuffer(raw, "<f8").astype(np.float32)
    return x.reshape(-1, channels)


def downmix_average(x):
    """select_channels(signal, "average") (audio.py:60-119): float32 mean over the channel axis; 1-D input passes through."""
    return x[:, 0] if x.shape[1] == 1 else np.mean(x, axis=-1)


def n_taps(sr_in, sr_out):
    low = min(sr_in, sr_out)
    delta = 0.5 * (STOP - PASS) * low / sr_in
    p = int(math.ceil((ATTEN_DB - 7.95) / (14.357 * delta)))
    return p + (p & 1)


def kernel(tau, sr_in, sr_out):
    """Continuous interpolation kernel g(tau), tau in input samples: Kaiser-windowed sinc, cut-off mid transition band."""
    low = min(sr_in, sr_out)
    fc = 0.25 * (PASS + STOP) * low / sr_in
    beta = 0.1102 * (ATTEN_DB - 8.7)
    W = 0.5 * n_taps(sr_in, sr_out)
    u = np.asarray(tau, np.float64) / W
    inside = np.abs(u) < 1.0
    win = np.i0(beta * np.sqrt(np.where(inside, 1.0 - u * u, 0.0))) / np.i0(beta)
    return np.where(inside, 2.0 * fc * np.sinc(2.0 * fc * np.asarray(tau, np.float64)) * win, 0.0)


def resample(x, sr_in, sr_out, chunk=4096):
    """y[n] = sum_k x[k] g(n sr_in / sr_out - k), n < ceil(len(x) sr_out / sr_in) (librosa fixes that length), zeros outside x."""
    x = np.asarray(x, np.float64)
    if sr_in == sr_out:
        return x.astype(np.float32)
    g = math.gcd(sr_in, sr_out)
    L, M = sr_out // g, sr_in // g
    n_out = -(-len(x) * L // M)
    P = n_taps(sr_in, sr_out)
    H = P // 2 - 1
    xp = np.concatenate([np.zeros(P), x, np.zeros(P + M)])
    y = np.empty(n_out, np.float64)
    j = np.arange(P)
    for a in range(0, n_out, chunk):
        n = np.arange(a, min(a + chunk, n_out), dtype=np.int64)
        k0 = (n * M) // L
        frac = ((n * M) % L) / L
        k = k0[:, None] - H + j[None, :]
        tau = frac[:, None] + (H - j)[None, :]
        y[a:a + len(n)] = np.sum(xp[np.clip(k + P, 0, len(xp) - 1)] * ((k >= -P) & (k < len(x) + P)) * kernel(tau, sr_in, sr_out), axis=1)
    return y.astype(np.float32)


def load(data, target_sr=16000):
    """bytes of a WAV file -> (mono float32 waveform at target_sr, original rate): AudioSegment.from_file(...).samples."""
    kind, ch, sr, off, n = parse_wav(data)
    bps = {FMT_U8: 1, FMT_S16: 2, FMT_S24: 3, FMT_S32: 4, FMT_F32: 4, FMT_F64: 8}[kind]
    mono = downmix_average(decode_frames(data[off:off + n * ch * bps], kind, ch))
    return resample(mono, sr, target_sr) if sr != target_sr else mono.astype(np.float32), sr

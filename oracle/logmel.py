"""Log-mel oracle (numpy, float64 internally).  TEST INFRASTRUCTURE ONLY.

Restates WhisperFeatureExtractor as the reference calls it
(ref: desta/models/modeling_desta25.py:1570, desta/trainer/data/simple_dataset.py:239-243):
  transformers/models/whisper/feature_extraction_whisper.py:135-164  (_torch_extract_fbank_features)
  transformers/models/whisper/feature_extraction_whisper.py:296-320  (pad / truncate to 480000)
  transformers/audio_utils.py:263-333, 453-546                       (slaney mel filter bank)
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_FREQ = 201


def _hz_to_mel_slaney(f):
    # audio_utils.py:263-296, mel_scale == "slaney"
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    mels = np.where(hi, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mels)
    return mels


def _mel_to_hz_slaney(m):
    # audio_utils.py:299-333
    m = np.asarray(m, dtype=np.float64)
    freq = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), freq)


def mel_filter_bank(n_mels: int) -> np.ndarray:
    """[201, n_mels] float64 slaney-normalised triangular filters, 0..8000 Hz.

    audio_utils.py:453-546 with norm="slaney", mel_scale="slaney"
    (feature_extraction_whisper.py:95-103)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(8000.0), n_mels + 2)
    filt_hz = _mel_to_hz_slaney(mel_pts)
    fft_hz = np.linspace(0, SAMPLE_RATE // 2, N_FREQ)
    # _create_triangular_filter_bank (audio_utils.py:356-376)
    fdiff = np.diff(filt_hz)
    slopes = filt_hz[None, :] - fft_hz[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filt_hz[2:n_mels + 2] - filt_hz[:n_mels])
    return fb * enorm[None, :]


def pad_or_truncate(wave) -> np.ndarray:
    """feature_extraction_whisper.py:296-320: first 480000 samples, right-padded with 0."""
    w = np.asarray(wave, dtype=np.float32).reshape(-1)[:N_SAMPLES]
    out = np.zeros(N_SAMPLES, dtype=np.float32)
    out[: w.shape[0]] = w
    return out


def logmel(waves, n_mels: int = 128) -> np.ndarray:
    """list of 1-D float arrays (any length) -> [N, n_mels, 3000] float32."""
    filt = mel_filter_bank(n_mels).astype(np.float32).astype(np.float64)  # cast to f32 as :152 does
    n = np.arange(N_FFT)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)  # periodic hann (torch.hann_window)
    out = np.empty((len(waves), n_mels, N_FRAMES), dtype=np.float32)
    for i, wave in enumerate(waves):
        w = pad_or_truncate(wave).astype(np.float64)
        wp = np.pad(w, (N_FFT // 2, N_FFT // 2), mode="reflect")  # torch.stft(center=True)
        idx = np.arange(N_FRAMES)[:, None] * HOP + n[None, :]    # frame 3000 dropped (:150)
        frames = wp[idx] * hann[None, :]
        spec = np.fft.rfft(frames, axis=1)                       # [3000, 201]
        power = spec.real ** 2 + spec.imag ** 2
        mel = power @ filt                                       # [3000, n_mels]
        logm = np.log10(np.maximum(mel, 1e-10))
        logm = np.maximum(logm, logm.max() - 8.0)                # per-clip max (:157-158)
        out[i] = ((logm + 4.0) / 4.0).T.astype(np.float32)
    return out

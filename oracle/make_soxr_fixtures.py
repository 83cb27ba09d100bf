"""Generate soxr_hq fixtures for scope row n3 (file -> 16 kHz waveform).  TEST INFRASTRUCTURE ONLY.

The reference resamples with ``librosa.core.resample(samples, orig_sr=sr, target_sr=16000)`` (default
``res_type="soxr_hq"``; ref: desta/utils/audio.py:161-166, called at desta/models/modeling_desta25.py:1553-1557 and
desta/trainer/data/simple_dataset.py:145-149).  Neither librosa nor python-soxr is installed in the build image, so this
script cannot run there; run it on any machine that has them (``pip install librosa soxr``):

    python oracle/make_soxr_fixtures.py            # writes tests/golden/resample_soxr.npz (~0.4 MB)

and commit the file.  ``tests/test_audio_io_cpu.py::test_oracle_vs_soxr_fixtures`` (oracle restatement) and
``tests/test_audio_io_gpu.py::test_device_resampler_vs_soxr_fixtures`` (CUDA kernel) load it when present and skip
otherwise; until then row n3's resampler stays "parity unpinned" (DESIGN.md §2).

Inputs are regenerated from seeds by the tests (``fixture_input`` below is imported by them), only soxr's OUTPUTS are
stored: for each source rate a 0.75 s mixture of pass-band sinusoids + a chirp that crosses the transition band +
low-level noise, mono, float32.  The tests compare inside the pass band to 2e-5 absolute and report the transition-band
deviation separately (another realisation of the same specification may differ there by design).
"""
import os
import sys

import numpy as np

RATES = (8000, 11025, 22050, 24000, 32000, 44100, 48000)
SECONDS = 0.75


def fixture_input(sr, seed=0):
    """float32 mono test signal at `sr` (deterministic): tones below 0.9 x the lower Nyquist, a chirp through it, noise."""
    rng = np.random.default_rng(1000 + sr + seed)
    n = int(sr * SECONDS)
    t = np.arange(n, dtype=np.float64) / sr
    nyq = 0.5 * min(sr, 16000)
    x = np.zeros(n)
    for f, a in ((0.031 * nyq, 0.30), (0.27 * nyq, 0.20), (0.61 * nyq, 0.15), (0.88 * nyq, 0.10)):
        x += a * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    f0, f1 = 0.80 * nyq, min(1.15 * nyq, 0.49 * sr)
    x += 0.05 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / SECONDS * t * t))
    x += 0.003 * rng.standard_normal(n)
    return x.astype(np.float32)


def main():
    try:
        import librosa
        import soxr
    except ImportError as e:
        sys.exit(f"needs librosa and python-soxr ({e}); see the module docstring")
    out = {"librosa_version": np.array(librosa.__version__), "soxr_version": np.array(soxr.__version__)}
    for sr in RATES:
        x = fixture_input(sr)
        y = librosa.core.resample(x, orig_sr=sr, target_sr=16000)      # exactly the reference's call
        out[f"y_{sr}"] = np.asarray(y, dtype=np.float32)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resample_soxr.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

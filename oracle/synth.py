"""Synthetic 16 kHz clips — numpy PCG64, identical on every machine.  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def synth_wave(kind: str, rng) -> np.ndarray:
    sr = 16000
    if kind == "noise_full":          # full 30 s, no padding
        return np.clip(0.1 * rng.standard_normal(480000), -1, 1).astype(np.float32)
    if kind == "speechy_12s":         # gated tone + noise, then zero padding -> exercises the max-8 floor
        t = np.arange(12 * sr) / sr
        gate = (np.sin(2 * np.pi * 0.7 * t) > 0).astype(np.float64)
        return (0.3 * np.sin(2 * np.pi * 220 * t) * gate + 0.05 * rng.standard_normal(t.size)).astype(np.float32)
    if kind == "short_3p5s":
        return (0.2 * rng.standard_normal(int(3.54 * sr))).astype(np.float32)
    if kind == "silence":             # all-zero clip -> every value -1.5
        return np.zeros(5 * sr, dtype=np.float32)
    if kind == "long_31s":            # longer than 30 s -> truncated
        return (0.05 * rng.standard_normal(31 * sr + 123)).astype(np.float32)
    if kind == "tone_clip":           # ragged length, not a multiple of the hop
        t = np.arange(7 * sr + 77) / sr
        return (0.8 * np.sin(2 * np.pi * 1000 * t) + 0.4 * np.sin(2 * np.pi * 3333 * t)).astype(np.float32)
    raise ValueError(kind)


def synth_waves(seed=7, kinds=("noise_full", "speechy_12s", "short_3p5s")):
    rng = np.random.default_rng(seed)
    return [synth_wave(k, rng) for k in kinds]

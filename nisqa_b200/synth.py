"""Deterministic "synthetic speech" used by the parity tests and bench.py.

Signal model: ~30 harmonics of f0 in [100, 250] Hz with 1/k roll-off and a slow vibrato, a
2-4 Hz syllabic AM envelope, white noise at -50..-20 dBFS, peak at -20..-6 dBFS, quantised
to PCM16.  The mel dB of such a clip spans about -60..+5 dB, so neither the -80 dB floor nor
the top_db clamp dominates.  ``numpy.random.default_rng(seed)`` (PCG64) is bit-stable across
platforms, so the same seed gives the same PCM here and on the GPU box.
"""
import numpy as np


def synth_speech_f32(seed, seconds=10.0, sr=48000):
    rng = np.random.default_rng(int(seed))
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    f0 = rng.uniform(100.0, 250.0)
    vib = 1.0 + 0.02 * np.sin(2 * np.pi * rng.uniform(3.0, 6.0) * t + rng.uniform(0, 2 * np.pi))
    phase = 2 * np.pi * f0 * np.cumsum(vib) / sr
    y = np.zeros(n, dtype=np.float64)
    n_h = 30
    for k in range(1, n_h + 1):
        if k * f0 * 1.02 >= 0.45 * sr:
            break
        y += (1.0 / k) * rng.uniform(0.5, 1.0) * np.sin(k * phase + rng.uniform(0, 2 * np.pi))
    am_f = rng.uniform(2.0, 4.0)
    env = 0.55 + 0.45 * np.sin(2 * np.pi * am_f * t + rng.uniform(0, 2 * np.pi))
    env *= 0.5 + 0.5 * (np.sin(2 * np.pi * 0.31 * t + rng.uniform(0, 2 * np.pi)) > -0.6)
    y *= env
    y /= max(np.max(np.abs(y)), 1e-9)
    peak = 10.0 ** (rng.uniform(-20.0, -6.0) / 20.0)
    noise = 10.0 ** (rng.uniform(-50.0, -20.0) / 20.0)
    y = peak * y + noise * rng.standard_normal(n)
    return np.clip(y, -1.0, 1.0).astype(np.float32)


def synth_speech_pcm16(seed, seconds=10.0, sr=48000):
    y = synth_speech_f32(seed, seconds, sr)
    return np.clip(np.round(y.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def ragged_durations(n, lo=2.0, hi=30.0, seed=1234):
    """config 3: durations ~ U(lo, hi) seconds, seeded."""
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=n)

"""Seeded synthetic 16 kHz audio shaped like the reference's inputs (no dataset is reachable).

Each clip is white noise + three stationary sinusoids (harmonic content) + unit impulses at
about four per second (percussive content), followed by the reference's own normalisation
(lib/preprocessing.py:114-132: subtract the mean, divide by the peak).  Seed = 1234 + clip index,
so the CPU oracle / baseline and the GPU path can regenerate identical bits.
"""
from __future__ import annotations

import numpy as np

FS = 16000


def synth_clip(index: int, n_samples: int, fs: int = FS) -> np.ndarray:
    rng = np.random.default_rng(1234 + int(index))
    t = np.arange(n_samples, dtype=np.float64) / fs
    x = 0.5 * rng.standard_normal(n_samples)
    for _ in range(3):
        f0 = rng.uniform(80.0, 4000.0)
        x += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 2 * np.pi))
    n_imp = max(1, int(round(4.0 * n_samples / fs)))
    pos = rng.integers(0, n_samples, size=n_imp)
    x[pos] += rng.uniform(4.0, 8.0, size=n_imp) * rng.choice([-1.0, 1.0], size=n_imp)
    x = x.astype(np.float32)
    x = x - np.mean(x)
    x = x / np.max(np.abs(x))
    return x.astype(np.float32)


def synth_batch(n_clips: int, n_samples: int, first_index: int = 0, fs: int = FS) -> np.ndarray:
    """(n_clips, n_samples) float32; clip i uses seed 1234 + first_index + i."""
    out = np.empty((n_clips, n_samples), dtype=np.float32)
    for i in range(n_clips):
        out[i] = synth_clip(first_index + i, n_samples, fs)
    return out


def synth_batch_fast(n_clips: int, n_samples: int, first_index: int = 0, fs: int = FS) -> np.ndarray:
    """Same statistics, vectorised over clips with one generator (seed 1234 + first_index); used to
    fill the full benchmark batch quickly.  Not bit-identical to synth_clip."""
    rng = np.random.default_rng(1234 + int(first_index))
    t = np.arange(n_samples, dtype=np.float32) / np.float32(fs)
    x = 0.5 * rng.standard_normal((n_clips, n_samples), dtype=np.float32)
    for _ in range(3):
        f0 = rng.uniform(80.0, 4000.0, size=(n_clips, 1)).astype(np.float32)
        ph = rng.uniform(0, 2 * np.pi, size=(n_clips, 1)).astype(np.float32)
        a = rng.uniform(0.3, 1.0, size=(n_clips, 1)).astype(np.float32)
        x += a * np.sin(np.float32(2 * np.pi) * f0 * t[None, :] + ph)
    n_imp = max(1, int(round(4.0 * n_samples / fs)))
    pos = rng.integers(0, n_samples, size=(n_clips, n_imp))
    amp = (rng.uniform(4.0, 8.0, size=(n_clips, n_imp)) * rng.choice([-1.0, 1.0], size=(n_clips, n_imp))).astype(np.float32)
    np.add.at(x, (np.arange(n_clips)[:, None], pos), amp)
    x -= x.mean(axis=1, keepdims=True)
    x /= np.abs(x).max(axis=1, keepdims=True)
    return x


def musan_like_durations(n_clips: int, seed: int = 7, mean_s: float = 30.0, min_s: float = 1.0,
                         max_s: float = 120.0) -> np.ndarray:
    """Variable clip lengths (samples) with a long-tailed distribution, for ragged-batch tests."""
    rng = np.random.default_rng(seed)
    d = np.clip(rng.exponential(mean_s, size=n_clips), min_s, max_s)
    return (d * FS).astype(np.int64)

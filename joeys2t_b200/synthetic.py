# coding: utf-8
"""
Synthetic LibriSpeech / MuST-C-shaped corpora for the benchmark configurations of BASELINE.json
(SURVEY.md §8d).  All audio is 16 kHz mono with **integer** sample values in int16 range (or
integer/32768 for the float32 variant) and silence as exact zeros — sub-LSB float amplitudes would
park mel energies at FLT_EPSILON where the log amplifies rounding noise for *any* implementation.
"""
from typing import List, Tuple

import numpy as np

SR = 16000


def _speechlike(rng: np.random.Generator, n: int) -> np.ndarray:
    """int16 'speech': pink-ish noise under a 4 Hz syllable envelope plus a white floor."""
    from scipy.signal import lfilter  # host-side data generation only
    white = rng.standard_normal(n)
    pink = np.zeros(n)
    for a, g in ((0.99, 0.55), (0.9, 0.3), (0.5, 0.15)):  # three leaky integrators
        y = lfilter([1.0 - a], [1.0, -a], white)
        pink += g * y / np.sqrt((1 - a) / (1 + a))
    t = np.arange(n) / SR
    env = 0.5 * (1.0 - np.cos(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi)))
    x = 3000.0 * env * pink + 30.0 * rng.standard_normal(n)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def _durations(rng, n, lo, hi):
    return rng.uniform(lo, hi, size=n)


def librispeech_batch(n_utts: int = 256, seed: int = 1234, lo: float = 10.0,
                      hi: float = 15.0) -> List[np.ndarray]:
    """Config 2: ``n_utts`` x U(10, 15) s int16; 5 % of utterances get a 0.3 s run of exact zeros."""
    rng = np.random.default_rng(seed)
    out = []
    for d in _durations(rng, n_utts, lo, hi):
        n = int(d * SR)
        x = _speechlike(rng, n)
        if rng.uniform() < 0.05:
            s = int(rng.integers(0, max(1, n - int(0.3 * SR))))
            x[s:s + int(0.3 * SR)] = 0
        out.append(x)
    return out


def mustc_batch(n_utts: int = 512, seed: int = 2345) -> List[np.ndarray]:
    """Config 3: ragged, log-normal durations (median 5 s) clipped to [1, 30] s, int16."""
    rng = np.random.default_rng(seed)
    dur = np.clip(np.exp(rng.normal(np.log(5.0), 0.7, size=n_utts)), 1.0, 30.0)
    return [_speechlike(rng, int(d * SR)) for d in dur]


def longform_batch(n_utts: int = 64, seed: int = 3456) -> List[np.ndarray]:
    """Config 4: U(30, 60) s; even indices int16, odd indices float32 in [-1, 1)."""
    rng = np.random.default_rng(seed)
    out = []
    for i, d in enumerate(_durations(rng, n_utts, 30.0, 60.0)):
        x = _speechlike(rng, int(d * SR))
        out.append(x if i % 2 == 0 else (x.astype(np.float32) / np.float32(32768.0)))
    return out


def pooled_batch(n_utts: int, seed: int, lo: float, hi: float, dtype=np.int16,
                 pool_seconds: float = 90.0) -> List[np.ndarray]:
    """LibriSpeech-shaped batch cut from one ``pool_seconds`` stretch of 'speech' (random offset,
    random integer-rounded gain per utterance; 5 % of utterances get 0.3 s of exact zeros).
    ~50x cheaper to generate than :func:`librispeech_batch`, same spectral character."""
    rng = np.random.default_rng(seed)
    pool = _speechlike(rng, int(pool_seconds * SR)).astype(np.float32)
    pool = np.concatenate([pool, pool])
    half = pool.shape[0] // 2
    out = []
    for d in _durations(rng, n_utts, lo, hi):
        n = int(d * SR)
        segs, need = [], n
        while need > 0:
            o = int(rng.integers(0, half))
            m = min(need, half)
            segs.append(pool[o:o + m])
            need -= m
        x = np.concatenate(segs) if len(segs) > 1 else segs[0]
        x = np.clip(np.rint(x * np.float32(rng.uniform(0.25, 2.0))), -32768, 32767).astype(np.int16)
        if rng.uniform() < 0.05:
            s = int(rng.integers(0, max(1, n - int(0.3 * SR))))
            x[s:s + int(0.3 * SR)] = 0
        out.append(x if dtype == np.int16 else x.astype(np.float32) / np.float32(32768.0))
    return out


def fast_noise_batch(n_utts: int, seed: int, lo: float, hi: float,
                     dtype=np.int16) -> List[np.ndarray]:
    """Cheap integer-valued noise (shaped by a one-tap low-pass) for large sweeps where generating
    'speech' on the host would dominate the run."""
    rng = np.random.default_rng(seed)
    out = []
    for d in _durations(rng, n_utts, lo, hi):
        n = int(d * SR)
        w = rng.integers(-3000, 3000, size=n + 1, dtype=np.int32)
        x = ((w[1:] + w[:-1]) // 2).astype(np.int16)
        out.append(x if dtype == np.int16 else x.astype(np.float32) / np.float32(32768.0))
    return out


def total_audio_hours(waves: List[np.ndarray]) -> float:
    return sum(int(w.shape[-1]) for w in waves) / SR / 3600.0


def corpus_shape(waves: List[np.ndarray]) -> Tuple[int, float, int]:
    from joeys2t_b200 import tables
    frames = sum(tables.num_frames(int(w.shape[-1])) for w in waves)
    return len(waves), total_audio_hours(waves), frames

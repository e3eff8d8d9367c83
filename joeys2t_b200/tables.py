# coding: utf-8
"""
Host-side constant tables of the Kaldi-compatible fbank front-end.

The reference gets these from ``torchaudio.compliance.kaldi`` (third-party, not vendored):
``_feature_window_function`` (kaldi.py:86-113, povey = ``hann(400, periodic=False)**0.85``) and
``get_mel_banks`` (kaldi.py:436-511).  torchaudio builds them with float32 torch CPU ops; the
same op sequence is restated here with torch float32 ops so the tables are *bit-identical* to
the reference's (checked against ``tests/golden/ref_tables.npz``), without importing torchaudio.

Everything here runs once per process on the host; it is not on the hot path.
"""
import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000
FRAME_LENGTH = 400  # int(16000 * 25 ms)      kaldi.py:139
FRAME_SHIFT = 160  # int(16000 * 10 ms)       kaldi.py:138
FFT_SIZE = 512  # next power of two           kaldi.py:140
NUM_FFT_BINS = FFT_SIZE // 2
NUM_MEL_BINS = 80
LOW_FREQ = 20.0
PREEMPHASIS = 0.97


def frame_geometry(sample_rate: int) -> Tuple[int, int, int]:
    """(shift, length, padded length) for a sample rate — kaldi.py:138-140."""
    shift = int(sample_rate * 10.0 * 0.001)
    length = int(sample_rate * 25.0 * 0.001)
    padded = 1 if length == 0 else 2**(length - 1).bit_length()
    return shift, length, padded


def num_frames(num_samples: int) -> int:
    """snip_edges=True frame count at 16 kHz — kaldi.py:63-67."""
    if num_samples < FRAME_LENGTH:
        return 0
    return 1 + (num_samples - FRAME_LENGTH) // FRAME_SHIFT


@lru_cache(maxsize=None)
def povey_window() -> np.ndarray:
    """(400,) float32, ``w[0] == w[399] == 0`` — kaldi.py:98-100."""
    w = torch.hann_window(FRAME_LENGTH, periodic=False, dtype=torch.float32).pow(0.85)
    return w.numpy().copy()


@lru_cache(maxsize=None)
def mel_banks(num_bins: int = NUM_MEL_BINS) -> np.ndarray:
    """(num_bins, 256) float32 triangular mel filters, 20 Hz … Nyquist — kaldi.py:436-511
    with ``vtln_warp_factor == 1``, ``high_freq = 0`` (→ Nyquist), 512-point FFT, 16 kHz."""
    assert num_bins > 3, "Must have at least 3 mel bins"
    nyquist = 0.5 * SAMPLE_RATE
    high_freq = 0.0 + nyquist
    fft_bin_width = SAMPLE_RATE / FFT_SIZE
    mel_low = 1127.0 * math.log(1.0 + LOW_FREQ / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)

    b = torch.arange(num_bins).unsqueeze(1)  # int64; promoted to float32 by the python scalars
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(float(NUM_FFT_BINS))) / 700.0).log())
    mel = mel.unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return bins.numpy().copy()


def mel_two_band(banks: np.ndarray):
    """Re-express the (num_bins, 256) bank as the streaming two-band form the kernel uses.

    Every FFT bin k feeds at most two *adjacent* filters (kaldi.py:500-505: consecutive
    triangles share their edges).  ``seg[k]`` is the index of the upper one, ``wu[k]`` its
    (rising-slope) weight and ``wd[k]`` the (falling-slope) weight of filter ``seg[k]-1``:

        mel[m] = sum_{k: seg[k]==m} wu[k] P[k]  +  sum_{k: seg[k]==m+1} wd[k] P[k]

    :returns: (seg int32[256] in 0..num_bins, wu float32[256], wd float32[256])
    :raises ValueError: if the bank does not have that structure (it always does for the
        reference's fixed arguments; other banks are rejected rather than silently mis-evaluated).
    """
    num_bins, nk = banks.shape
    seg = np.zeros(nk, np.int32)
    wu = np.zeros(nk, np.float32)
    wd = np.zeros(nk, np.float32)
    peak = np.array([int(np.argmax(r)) for r in banks])
    for k in range(nk):
        nz = np.nonzero(banks[:, k])[0]
        if len(nz) == 0:
            # no weight: keep the segment monotone
            seg[k] = seg[k - 1] if k else 0
            continue
        if len(nz) == 2 and nz[1] - nz[0] == 1:
            s = int(nz[1])
        elif len(nz) == 1:
            m = int(nz[0])
            s = m if k <= peak[m] else m + 1
        else:
            raise ValueError(f"mel bank is not two-band at FFT bin {k}: filters {nz.tolist()}")
        seg[k] = s
        if s < num_bins:
            wu[k] = banks[s, k]
        if s >= 1:
            wd[k] = banks[s - 1, k]
    if np.any(np.diff(seg) < 0):
        raise ValueError("mel bank segments are not monotone in frequency")
    # exact reconstruction check
    dense = np.zeros_like(banks)
    for k in range(nk):
        if seg[k] < num_bins:
            dense[seg[k], k] = wu[k]
        if seg[k] >= 1:
            dense[seg[k] - 1, k] = wd[k]
    if not np.array_equal(dense, banks):
        raise ValueError("mel bank cannot be represented in two-band form")
    return seg, wu, wd

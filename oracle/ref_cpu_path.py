# coding: utf-8
"""
ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ``joeys2t_b200/``).

The reference's CPU path for one utterance.  Three tiers, best available first (``KIND`` / ``KIND_TAG``
say which one is in use so that bench.py can report it):

1. ``"reference"`` — the **unmodified reference**: ``joeynmt.helpers_for_audio.extract_fbank_features``
   (helpers_for_audio.py:41-68) -> ``joeynmt.data_augmentation.CMVN()`` (data_augmentation.py:96-109),
   imported from ``/root/reference`` in the build container or from the installed copy ``oracle/_ref``
   (``oracle/build_ref.sh``) on the GPU box, through the two import shims of ``oracle/ref_shims.py``.
2. ``"port"`` — ``torchaudio.compliance.kaldi.fbank`` (the third-party dependency that holds all of the
   arithmetic; part of the image) called with exactly the reference's arguments + the restated glue /
   CMVN of ``oracle/fbank_numpy.py``.
3. ``"port"`` — the pure numpy restatement, when torchaudio is missing too.
"""
import os

import numpy as np

from oracle import fbank_numpy as O

_ref_extract = None
_ref_cmvn = None
_ta_kaldi = None
try:
    import torch
except Exception:  # pylint: disable=broad-except
    torch = None

if torch is not None and not os.environ.get("JS2T_CPU_BASELINE_PORT"):
    try:
        from oracle import ref_shims
        if ref_shims.reference_available():
            _helpers = ref_shims.install(full_stack=False)
            import importlib
            _ref_extract = _helpers.extract_fbank_features
            _ref_cmvn = importlib.import_module("joeynmt.data_augmentation").CMVN()
    except Exception:  # pylint: disable=broad-except
        _ref_extract = None

if _ref_extract is not None:
    KIND_TAG = "reference"
    KIND = ("unmodified reference: joeynmt.helpers_for_audio.extract_fbank_features -> "
            f"joeynmt.data_augmentation.CMVN() imported from {ref_shims.REFERENCE_ROOT}")
else:
    KIND_TAG = "port"
    try:  # the reference's own dependency (requirements.txt:6)
        import torchaudio.compliance.kaldi as _ta_kaldi
        KIND = "torchaudio.compliance.kaldi.fbank (the reference's own dependency) + restated joeynmt glue/CMVN"
    except Exception:  # pylint: disable=broad-except
        _ta_kaldi = None
        KIND = "numpy restatement (oracle/fbank_numpy.py)"


def _as_loader_waveform(wave: np.ndarray):
    """(1, N) float32 in [-1, 1) — what ``torchaudio.load`` hands to the reference
    (helpers_for_audio.py:115; int16 / 32768 is exact)."""
    if wave.dtype == np.int16:
        return torch.from_numpy(wave.astype(np.float32) / np.float32(32768.0))[None]
    return torch.from_numpy(np.asarray(wave, np.float32))[None]


def fbank(wave: np.ndarray) -> np.ndarray:
    """(T, 80) log-mel of one 16 kHz utterance (int16 PCM, or float in [-1, 1))."""
    if _ref_extract is not None:
        return _ref_extract(_as_loader_waveform(wave), 16000)
    if _ta_kaldi is None:
        return O.extract_fbank_features(wave)
    if wave.dtype == np.int16:
        x = torch.from_numpy(wave.astype(np.float32))  # == (int16 / 32768) * 2**15, exact (Q4)
    else:
        x = torch.from_numpy(np.asarray(wave, np.float32)) * (2**15)  # helpers_for_audio.py:54
    return _ta_kaldi.fbank(x[None], num_mel_bins=80, sample_frequency=16000).numpy()


def fbank_cmvn(wave: np.ndarray) -> np.ndarray:
    if _ref_extract is not None:
        return _ref_cmvn(fbank(wave))
    return O.cmvn(fbank(wave))


def single_thread() -> None:
    """One thread per worker process (the reference's prep scripts fan out over processes,
    scripts/prepare_mustc.py:53,118)."""
    if torch is not None:
        torch.set_num_threads(1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # pylint: disable=broad-except
        pass

# coding: utf-8
"""
ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ``joeys2t_b200/``).

The reference's CPU path for one utterance, as close to the real thing as the GPU box allows:

    joeynmt/helpers_for_audio.py:41-68   extract_fbank_features  (waveform * 2**15 -> fbank)
    joeynmt/helpers_for_audio.py:30-37   _get_torchaudio_fbank   (ta_kaldi.fbank(..., num_mel_bins=80))
    joeynmt/data_augmentation.py:96-109  CMVN.__call__

``/root/reference`` does not travel to the GPU box, but the third-party dependency that holds all
of the arithmetic does: ``torchaudio.compliance.kaldi.fbank`` is part of the image.  When it is
importable this module calls *it* with exactly the reference's arguments and adds the restated glue
(``oracle/fbank_numpy.cmvn``); otherwise it falls back to the pure numpy restatement.  ``KIND``
says which one is in use so that bench.py can report it.
"""
import numpy as np

from oracle import fbank_numpy as O

try:  # the reference's own dependency (requirements.txt:6)
    import torch
    import torchaudio.compliance.kaldi as _ta_kaldi
    KIND = "torchaudio.compliance.kaldi.fbank (the reference's own dependency) + restated joeynmt glue/CMVN"
except Exception:  # pylint: disable=broad-except
    _ta_kaldi = None
    KIND = "numpy restatement (oracle/fbank_numpy.py)"


def fbank(wave: np.ndarray) -> np.ndarray:
    """(T, 80) log-mel of one 16 kHz utterance (int16 PCM, or float in [-1, 1))."""
    if _ta_kaldi is None:
        return O.extract_fbank_features(wave)
    if wave.dtype == np.int16:
        x = torch.from_numpy(wave.astype(np.float32))  # == (int16 / 32768) * 2**15, exact (Q4)
    else:
        x = torch.from_numpy(np.asarray(wave, np.float32)) * (2**15)  # helpers_for_audio.py:54
    return _ta_kaldi.fbank(x[None], num_mel_bins=80, sample_frequency=16000).numpy()


def fbank_cmvn(wave: np.ndarray) -> np.ndarray:
    return O.cmvn(fbank(wave))


def single_thread() -> None:
    """One thread per worker process (the reference's prep scripts fan out over processes,
    scripts/prepare_mustc.py:53,118)."""
    if _ta_kaldi is not None:
        torch.set_num_threads(1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # pylint: disable=broad-except
        pass

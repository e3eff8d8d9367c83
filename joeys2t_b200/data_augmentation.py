# coding: utf-8
"""
Data Augmentation — B200 drop-in for ``joeynmt/data_augmentation.py`` (same class names,
constructor arguments, ``__call__`` signature, ``__repr__`` and ``.before`` attribute; host numpy
float32 in, host numpy float32 out, inputs never mutated).

* ``CMVN``         ← joeynmt/data_augmentation.py:83-115
* ``SpecAugment``  ← joeynmt/data_augmentation.py:15-80  (mask positions are drawn on the host from
  the global ``np.random`` generator in exactly the reference's order, so a seeded run masks the
  same cells bit for bit; the fill happens on the GPU)

Both run on the GPU through the C ABI (``js2t_features_execute``).  The per-item call pays a
host↔device round trip; throughput comes from the batched API in :mod:`joeys2t_b200.frontend`.
"""
import math
from typing import Optional, Tuple

import numpy as np

from joeys2t_b200 import frontend

NUM_FREQ = frontend.NUM_MEL


def _check_2d(x: np.ndarray, what: str):
    assert len(x.shape) == 2, f"{what} must be a 2-D tensor."
    if x.shape[1] != NUM_FREQ:
        raise ValueError(
            f"joeys2t_b200 is specialised for num_freq={NUM_FREQ}; got {x.shape[1]} (no CPU fallback)")


def draw_masks(num_frames: int, num_freqs: int, freq_mask_n: int, freq_mask_f: int,
               time_mask_n: int, time_mask_t: int, time_mask_p: float) -> Optional[np.ndarray]:
    """Replay of the RNG draws of joeynmt/data_augmentation.py:48-70 on the global ``np.random``.

    :returns: int32 (freq_mask_n + time_mask_n, 2) table of (start, width) — width 0 rows are
        no-ops whose draws were still consumed, like the reference — or ``None`` when the reference
        returns its input untouched (no frames, or fewer frequency bins than ``freq_mask_f``).
    """
    rows = _draw_mask_rows(num_frames, num_freqs, freq_mask_n, freq_mask_f, time_mask_n, time_mask_t,
                           time_mask_p)
    return None if rows is None else np.array(rows, np.int32)


def _draw_mask_rows(num_frames, num_freqs, freq_mask_n, freq_mask_f, time_mask_n, time_mask_t, time_mask_p):
    """The draws themselves, as a list of (start, width) tuples (one numpy array per BATCH is built by
    :func:`mask_tables_for_batch`; per-row array writes were most of the per-batch host time)."""
    if num_frames == 0 or num_freqs < freq_mask_f:  # :48-52
        return None
    randint = np.random.randint
    rows = []
    for _ in range(freq_mask_n):  # :54-58
        f = int(randint(0, freq_mask_f))
        f0 = int(randint(0, num_freqs - f))
        rows.append((f0, f))
    max_time_mask_t = min(time_mask_t, math.floor(num_frames * time_mask_p))  # :60-62
    if max_time_mask_t < 1:  # :63-64 frequency masks only
        rows.extend([(0, 0)] * time_mask_n)
        return rows
    for _ in range(time_mask_n):  # :66-70
        t = int(randint(0, max_time_mask_t))
        t0 = int(randint(0, num_frames - t))
        rows.append((t0, t))
    return rows


class SpecAugment:
    """
    SpecAugment (https://arxiv.org/abs/1904.08779); interface of joeynmt/data_augmentation.py:15-80
    """

    def __init__(
        self,
        freq_mask_n: int = 2,
        freq_mask_f: int = 27,
        time_mask_n: int = 2,
        time_mask_t: int = 40,
        time_mask_p: float = 1.0,
        mask_value: Optional[float] = None
    ):
        self.freq_mask_n = freq_mask_n
        self.freq_mask_f = freq_mask_f
        self.time_mask_n = time_mask_n
        self.time_mask_t = time_mask_t
        self.time_mask_p = time_mask_p
        self.mask_value = mask_value

    def draw(self, num_frames: int, num_freqs: int = NUM_FREQ) -> Optional[np.ndarray]:
        """Host-side mask table for one spectrogram (consumes the global numpy RNG)."""
        return draw_masks(num_frames, num_freqs, self.freq_mask_n, self.freq_mask_f,
                          self.time_mask_n, self.time_mask_t, self.time_mask_p)

    def __call__(self, spectrogram: np.ndarray) -> np.ndarray:
        _check_2d(spectrogram, "spectrogram")
        num_frames, num_freqs = spectrogram.shape
        table = self.draw(num_frames, num_freqs)
        if table is None:
            return spectrogram  # :48-52 (the reference returns the input object itself)
        out, _ = frontend.features_cmvn_specaug_ragged(
            [spectrogram], masks=table[None], n_fmask=self.freq_mask_n, n_tmask=self.time_mask_n,
            mask_value=self.mask_value)
        distorted = out.cpu().numpy()
        assert distorted.shape == spectrogram.shape
        return distorted

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(freq_mask_n={self.freq_mask_n}, "
            f"freq_mask_f={self.freq_mask_f}, time_mask_n={self.time_mask_n}, "
            f"time_mask_t={self.time_mask_t}, time_mask_p={self.time_mask_p})"
        )


class CMVN:
    """
    CMVN: Cepstral Mean and Variance Normalization (Utterance-level);
    interface of joeynmt/data_augmentation.py:83-115
    """

    def __init__(
        self, norm_means: bool = True, norm_vars: bool = True, before: bool = True
    ):
        self.norm_means = norm_means
        self.norm_vars = norm_vars
        self.before = before

    def config(self) -> dict:
        return dict(norm_means=self.norm_means, norm_vars=self.norm_vars, before=self.before)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        _check_2d(x, "x")
        orig_shape = x.shape
        out, _ = frontend.features_cmvn_specaug_ragged(
            [x], cmvn=dict(norm_means=self.norm_means, norm_vars=self.norm_vars, before=True))
        y = out.cpu().numpy()
        assert orig_shape == y.shape
        return y

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(norm_means={self.norm_means}, "
            f"norm_vars={self.norm_vars}, before={self.before})"
        )


def mask_tables_for_batch(specaugment: "SpecAugment", n_frames,
                          num_freqs: int = NUM_FREQ) -> Tuple[np.ndarray, int, int]:
    """Draw the SpecAugment tables of a whole batch in utterance order (= the order in which the
    reference's per-item loop would consume the RNG).  Utterances the reference leaves untouched
    get all-zero-width rows.

    The draws are replayed from raw outputs of the global ``np.random`` stream by ``js2t_specaug_replay``
    (host C code): exactly the values and exactly the stream position the per-item ``randint`` calls of
    :func:`draw_masks` give, at a tenth of their cost for a 16-utterance batch."""
    sa = specaugment
    n_masks = sa.freq_mask_n + sa.time_mask_n
    nf = np.ascontiguousarray(n_frames, np.int32)
    if _FAST_DRAWS and n_masks > 0 and len(nf) > 1 and sa.freq_mask_f >= 1 and num_freqs >= sa.freq_mask_f \
            and int(nf.min()) >= 0:
        table = _replay_tables(sa, nf, num_freqs, n_masks)
        if table is not None:
            return table, sa.freq_mask_n, sa.time_mask_n
    empty = [(0, 0)] * n_masks
    rows = []
    for t in n_frames:
        r = _draw_mask_rows(int(t), num_freqs, sa.freq_mask_n, sa.freq_mask_f, sa.time_mask_n,
                            sa.time_mask_t, sa.time_mask_p)
        rows.append(empty if r is None else r)
    table = np.array(rows, np.int32).reshape(len(rows), n_masks, 2)
    return table, sa.freq_mask_n, sa.time_mask_n


_FAST_DRAWS = True  # (tests switch it off to compare the two routes)


_RNG_HOOK = None  # (bit generator object, address of its next_uint32, its state pointer, its lock)


def _replay_tables(sa, nf: np.ndarray, num_freqs: int, n_masks: int) -> Optional[np.ndarray]:
    """See :func:`mask_tables_for_batch`.  ``None`` = take the per-item route (the C library is not
    available, or the configuration is one the reference itself rejects)."""
    global _RNG_HOOK
    import ctypes

    from joeys2t_b200 import _lib
    try:
        lib = _lib.load()
    except Exception:  # pylint: disable=broad-except
        return None
    bg = np.random.mtrand._rand._bit_generator  # behind np.random.randint / np.random.seed
    hook = _RNG_HOOK
    if hook is None or hook[0] is not bg:
        iface = bg.ctypes
        hook = _RNG_HOOK = (bg, ctypes.cast(iface.next_uint32, ctypes.c_void_p).value, iface.state, bg.lock)
    table = np.empty((len(nf), n_masks, 2), np.int32)
    with hook[3]:
        rc = lib.js2t_specaug_replay(hook[1], hook[2], len(nf), nf.ctypes.data, int(num_freqs),
                                     int(sa.freq_mask_n), int(sa.freq_mask_f), int(sa.time_mask_n),
                                     int(sa.time_mask_t), float(sa.time_mask_p), table.ctypes.data)
    return table if rc == _lib.OK else None

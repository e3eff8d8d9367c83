# coding: utf-8
"""
``SpeechProcessor`` — B200 drop-in for ``joeynmt/tokenizers.py:433-508`` (same constructor
arguments, ``__call__(line, is_train)`` contract, length filters and CMVN / SpecAugment order),
plus :meth:`SpeechProcessor.process_batch`, the batched entry point a sampler / collate function
hands whole index lists to: one fused GPU pass for the batch, output already in
the padded ``(B, Tmax, 80)`` layout of ``pad_features`` on the device.
"""
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from joeys2t_b200 import frontend, tables
from joeys2t_b200.data_augmentation import CMVN, SpecAugment, mask_tables_for_batch
from joeys2t_b200.helpers_for_audio import get_features, load_waveform


class SpeechProcessor:
    """SpeechProcessor (joeynmt/tokenizers.py:433-508)"""

    def __init__(
        self,
        level: str = "frame",
        num_freq: int = 80,
        normalize: bool = False,
        max_length: int = -1,
        min_length: int = -1,
        **kwargs,
    ):
        self.level = level
        self.num_freq = num_freq
        self.normalize = normalize

        # filter by length
        self.max_length = max_length
        self.min_length = min_length

        self.specaugment: Callable = SpecAugment(**kwargs["specaugment"]) \
            if "specaugment" in kwargs else None
        self.cmvn: Callable = CMVN(**kwargs["cmvn"]) if "cmvn" in kwargs else None
        self.root_path = ""  # assigned later in dataset.__init__()

    # ---- per item (reference contract) ------------------------------------------------------
    def __call__(self, line: str, is_train: bool = False) -> np.ndarray:
        """
        get features

        :param line: path to audio file or pre-extracted features
        :param is_train:

        :return: spectrogram in shape (num_frames, num_freq), or None if filtered out
        """
        path, *extra = line.split(":")
        full = Path(self.root_path) / path
        if len(extra) == 0 and full.suffix in (".wav", ".mp3"):
            if not full.is_file():
                raise FileNotFoundError(f"File not found: {full}")
            waveform, sample_rate = load_waveform(full)
            if int(sample_rate) != tables.SAMPLE_RATE:
                raise ValueError(f"{full}: {sample_rate} Hz audio; the front-end needs 16 kHz")
            out = self.process_batch([waveform], is_train=is_train, layout="ragged")
            feats, lengths, keep = out
            if not keep[0]:
                return None
            return feats.cpu().numpy()

        # pre-extracted features (.npy / zip): length filters on the host, CMVN / SpecAugment fused
        item = get_features(self.root_path, line)  # shape = (num_frames, num_freq)
        num_frames, num_freq = item.shape
        assert num_freq == self.num_freq

        if self._filter_too_short_item(num_frames):
            return None
        if self._filter_too_long_item(num_frames):
            if is_train:  # pylint: disable=no-else-return
                return None
            else:  # in test, truncate the sequence
                item = item[:self.max_length, :]
                num_frames = item.shape[0]
                assert num_frames <= self.max_length
        if self.cmvn is None and not (is_train and self.specaugment):
            return item
        table, nf, nt = self._draw([num_frames], is_train)
        out, _ = frontend.features_cmvn_specaug_ragged(
            [item], cmvn=self.cmvn.config() if self.cmvn else None, masks=table, n_fmask=nf,
            n_tmask=nt, mask_value=self.specaugment.mask_value if table is not None else None)
        return out.cpu().numpy()

    # ---- batched (device-resident) ----------------------------------------------------------
    def _draw(self, n_frames: Sequence[int], is_train: bool):
        if not (is_train and self.specaugment):
            return None, 0, 0
        return mask_tables_for_batch(self.specaugment, n_frames, self.num_freq)

    def process_batch(self, waveforms: Sequence, is_train: bool = False, layout: str = "padded",
                      pad_index: int = 1) -> Tuple[Optional[torch.Tensor], List[int], List[bool]]:
        """Whole batch of waveforms → features on the GPU in one fused pass.

        Applies, per utterance, exactly what ``__call__`` does (tokenizers.py:458-494): drop if
        ``0 < T < min_length``; if ``T > max_length > 0`` drop (train) or truncate to the first
        ``max_length`` frames *before* CMVN (eval); CMVN(before) → SpecAugment (train) →
        CMVN(after).  SpecAugment tables are drawn on the host in utterance order from the global
        ``np.random`` — the order the reference's per-item loop consumes it.

        :returns: (features, lengths, keep) — features is ``(B', Tmax, 80)`` padded with
            ``float(pad_index)`` (or ragged ``(sum T, 80)``) for the B' kept utterances on the GPU,
            ``lengths`` their frame counts, ``keep[i]`` whether input i survived the filters.
        """
        assert self.num_freq == tables.NUM_MEL_BINS
        n_all = [tables.num_frames(int(np.asarray(w).shape[-1])) for w in waveforms]
        keep, max_frames = [], []
        for t in n_all:
            k = True
            if t <= 0:
                raise ValueError("choose a window size 400 that is [2, N]: waveform shorter than "
                                 "one 25 ms frame")
            if self._filter_too_short_item(t):
                k = False
            elif self._filter_too_long_item(t):
                if is_train:
                    k = False
            keep.append(k)
            if k:
                max_frames.append(self.max_length if self._filter_too_long_item(t) else 0)
        kept = [w for w, k in zip(waveforms, keep) if k]
        if not kept:
            return None, [], keep
        lengths = [min(t, m) if m > 0 else t
                   for t, m in zip((t for t, k in zip(n_all, keep) if k), max_frames)]
        table, nf, nt = self._draw(lengths, is_train)
        feats, n_frames = frontend.fbank_cmvn_specaug_ragged(
            kept, cmvn=self.cmvn.config() if self.cmvn else None, masks=table, n_fmask=nf,
            n_tmask=nt, mask_value=self.specaugment.mask_value if table is not None else None,
            max_frames=max_frames, layout=layout, pad_value=float(pad_index))
        assert n_frames.tolist() == lengths
        return feats, lengths, keep

    def _filter_too_short_item(self, length: int) -> bool:
        return self.min_length > length > 0

    def _filter_too_long_item(self, length: int) -> bool:
        return length > self.max_length > 0

    def __repr__(self):
        return (
            f"{self.__class__.__name__}("
            f"level={self.level}, normalize={self.normalize}, "
            f"filter_by_length=({self.min_length}, {self.max_length}), "
            f"cmvn={self.cmvn}, specaugment={self.specaugment})"
        )

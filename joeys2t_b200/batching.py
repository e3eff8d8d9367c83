# coding: utf-8
"""
Batched sampler / collate path — SURVEY.md §8(f-2).

The reference builds its batches by *computing every item's features twice*: the batch samplers
call ``d[idx]`` only to learn whether the item survives the length filters and how long it is
(``joeynmt/datasets.py:1199``, ``:1274``), and the ``DataLoader`` then calls ``__getitem__`` again
for the items of the batch (quirk Q5).  Here

* :class:`FrameCountBatchSampler` takes those decisions from the manifest's ``n_frames`` column
  (and the tokenised target lengths) with the very same rules — ``SpeechProcessor``'s length
  filters (``tokenizers.py:473-484``), ``SentenceBatchSampler`` / ``TokenBatchSampler`` batching
  (``datasets.py:1194-1211``, ``:1267-1292``) — and yields the same index lists without touching
  any audio, and
* :class:`SpeechBatchCollator` hands the whole index list to the fused GPU front-end once and
  returns ``src`` already in ``pad_features``' ``(B, Tmax, 80)`` layout (pad value
  ``float(pad_index)``, ``helpers_for_audio.py:151-152``) on the device together with
  ``src_length`` — what ``collate_fn`` (``datasets.py:207-225``) + ``Batch._make_cuda``
  (``batch.py:114-121``) produce with three host copies.
"""
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


def kept_length(n_frames: int, min_length: int = -1, max_length: int = -1,
                is_train: bool = False) -> Optional[int]:
    """Length of an item after ``SpeechProcessor.__call__``'s filters (tokenizers.py:473-484,496-500),
    or ``None`` if the item is dropped: ``0 < T < min_length`` → dropped; ``T > max_length > 0`` →
    dropped when training, truncated to ``max_length`` otherwise."""
    t = int(n_frames)
    if min_length > t > 0:
        return None
    if t > max_length > 0:
        return None if is_train else int(max_length)
    return t


class FrameCountBatchSampler:
    """Yields lists of dataset indices exactly like the reference's ``SentenceBatchSampler``
    (``batch_type="sentence"``: ``batch_size`` items per batch) or ``TokenBatchSampler``
    (``batch_type="token"``: close the batch once ``max_tokens * len(batch) >= batch_size`` with
    ``n_tokens = max(src_len + 1, trg_len + 1)``), but from frame counts instead of features.

    :param sampler: any iterable of dataset indices (the reference's ``RandomSubsetSampler`` /
        ``DistributedSubsetSampler`` objects work unchanged)
    :param n_frames: frame count per dataset index (TSV column ``n_frames``)
    :param trg_len: tokenised target length per dataset index; ``< 0`` = the target was filtered out,
        which drops the item (``datasets.py:650-653``); ``None`` = no targets (test without ``trg``)
    """

    def __init__(self, sampler: Iterable[int], batch_size: int, batch_type: str = "sentence",
                 drop_last: bool = False, *, n_frames: Sequence[int],
                 trg_len: Optional[Sequence[int]] = None, min_length: int = -1,
                 max_length: int = -1, is_train: bool = False):
        if batch_type not in ("sentence", "token"):
            raise ValueError(f"{batch_type}: Unknown batch type")
        self.sampler = sampler
        self.batch_size = int(batch_size)
        self.batch_type = batch_type
        self.drop_last = drop_last
        self.n_frames = np.asarray(n_frames, dtype=np.int64)
        self.trg_len = None if trg_len is None else np.asarray(trg_len, dtype=np.int64)
        self.min_length, self.max_length, self.is_train = int(min_length), int(max_length), is_train

    def src_length(self, idx: int) -> Optional[int]:
        """Frames the model will see for item ``idx`` (None = dropped)."""
        if self.trg_len is not None and self.trg_len[idx] < 0:
            return None
        return kept_length(self.n_frames[idx], self.min_length, self.max_length, self.is_train)

    def __iter__(self) -> Iterator[List[int]]:
        batch: List[int] = []
        max_tokens = 0
        for idx in self.sampler:
            idx = int(idx)
            src_len = self.src_length(idx)
            if src_len is None:  # otherwise drop instance
                continue
            batch.append(idx)
            if self.batch_type == "sentence":
                full = len(batch) >= self.batch_size
            else:
                trg_len = 0 if self.trg_len is None else int(self.trg_len[idx])
                n_tokens = 0 if src_len == 0 else max(src_len + 1, trg_len + 1)
                max_tokens = max(max_tokens, n_tokens)
                full = max_tokens * len(batch) >= self.batch_size
            if full:
                yield batch
                batch, max_tokens = [], 0
        if len(batch) > 0 and not self.drop_last:
            yield batch

    @property
    def num_samples(self) -> int:
        """``len(sampler)`` — UNFILTERED, like the reference's ``SentenceBatchSampler.num_samples``
        (datasets.py:1180-1192): schedulers and epoch-length computations built on ``len(batch_sampler)``
        must see the same number with either sampler."""
        try:
            return len(self.sampler)
        except (NotImplementedError, TypeError):
            return len(self.n_frames)

    def __len__(self) -> int:
        if self.batch_type == "token":
            raise NotImplementedError  # like TokenBatchSampler.__len__ (datasets.py:1294-1295)
        n = self.num_samples  # datasets.py:1213-1218: items the filters will drop are counted too
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


class SpeechBatchCollator:
    """Index list → ``(src, src_length, kept_indices)`` with ``src`` on the GPU.

    :param processor: :class:`joeys2t_b200.speech_processor.SpeechProcessor` (length filters, CMVN,
        SpecAugment configuration — the reference's ``tokenizer["src"]``)
    :param load_fn: ``idx -> waveform`` (int16 PCM or float in [-1, 1), shape (N,) or (C, N))
    :param is_train: training split (drops over-long items, applies SpecAugment)
    :param pad_index: padding value of ``pad_features`` (``float(pad_index)``)
    """

    def __init__(self, processor, load_fn: Callable[[int], "np.ndarray"], is_train: bool = False,
                 pad_index: int = 1):
        self.processor = processor
        self.load_fn = load_fn
        self.is_train = is_train
        self.pad_index = pad_index

    def __call__(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
        waves = [self.load_fn(int(i)) for i in indices]
        src, lengths, keep = self.processor.process_batch(
            waves, is_train=self.is_train, layout="padded", pad_index=self.pad_index)
        kept = [int(i) for i, k in zip(indices, keep) if k]
        if src is None:
            raise ValueError(f"every item of the batch {list(indices)} was filtered out")
        # pinned + non-blocking: a pageable host-to-device copy would wait for everything enqueued so far
        # (the batch's own H2D and kernels) and serialise the caller with the GPU.  The pinned memory comes
        # from a small ring owned by the collator (a fresh pin_memory() per batch costs more than the batch's
        # kernel launches); a ring entry is rewritten only after the copy that read it has completed.
        host_len, ev = self._length_slot(len(lengths))
        host_len[:len(lengths)] = torch.as_tensor(lengths, dtype=torch.long)
        src_length = host_len[:len(lengths)].to(src.device, non_blocking=True)
        ev.record(torch.cuda.current_stream(src.device))
        return src, src_length, kept

    def _length_slot(self, n: int):
        ring = self.__dict__.setdefault("_len_ring", [])
        if len(ring) < 8:
            ring.append([torch.empty(max(n, 1024), dtype=torch.long).pin_memory(), torch.cuda.Event()])
            self._len_next = len(ring) - 1
        else:
            self._len_next = (self._len_next + 1) % len(ring)
        slot = ring[self._len_next]
        slot[1].synchronize()  # (never recorded: returns at once)
        if slot[0].numel() < n:
            slot[0] = torch.empty(2 * n, dtype=torch.long).pin_memory()
        return slot[0], slot[1]


def lengths_from_dataset(dataset) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """``(n_frames, trg_len)`` per index for a reference ``SpeechDataset`` (duck-typed: ``df`` with an
    ``n_frames`` column, ``has_trg`` / ``has_prompt`` and ``get_item(idx, lang)`` which tokenises text
    only — no audio is touched).  See INTEGRATION.md."""
    n_frames = np.asarray(dataset.df["n_frames"].tolist(), dtype=np.int64)
    if not (dataset.has_trg or dataset.has_prompt.get("trg", False)):
        return n_frames, None
    trg_len = np.empty(len(n_frames), dtype=np.int64)
    for idx in range(len(n_frames)):
        trg = dataset.get_item(idx=idx, lang="trg")
        trg_len[idx] = -1 if trg is None else len(trg)
    return n_frames, trg_len

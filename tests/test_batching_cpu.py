# coding: utf-8
"""
Batched sampler path, CPU only: ``FrameCountBatchSampler`` must yield the index
lists the reference's ``SentenceBatchSampler`` / ``TokenBatchSampler`` yield — pinned to golden
batches produced by the unmodified reference (``oracle/make_golden_batches.py`` →
``tests/golden/ref_batches.npz``: ``load_data`` → ``make_iter`` on the 10 speech fixtures,
configuration of ``test/unit/test_data.py:185-214``) — without computing any features.
"""
import numpy as np
import pytest

from joeys2t_b200.batching import FrameCountBatchSampler, kept_length

GOLD_CASES = [(split, bt, bs) for split in ("train", "test")
              for bt, bs in (("sentence", 2), ("sentence", 3), ("token", 600), ("token", 1500))]


@pytest.fixture(scope="module")
def ref_batches():
    from tests.conftest import GOLD
    return np.load(GOLD / "ref_batches.npz")


def _split_batches(flat, sizes):
    out, o = [], 0
    for n in sizes:
        out.append([int(i) for i in flat[o:o + n]])
        o += n
    return out


@pytest.mark.parametrize("split,batch_type,batch_size", GOLD_CASES)
def test_index_batches_match_reference_golden(ref_batches, fixtures_pcm, split, batch_type, batch_size):
    z = ref_batches
    key = f"{split}_{batch_type}{batch_size}"
    _, n_frames = fixtures_pcm  # TSV column n_frames of test/data/speech/test.tsv
    assert np.array_equal(z["tsv_n_frames"], n_frames)
    sampler = FrameCountBatchSampler(
        z[f"{key}_order"].tolist(), batch_size, batch_type, n_frames=n_frames,
        trg_len=z[f"{split}_item_trg_len"], max_length=500, is_train=split == "train")
    got = [b for b in sampler]  # (list() would call __len__, which token batching does not define)
    want = _split_batches(z[f"{key}_indices"], z[f"{key}_batches"])
    assert got == want
    # the lengths the sampler assumed are the lengths of the features the reference then produced
    lens = [sampler.src_length(i) for b in got for i in b]
    assert lens == z[f"{key}_lengths"].tolist()
    # and the padded batch shape follows from them (test_data.py:251,270: (2, 310, 80), (2, 500, 80))
    for b, shape in zip(got, z[f"{key}_shapes"]):
        assert (len(b), max(sampler.src_length(i) for i in b), 80) == tuple(shape)
    if batch_type == "sentence":
        # like the reference (datasets.py:1180-1192, 1213-1218): from the UNFILTERED sampler length, so it
        # can exceed the number of batches actually yielded when the length filters drop items
        n_visited = len(z[f"{key}_order"])
        assert len(sampler) == -(-n_visited // batch_size) >= len(want)
    else:
        with pytest.raises(NotImplementedError):
            len(sampler)


def test_dropped_items_match_reference(ref_batches, fixtures_pcm):
    z = ref_batches
    _, n_frames = fixtures_pcm
    for split in ("train", "test"):
        s = FrameCountBatchSampler(range(10), 2, n_frames=n_frames, trg_len=z[f"{split}_item_trg_len"],
                                   max_length=500, is_train=split == "train")
        dropped = [s.src_length(i) is None for i in range(10)]
        assert dropped == z[f"{split}_item_dropped"].tolist()
        frames = [-1 if s.src_length(i) is None else s.src_length(i) for i in range(10)]
        assert frames == z[f"{split}_item_frames"].tolist()


def test_length_rule_edges():
    # tokenizers.py:473-484,496-500: strict inequalities, 0 frames never "too short"
    assert kept_length(199, min_length=200) is None
    assert kept_length(200, min_length=200) == 200
    assert kept_length(0, min_length=200) == 0
    assert kept_length(500, max_length=500, is_train=True) == 500
    assert kept_length(501, max_length=500, is_train=True) is None
    assert kept_length(501, max_length=500, is_train=False) == 500
    assert kept_length(10**6) == 10**6  # max_length = -1: no limit


def test_drop_last_and_empty_and_no_targets():
    n = [100, 120, 90, 300, 80]
    s = FrameCountBatchSampler(range(5), 2, n_frames=n, drop_last=True)
    assert list(s) == [[0, 1], [2, 3]] and len(s) == 2
    s = FrameCountBatchSampler(range(5), 2, n_frames=n)
    assert list(s) == [[0, 1], [2, 3], [4]] and len(s) == 3
    assert list(FrameCountBatchSampler([], 2, n_frames=n)) == []
    # token batches without targets: n_tokens = src_len + 1
    s = FrameCountBatchSampler(range(5), 250, "token", n_frames=n)
    assert [b for b in s] == [[0, 1, 2], [3], [4]]
    with pytest.raises(ValueError):
        FrameCountBatchSampler(range(5), 2, "bucket", n_frames=n)

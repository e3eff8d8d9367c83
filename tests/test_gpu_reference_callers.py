# coding: utf-8
"""
The drop-in under the reference's OWN callers, on the GPU (north_star: "datasets.SpeechDataset,
hub_interface.generate and the training/prediction loops pick it up as a drop-in").

``joeys2t_b200.install()`` is patched into the *unmodified* reference package — ``oracle/_ref``, installed
by ``oracle/build_ref.sh`` (byte-identical to ``/root/reference/joeynmt``; ``/root/reference`` itself does
not exist on the GPU box) — and the reference's own code then drives the B200 front-end:

* ``joeynmt.data.load_data`` -> ``build_tokenizer`` (``tokenizers.py:611-619``: now hands out the B200
  ``SpeechProcessor``) -> ``SpeechDataset.__getitem__`` (``datasets.py:636-656``) -> ``make_iter`` with
  ``SentenceBatchSampler`` / ``TokenBatchSampler`` (``datasets.py:1194-1292``, which call ``d[idx]`` per
  item) -> ``collate_fn`` (``datasets.py:186-242``: ``pad_features``) -> ``Batch``;
* ``SpeechStreamDataset.set_item`` / ``__getitem__`` (``datasets.py:792-863``), the dataset
  ``hub_interface.generate`` fills and ``predict`` iterates (``hub_interface.py:146-221``).

Everything is compared with ``tests/golden/ref_batches.npz`` / ``ref_fbank.npz``, which the same
reference code produced on the CPU (``oracle/make_golden_batches.py``): batch composition, shapes
and lengths exactly, features within the log-mel tolerance.
"""
import importlib
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import ref_shims

from .conftest import GOLD

pytestmark = pytest.mark.gpu

LOGMEL_ATOL = 1e-3  # BASELINE.json north_star: log-mel max-abs tolerance vs the CPU reference


@pytest.fixture(scope="module")
def patched_reference():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not ref_shims.reference_available():
        pytest.fail("the reference is neither mounted nor installed: run `bash oracle/build_ref.sh` in the "
                    "build container (oracle/_ref travels to the GPU box)")
    ref_shims.install(full_stack=True)
    names = ("joeynmt.helpers_for_audio", "joeynmt.data_augmentation", "joeynmt.tokenizers")
    saved = {n: dict(vars(importlib.import_module(n))) for n in names}
    import joeys2t_b200
    joeys2t_b200.install()
    yield ref_shims.speech_fixture_dir()
    for n in names:  # un-patch: later tests in this process see the pristine reference again
        mod = sys.modules[n]
        for k, v in saved[n].items():
            setattr(mod, k, v)


def _data_cfg(speech, **src_extra):
    src = {"lang": "en", "level": "frame", "num_freq": 80, "max_length": 500, "tokenizer_type": "speech"}
    src.update(src_extra)
    return {
        "train": str(speech / "test"), "test": str(speech / "test"),
        "src": src,
        "trg": {"lang": "en", "level": "char", "lowercase": True, "max_length": 50,
                "voc_file": str(speech / "char.txt")},
        "dataset_type": "speech",
        "special_symbols": SimpleNamespace(**{
            "unk_token": "<unk>", "pad_token": "<pad>", "bos_token": "<s>", "eos_token": "</s>",
            "sep_token": None, "unk_id": 0, "pad_id": 1, "bos_id": 2, "eos_id": 3, "sep_id": None,
            "lang_tags": []}),
    }


def test_reference_builder_hands_out_the_b200_processor(patched_reference):
    from joeynmt.data import load_data
    from joeys2t_b200.speech_processor import SpeechProcessor
    _, _, train_data, _, _ = load_data(_data_cfg(patched_reference), datasets=["train"], task="S2T")
    assert type(train_data.tokenizer["src"]) is SpeechProcessor
    assert train_data.tokenizer["src"].root_path == patched_reference


def test_speech_dataset_items_and_batches_match_the_reference(patched_reference, ref_fbank):
    """SpeechDataset.__getitem__ + make_iter (sentence and token batching, train and test split) running
    on the B200 front-end reproduce the reference's own batches (ref_batches.npz)."""
    import torch
    from joeynmt.data import load_data
    gold = np.load(GOLD / "ref_batches.npz")
    _, trg_vocab, train_data, _, test_data = load_data(_data_cfg(patched_reference),
                                                       datasets=["train", "test"], task="S2T")
    # per item: what SpeechDataset.__getitem__ hands to the samplers (datasets.py:636-656)
    for split, data in (("train", train_data), ("test", test_data)):
        for idx in range(len(data)):
            _, src, trg = data[idx]
            want_t = int(gold[f"{split}_item_frames"][idx])
            assert (src is None) == bool(gold[f"{split}_item_dropped"][idx]), (split, idx)
            if src is None:
                continue
            assert src.dtype == np.float32 and src.shape == (want_t, 80), (split, idx, src.shape)
            assert len(trg) == int(gold[f"{split}_item_trg_len"][idx])
            # test split truncates to max_length=500 (tokenizers.py:477-484); train drops instead
            assert np.abs(src - ref_fbank[idx][:want_t]).max() <= LOGMEL_ATOL, (split, idx)

    seed = 42
    for split, data, shuffle in (("train", train_data, True), ("test", test_data, False)):
        for batch_type, batch_size in (("sentence", 2), ("sentence", 3), ("token", 600), ("token", 1500)):
            key = f"{split}_{batch_type}{batch_size}"
            loader = data.make_iter(batch_size=batch_size, batch_type=batch_type, shuffle=shuffle, seed=seed,
                                    pad_index=trg_vocab.pad_index, eos_index=trg_vocab.eos_index,
                                    device=torch.device("cpu"), num_workers=0)
            loader.batch_sampler.set_seed(seed)
            index_batches = [list(b) for b in loader.batch_sampler]
            assert [len(b) for b in index_batches] == gold[f"{key}_batches"].tolist(), key
            assert [i for b in index_batches for i in b] == gold[f"{key}_indices"].tolist(), key
            loader.batch_sampler.set_seed(seed)
            np.random.seed(seed)
            shapes, lens = [], []
            for bi, batch in enumerate(loader):
                src = batch.src.numpy()
                assert src.dtype == np.float32
                assert batch.indices.tolist() == index_batches[bi]
                shapes.append(list(src.shape))
                lens += batch.src_length.tolist()
                # padding rows hold float(pad_index) exactly (helpers_for_audio.py:151-152)
                for row, t in enumerate(batch.src_length.tolist()):
                    assert (src[row, t:] == float(trg_vocab.pad_index)).all()
                    want = ref_fbank[index_batches[bi][row]][:t]
                    assert np.abs(src[row, :t] - want).max() <= LOGMEL_ATOL
                if f"{key}_full2" in gold and bi == 2:
                    full = gold[f"{key}_full2"]  # the batch test/unit/test_data.py:251,270 asserts the shape of
                    assert src.shape == full.shape
                    assert np.abs(src - full).max() <= LOGMEL_ATOL
            assert shapes == gold[f"{key}_shapes"].tolist(), key
            assert lens == gold[f"{key}_lengths"].tolist(), key


def test_speech_dataset_with_cmvn_and_specaugment(patched_reference, ref_processor):
    """The reference's dataset with CMVN + SpecAugment configured (configs/mustc_st.yaml:23-32 shape of
    config): items equal the goldens of the reference's own SpeechProcessor on the same RNG seed."""
    from joeynmt.data import load_data
    sa = dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0)
    for vname, before in (("before", True), ("after", False)):
        # mustc_st.yaml:21-32 nests both under src.tokenizer_cfg (tokenizers.py:559, :611-619)
        cfg = _data_cfg(patched_reference, min_length=200, tokenizer_cfg=dict(
            specaugment=sa, cmvn=dict(norm_means=True, norm_vars=True, before=before)))
        _, _, train_data, _, test_data = load_data(cfg, datasets=["train", "test"], task="S2T")
        ids = test_data.df["id"].tolist()
        for data, tag in ((train_data, "train"), (test_data, "eval")):
            rows = data.df["id"].tolist()
            for idx, uid in enumerate(rows):
                clip = int(uid.rsplit("-", 1)[1])
                key = f"{vname}_{tag}_c{clip}"
                np.random.seed(1000 + clip)
                _, src, trg = data[idx]
                if trg is None:  # target over trg.max_length: the dataset drops the pair (datasets.py:652-654)
                    assert src is None
                    continue
                assert (src is None) == bool(ref_processor[key + "_none"]), key
                if src is None:
                    continue
                assert list(src.shape) == ref_processor[key + "_shape"].tolist(), key
                if key + "_full" in ref_processor:
                    ref = ref_processor[key + "_full"]
                    # CMVN after SpecAugment: a frequency-masked column is constant, and the reference divides
                    # its own rounding noise by std = 1e-5 there (ill-conditioned, as in test_gpu_parity.py)
                    ok = np.ones(80, bool) if before else np.ptp(ref, axis=0) >= 1e-2
                    assert (np.abs(src - ref) <= 5e-4 + 1e-4 * np.abs(ref))[:, ok].all(), key
                    assert np.abs(src[:, ~ok]).max(initial=0) < 0.2, key
        assert ids


def test_speech_stream_dataset_the_hub_generate_route(patched_reference, ref_fbank):
    """hub_interface.generate (hub_interface.py:146-221) fills a SpeechStreamDataset with absolute wav
    paths and predict() iterates it with sentence batching: same route, no model."""
    import torch
    from joeynmt.data import load_data
    from joeynmt.datasets import SpeechStreamDataset
    # the dataset hub_interface builds: load_data(..., datasets=["stream"]) -> build_dataset("speech_stream")
    # (data.py:165-178, datasets.py:1128-1142)
    _, _, _, _, data = load_data(_data_cfg(patched_reference, min_length=10), datasets=["stream"], task="S2T")
    assert isinstance(data, SpeechStreamDataset)
    clips = [3, 0, 7]
    for c in clips:
        data.set_item(str(patched_reference / "wav" / f"260-123440-{c}.wav"))
    assert len(data) == len(clips)
    for i, c in enumerate(clips):  # SpeechStreamDataset.__getitem__ (datasets.py:852-863)
        _, src, _ = data[i]
        want = ref_fbank[c][:500]
        assert src.shape == want.shape and np.abs(src - want).max() <= LOGMEL_ATOL
    loader = data.make_iter(batch_size=len(clips), batch_type="sentence", shuffle=False, pad_index=1,
                            eos_index=3, device=torch.device("cpu"), num_workers=0)
    batches = list(loader)
    assert len(batches) == 1
    b = batches[0]
    # Batch sorts by source length (batch.py) — map rows back through batch.indices
    for row, i in enumerate(b.indices.tolist()):
        t = int(b.src_length[row])
        want = ref_fbank[clips[i]][:500]
        assert t == want.shape[0]
        assert np.abs(b.src[row, :t].numpy() - want).max() <= LOGMEL_ATOL
        assert (b.src[row, t:] == 1.0).all()
    data.reset_cache()
    assert len(data) == 0

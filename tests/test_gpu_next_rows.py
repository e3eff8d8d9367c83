# coding: utf-8
"""
GPU parity of the rows either side of the hot path (SURVEY.md §8 f-2 … f-4), ``-m gpu`` only:

* f-2  ``SpeechBatchCollator`` against the reference's own batches (``make_iter`` → ``collate_fn``:
       ``pad_features`` + ``torch.tensor(src).float()``), golden ``tests/golden/ref_batches.npz``;
* f-3  ``feature_store.extract_corpus``: batched GPU extraction into the reference's npy-in-zip format,
       read back through ``get_features("name.zip:offset:size")``;
* f-4  ``reformat_freq`` (48 kHz → 16 kHz ingest) bit-exact against the numpy expression of
       ``scripts/gradio_demo.py:35-45``.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import fbank_numpy as O  # noqa: E402

LOGMEL_ATOL = 1e-3


@pytest.fixture(scope="module")
def fe():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    from joeys2t_b200 import _lib, frontend
    if _lib.is_stale():
        _lib.build()
    return frontend


@pytest.fixture(scope="module")
def ref_batches():
    from tests.conftest import GOLD
    return np.load(GOLD / "ref_batches.npz")


# ---- f-2 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("split,batch_type,batch_size", [
    ("train", "sentence", 2), ("train", "token", 600), ("test", "sentence", 2), ("test", "sentence", 3),
    ("test", "token", 1500)])
def test_collated_batches_match_reference(fe, fixtures_pcm, ref_fbank, ref_batches, split, batch_type, batch_size):
    from joeys2t_b200.batching import FrameCountBatchSampler, SpeechBatchCollator
    from joeys2t_b200.speech_processor import SpeechProcessor
    z = ref_batches
    pcm, n_frames = fixtures_pcm
    key = f"{split}_{batch_type}{batch_size}"
    proc = SpeechProcessor(level="frame", num_freq=80, max_length=500)  # test_data.py:185-214
    sampler = FrameCountBatchSampler(
        z[f"{key}_order"].tolist(), batch_size, batch_type, n_frames=n_frames,
        trg_len=z[f"{split}_item_trg_len"], max_length=500, is_train=split == "train")
    collate = SpeechBatchCollator(proc, lambda i: pcm[i], is_train=split == "train", pad_index=1)
    lengths_ref = z[f"{key}_lengths"].tolist()
    o = 0
    for bi, indices in enumerate(sampler):
        src, src_length, kept = collate(indices)
        assert kept == indices and src.is_cuda and src.dtype == torch.float32
        assert tuple(src.shape) == tuple(z[f"{key}_shapes"][bi])  # test_data.py:251,270
        assert src_length.tolist() == lengths_ref[o:o + len(indices)]
        o += len(indices)
        got = src.cpu().numpy()
        # rows of every item: the reference's features (raw log-mel, truncated at 500 in eval)
        for j, i in enumerate(indices):
            t = int(src_length[j])
            ref = ref_fbank[i][:t]  # the reference's own features of this clip (golden)
            assert np.abs(got[j, :t] - ref).max() <= LOGMEL_ATOL
            assert (got[j, t:] == 1.0).all()  # pad_features fills with float(pad_index)
        if f"{key}_full2" in z.files and bi == 2:
            full = z[f"{key}_full2"]  # the reference's own collated tensor
            assert full.shape == got.shape
            assert np.abs(got - full).max() <= LOGMEL_ATOL
            assert ((full == 1.0) == (got == 1.0)).all()


# ---- f-3 ----------------------------------------------------------------------------------------
def test_extract_corpus_into_zip_store(fe, fixtures_pcm, ref_fbank, tmp_path):
    from joeys2t_b200 import feature_store as FS
    from joeys2t_b200 import helpers_for_audio as HA
    pcm, n_frames = fixtures_pcm
    items = [(f"260-123440-{i}", x.astype(np.float32) / np.float32(32768.0)) for i, x in enumerate(pcm)]
    items.insert(3, ("too-short", np.zeros(399, np.float32)))  # prepare_librispeech.py:86-88: reported, n_frames 0
    zip_path = tmp_path / "fbank80.zip"
    manifest, frames, failed = FS.extract_corpus(items, zip_path, batch_utterances=4)
    assert [f[0] for f in failed] == ["too-short"] and frames["too-short"] == 0
    assert set(manifest) == {f"260-123440-{i}" for i in range(10)}
    assert FS.get_zip_manifest(zip_path) == manifest
    for i in range(10):
        uid = f"260-123440-{i}"
        got = HA.get_features(tmp_path, manifest[uid])
        assert got.dtype == np.float32 and got.shape == (n_frames[i], 80) and frames[uid] == n_frames[i]
        assert np.abs(got - ref_fbank[i]).max() <= LOGMEL_ATOL
    # stored features feed the CMVN / SpecAugment path like the reference's zip branch does
    from joeys2t_b200.speech_processor import SpeechProcessor
    proc = SpeechProcessor(level="frame", num_freq=80, cmvn=dict(norm_means=True, norm_vars=True, before=True))
    proc.root_path = tmp_path
    y = proc(manifest["260-123440-1"], is_train=False)
    ref = O.cmvn(ref_fbank[1])
    assert np.abs(y - ref).max() <= 5e-4 + 1e-4 * np.abs(ref).max()


# ---- f-4 ----------------------------------------------------------------------------------------
def _ingest_signals():
    rng = np.random.default_rng(5)
    for n in (3, 24, 3 * 8 * 256 + 3, 3 * 100003):
        t = np.arange(n)
        yield f"speech-like int16 n={n}", (3000 * np.sin(t / 37.0) * rng.random(n)).astype(np.int16)
        yield f"full-scale int16 n={n}", rng.integers(-32768, 32768, n).astype(np.int16)
        yield f"float32 int16-range n={n}", (rng.random(n) * 65535 - 32768).astype(np.float32)
    n = 3 * 4001
    t = np.arange(n)
    yield "negative peak larger than positive (overflowing cast)", np.where(t % 7 == 0, -3000, 1000).astype(np.int16)
    yield "all non-positive (max -> 1)", (-np.abs(rng.integers(0, 200, n))).astype(np.int16)
    yield "zeros", np.zeros(n, np.int16)
    yield "float32 in [-1, 1) (divisor 1)", (rng.random(n) * 2 - 1).astype(np.float32)


def test_reformat_freq_bit_exact(fe):
    from joeys2t_b200 import helpers_for_audio as HA
    for name, y in _ingest_signals():
        with np.errstate(all="ignore"):
            ref, sr_ref = O.reformat_freq(48000, y.copy())
        got, sr = HA.reformat_freq(48000, y)
        assert sr == sr_ref == 16000 and got.dtype == np.int16 and got.shape == ref.shape, name
        assert np.array_equal(got, ref), f"{name}: {(got != ref).sum()} of {ref.size} samples differ"
    y = np.arange(30, dtype=np.int16)
    same, sr = HA.reformat_freq(16000, y)
    assert same is y and sr == 16000
    with pytest.raises(ValueError):
        HA.reformat_freq(44100, y)
    with pytest.raises(ValueError):
        HA.reformat_freq(48000, np.zeros(31, np.int16))
    with pytest.raises(ValueError):
        HA.reformat_freq(48000, np.zeros(30, np.float64))


def test_reformat_then_fbank_pipeline(fe, fixtures_pcm):
    """The demo's front door (gradio_demo.py:48-55): 48 kHz microphone PCM → reformat_freq →
    extract_fbank_features, against the same chain on the CPU oracle."""
    from joeys2t_b200 import helpers_for_audio as HA
    pcm, _ = fixtures_pcm
    y48 = np.repeat(pcm[1], 3)  # a 48 kHz signal whose block means are the 16 kHz fixture
    got16, sr = HA.reformat_freq(48000, y48)
    ref16, _ = O.reformat_freq(48000, y48)
    assert np.array_equal(got16, ref16)
    feats = HA.extract_fbank_features(torch.tensor(got16[np.newaxis, :]).float() / 32768.0, sr)
    ref = O.extract_fbank_features(ref16)
    assert np.abs(feats - ref).max() <= LOGMEL_ATOL

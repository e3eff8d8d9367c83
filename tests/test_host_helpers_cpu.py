# coding: utf-8
"""
Host-side helpers of the path against outputs of the unmodified reference (``tests/golden/ref_misc.npz``,
``ref_tables.npz``; generators ``oracle/make_golden_misc.py``, ``oracle/make_golden.py``).  CPU only: these
functions never touch the device.
"""
import numpy as np
import pytest

from oracle import make_golden_misc as G

from .conftest import GOLD


@pytest.fixture(scope="module")
def ref_misc():
    return np.load(GOLD / "ref_misc.npz")


def test_get_n_frames_equals_the_reference_over_a_length_sweep(ref_misc):
    """helpers_for_audio.py:93-96.  The float expression is part of the contract: int(N / sr * 1000) and
    int(1000 * N / sr) differ for 120 of these lengths (16080 samples at 16 kHz: 98 vs 99 frames)."""
    from joeys2t_b200.helpers_for_audio import get_n_frames
    lengths = G.n_frames_lengths()
    for sr in G.RATES:
        ours = np.array([get_n_frames(int(n), sr) for n in lengths], np.int32)
        assert np.array_equal(ours, ref_misc[f"n_frames_sr{sr}"]), sr
    assert get_n_frames(16080, 16000) == 98 and get_n_frames(64240, 16000) == 399


def test_pad_features_equals_the_reference(ref_misc):
    """helpers_for_audio.py:130-170 on seeded ragged lists (default and explicit embed_size / pad_index)."""
    from joeys2t_b200.helpers_for_audio import pad_features
    for seed, embed, pad in G.pad_cases():
        feats, lens, third = pad_features(G.pad_input(seed, embed), embed_size=embed, pad_index=pad)
        ref = ref_misc[f"pad{seed}_features"]
        assert third is None and isinstance(lens, list)
        assert feats.dtype == np.float32 and feats.shape == ref.shape
        assert np.array_equal(feats, ref), seed
        assert lens == ref_misc[f"pad{seed}_lengths"].tolist()
    feats, lens, _ = pad_features(G.pad_input(0, 80))  # defaults: embed_size=80, pad_index=1
    assert np.array_equal(feats, ref_misc["pad0_features"])
    with pytest.raises(AssertionError):
        pad_features([np.zeros((0, 80), np.float32)])


def test_tables_are_bit_identical_to_torchaudio(ref_tables):
    """povey window (kaldi.py:98-100) and 80x256 mel bank (kaldi.py:436-511) as torchaudio builds them."""
    from joeys2t_b200 import tables
    win, mel = tables.povey_window(), tables.mel_banks()
    assert win.dtype == np.float32 and mel.dtype == np.float32
    assert np.array_equal(win.view(np.uint32), ref_tables["povey400"].view(np.uint32))
    assert np.array_equal(mel.view(np.uint32), ref_tables["mel80x256"].view(np.uint32))
    assert win[0] == 0.0 and win[399] == 0.0
    assert int((mel != 0).sum()) == 501 and mel[:, 0].max() == 0.0


@pytest.mark.parametrize("cfg", [
    dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),   # mustc_st.yaml:23-32
    dict(freq_mask_n=1, freq_mask_f=80, time_mask_n=3, time_mask_t=40, time_mask_p=0.2),
    dict(freq_mask_n=0, freq_mask_f=5, time_mask_n=2, time_mask_t=1, time_mask_p=1.0),      # randint(0, 1): no draw
    dict(freq_mask_n=3, freq_mask_f=1, time_mask_n=0, time_mask_t=10, time_mask_p=1.0),
    dict(freq_mask_n=2, freq_mask_f=33, time_mask_n=2, time_mask_t=100000, time_mask_p=0.0),  # max_t = 0
])
def test_batched_mask_draws_replay_numpy_randint_exactly(cfg):
    """``js2t_specaug_replay`` (host C code behind ``mask_tables_for_batch``) must give the values AND leave the
    global ``np.random`` stream where the reference's per-item ``np.random.randint`` calls leave it
    (joeynmt/data_augmentation.py:48-70) — compared with the per-item route that replays them call by call."""
    from joeys2t_b200 import data_augmentation as D

    sa = D.SpecAugment(**cfg)
    n_frames = [1, 2, 3, 0, 17, 98, 1000, 1499, 0, 65536, 400000, 5, 31, 32, 33, 4097]
    for seed in (0, 1, 1234):
        np.random.seed(seed)
        np.random.randint(0, 10, 3)  # (the stream need not be at a fresh seed)
        fast, nf, nt = D.mask_tables_for_batch(sa, n_frames)
        assert D._replay_tables(sa, np.array([5, 7], np.int32), 80, nf + nt) is not None  # the C route is live
        np.random.seed(seed)
        np.random.randint(0, 10, 3)
        fast, nf, nt = D.mask_tables_for_batch(sa, n_frames)
        state_fast = np.random.get_state()
        follow_fast = np.random.randint(0, 1 << 30, 4)
        np.random.seed(seed)
        np.random.randint(0, 10, 3)
        D._FAST_DRAWS = False
        try:
            slow, _, _ = D.mask_tables_for_batch(sa, n_frames)
        finally:
            D._FAST_DRAWS = True
        state_slow = np.random.get_state()
        follow_slow = np.random.randint(0, 1 << 30, 4)
        assert (nf, nt) == (cfg["freq_mask_n"], cfg["time_mask_n"])
        assert fast.dtype == slow.dtype == np.int32 and fast.shape == slow.shape
        assert np.array_equal(fast, slow)
        assert state_fast[0] == state_slow[0] and np.array_equal(state_fast[1], state_slow[1])
        assert state_fast[2:] == state_slow[2:]
        assert np.array_equal(follow_fast, follow_slow)

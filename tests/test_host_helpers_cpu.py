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

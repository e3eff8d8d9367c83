# coding: utf-8
"""
Pins the CPU oracle (``oracle/fbank_numpy.py``) against everything the reference's own tests hold
for this path and against golden vectors produced by the unmodified reference
(``oracle/make_golden.py``).  CPU only.
"""
import hashlib

import numpy as np
import pytest

from oracle import fbank_numpy as O

# the reference's single known-answer vector: test/unit/test_tokenizer.py:322-329
# (fbank + CMVN of 260-123440-1.wav, frame 0, mel bins 0..9, atol = rtol = 1e-5)
KNOWN_ANSWER = np.array([
    -1.0788909, -1.0076448, -1.0421542, -1.0393586, -1.0239305,
    -0.9921213, -0.95107234, -0.9340749, -0.9119267, -0.8962079,
], np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_known_answer_vector(fixtures_pcm):
    pcm, _ = fixtures_pcm
    feats = O.cmvn(O.extract_fbank_features(pcm[1]))
    np.testing.assert_allclose(feats[0, :10], KNOWN_ANSWER, atol=1e-5, rtol=1e-5)


def test_known_answer_vector_on_golden(ref_fbank):
    # the golden fbank itself reproduces the reference's vector through the reference CMVN formula
    np.testing.assert_allclose(O.cmvn(ref_fbank[1])[0, :10], KNOWN_ANSWER, atol=1e-5, rtol=1e-5)


def test_frame_counts_match_manifest(fixtures_pcm, ref_fbank):
    pcm, n_frames = fixtures_pcm
    for x, t, f in zip(pcm, n_frames, ref_fbank):
        assert O.num_frames(len(x)) == t == f.shape[0]
    assert [O.num_frames(n) for n in (399, 400, 559, 560, 720)] == [0, 1, 1, 2, 3]


def test_fbank_matches_reference_golden(fixtures_pcm, ref_fbank):
    pcm, _ = fixtures_pcm
    for i, (x, ref) in enumerate(zip(pcm, ref_fbank)):
        got = O.extract_fbank_features(x)
        assert got.dtype == np.float32 and got.shape == ref.shape
        # north_star tolerance: log-mel max-abs <= 1e-3 (observed <= 4e-4: fp32 FFT order noise)
        assert np.abs(got - ref).max() <= 1e-3, i


def test_fbank_float_input_equals_int16_input(fixtures_pcm):
    pcm, _ = fixtures_pcm
    a = O.extract_fbank_features(pcm[0])
    b = O.extract_fbank_features((pcm[0].astype(np.float32) / 32768.0)[None, :])
    assert np.array_equal(a, b)  # quirk Q4: /32768 then *2**15 is exact


def test_fbank_fp64_arbiter_close(fixtures_pcm, ref_fbank):
    pcm, _ = fixtures_pcm
    hi = O.fbank_kaldi(pcm[3].astype(np.float64), dtype=np.float64)
    assert np.abs(hi - ref_fbank[3]).max() <= 1e-3


def test_short_input_raises():
    with pytest.raises(ValueError):
        O.extract_fbank_features(np.zeros(399, np.int16))
    assert O.extract_fbank_features(np.zeros(400, np.int16)).shape == (1, 80)


def test_multichannel_takes_channel0(fixtures_pcm):
    pcm, _ = fixtures_pcm
    x = pcm[0][:8000]
    stereo = np.stack([x, x[::-1]])
    assert np.array_equal(O.extract_fbank_features(stereo), O.extract_fbank_features(x))


def test_tables_close_to_torchaudio(ref_tables):
    # numpy's float32 log/cos differ from torch's by an ulp; the oracle tables are within 2e-5
    assert np.abs(O.povey_window() - ref_tables["povey400"]).max() < 1e-6
    assert np.abs(O.mel_banks() - ref_tables["mel80x256"]).max() < 5e-5
    assert (ref_tables["mel80x256"] != 0).sum() == 501


@pytest.mark.parametrize("nm", [True, False])
@pytest.mark.parametrize("nv", [True, False])
def test_cmvn_matches_reference_golden(ref_fbank, ref_cmvn, nm, nv):
    tag = f"m{int(nm)}v{int(nv)}"
    for i, f in enumerate(ref_fbank):
        y = O.cmvn(f, nm, nv)
        assert y.dtype == np.float32
        assert np.array_equal(y[:4], ref_cmvn[f"{tag}_head{i}"])
        assert np.array_equal(y[-4:], ref_cmvn[f"{tag}_tail{i}"])
        s = np.array([y.astype(np.float64).sum(), (y.astype(np.float64)**2).sum()])
        np.testing.assert_allclose(s, ref_cmvn[f"{tag}_sum{i}"], rtol=1e-12, atol=1e-9)
    assert np.array_equal(O.cmvn(ref_fbank[1], nm, nv), ref_cmvn[f"{tag}_full1"])


def test_cmvn_silent_utterance(ref_cmvn):
    silent = ref_cmvn["silent_fbank"]
    assert np.all(silent == np.float32(np.log(np.float32(O.FLT_EPSILON))))
    got = O.cmvn(O.extract_fbank_features(np.zeros(4000, np.int16)))
    assert np.array_equal(got, ref_cmvn["silent_cmvn"])


def test_cmvn_fp64_variant_close(ref_fbank):
    for f in ref_fbank:
        assert np.abs(O.cmvn_fp64(f) - O.cmvn(f)).max() < 5e-4


def test_specaugment_matches_reference_golden(ref_fbank, ref_specaugment):
    g = ref_specaugment
    cfgs = {
        "mustc": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
        "test": dict(freq_mask_n=1, freq_mask_f=5, time_mask_n=1, time_mask_t=10, time_mask_p=1.0),
        "default": dict(),
        "smallp": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=0.05),
        "zerop": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=0.001),
        "widef": dict(freq_mask_n=2, freq_mask_f=81, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
        "const": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0,
                      mask_value=0.0),
    }
    shas = dict(zip(g["sha_keys"].tolist(), g["sha_vals"].tolist()))
    cm = [O.cmvn(f) for f in ref_fbank]
    for cname, cfg in cfgs.items():
        for seed in (0, 1, 2345):
            for i, x in enumerate(cm):
                key = f"{cname}_s{seed}_c{i}"
                np.random.seed(seed)
                y = O.specaugment(x, **cfg)
                assert sha(y) == shas[key], key  # bit-exact incl. fill value
                if i == 1:
                    assert np.array_equal(y, g[key + "_full"])
                assert np.array_equal(np.packbits(y != x), g[key + "_changed"])
    # degenerate branches really are exercised by the goldens
    assert int(g["widef_s0_c0_untouched"]) == 1
    assert g["zerop_s0_c0_tm"].shape[0] == 0 and g["zerop_s0_c0_fm"].shape[0] == 2


def test_speech_processor_order_and_filters(ref_fbank, ref_processor):
    g = ref_processor
    sa = dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0)
    for vname, before in (("before", True), ("after", False)):
        ccfg = dict(norm_means=True, norm_vars=True, before=before)
        for i, f in enumerate(ref_fbank):
            for is_train in (True, False):
                key = f"{vname}_{'train' if is_train else 'eval'}_c{i}"
                np.random.seed(1000 + i)
                y = O.speech_processor(f, is_train, min_length=200, max_length=500, cmvn_cfg=ccfg,
                                       specaug_cfg=sa)
                if int(g[key + "_none"]):
                    assert y is None, key
                    continue
                assert list(y.shape) == g[key + "_shape"].tolist(), key
                assert sha(y.astype(np.float32)) == str(g[key + "_sha"]), key
    # the shapes the reference asserts in test/unit/test_data.py:251,270:
    # train drops T > max_length (1470- and 1200-frame clips), eval truncates them to 500
    assert int(g["before_train_c2_none"]) == 1 and int(g["before_train_c4_none"]) == 1
    assert g["before_eval_c2_shape"].tolist() == [500, 80]
    assert g["before_eval_c4_shape"].tolist() == [500, 80]
    # min_length=200 drops the 172-frame clip everywhere (tokenizers.py:473-476)
    assert int(g["before_eval_c1_none"]) == 1


def test_pad_features_layout(ref_fbank):
    feats, lengths, _ = O.pad_features([ref_fbank[1], ref_fbank[0]])
    assert feats.shape == (2, 215, 80) and lengths == [172, 215]
    assert np.all(feats[0, 172:] == 1.0) and np.array_equal(feats[1], ref_fbank[0])


def test_global_cmvn_is_cmvn_of_concatenation(ref_fbank):
    got = O.global_cmvn(ref_fbank[:4])
    want = O.cmvn_fp64(np.concatenate(ref_fbank[:4], 0))
    np.testing.assert_allclose(np.concatenate(got, 0), want, rtol=0, atol=1e-6)


def test_oracle_vs_installed_torchaudio_on_synthetic():
    ta = pytest.importorskip("torchaudio.compliance.kaldi")
    import torch
    from joeys2t_b200 import synthetic
    for w in synthetic.librispeech_batch(2, seed=7, lo=1.0, hi=2.0):
        ref = ta.fbank(torch.from_numpy(w.astype(np.float32))[None], num_mel_bins=80,
                       sample_frequency=16000).numpy()
        assert np.abs(O.extract_fbank_features(w) - ref).max() <= 1e-3

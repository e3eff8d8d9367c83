# coding: utf-8
"""
The drop-in contract, CPU only: every symbol ``joeys2t_b200.install()`` swaps into an
importable ``joeynmt`` has the reference's own signature (names, order, defaults) and — for the classes
— the reference's ``__repr__`` text, which the reference logs (``tokenizers.py:502-508``).  Needs the
reference itself (``/root/reference``, through the import shims of ``oracle/ref_shims.py``).
"""
import importlib
import inspect
import sys

import pytest

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference is not mounted")


@pytest.fixture(scope="module")
def reference_modules():
    ref_shims.install(full_stack=True)
    originals = {}
    for name in ("joeynmt.helpers_for_audio", "joeynmt.data_augmentation", "joeynmt.tokenizers"):
        mod = importlib.import_module(name)
        originals[name] = {k: getattr(mod, k) for k in dir(mod) if not k.startswith("__")}
    return originals


def _params(fn):
    return [(p.name, p.kind, p.default) for p in inspect.signature(fn).parameters.values()]


def test_function_signatures_match_the_reference(reference_modules):
    from joeys2t_b200 import helpers_for_audio as ours
    ref = reference_modules["joeynmt.helpers_for_audio"]
    for name in ("extract_fbank_features", "_get_torchaudio_fbank", "get_features", "pad_features",
                 "get_n_frames", "_convert_to_mono", "_is_npy_data", "_get_features_from_zip"):
        assert _params(getattr(ours, name)) == _params(ref[name]), name


def test_class_signatures_and_repr_match_the_reference(reference_modules):
    from joeys2t_b200 import data_augmentation as da
    from joeys2t_b200.speech_processor import SpeechProcessor
    ref_da = reference_modules["joeynmt.data_augmentation"]
    ref_tk = reference_modules["joeynmt.tokenizers"]
    for cls_name, ours in (("CMVN", da.CMVN), ("SpecAugment", da.SpecAugment)):
        theirs = ref_da[cls_name]
        assert _params(ours.__init__) == _params(theirs.__init__), cls_name
        assert _params(ours.__call__) == _params(theirs.__call__), cls_name
        assert repr(ours()) == repr(theirs()), cls_name
    kw = dict(freq_mask_n=1, freq_mask_f=5, time_mask_n=3, time_mask_t=10, time_mask_p=0.5)
    assert repr(da.SpecAugment(**kw)) == repr(ref_da["SpecAugment"](**kw))
    assert repr(da.CMVN(False, True, False)) == repr(ref_da["CMVN"](False, True, False))
    assert da.CMVN(before=False).before is False  # read by tokenizers.py:488,492
    theirs = ref_tk["SpeechProcessor"]
    assert _params(SpeechProcessor.__init__) == _params(theirs.__init__)
    assert _params(SpeechProcessor.__call__) == _params(theirs.__call__)
    cfg = dict(level="frame", num_freq=80, normalize=False, max_length=500, min_length=10,
               specaugment=kw, cmvn=dict(norm_means=True, norm_vars=False, before=True))
    assert repr(SpeechProcessor(**cfg)) == repr(theirs(**cfg))


def test_install_swaps_every_symbol_and_restores_cleanly(reference_modules):
    import joeys2t_b200
    from joeys2t_b200 import data_augmentation as da
    from joeys2t_b200 import helpers_for_audio as ha
    from joeys2t_b200.speech_processor import SpeechProcessor
    jha, jda, jtk = joeys2t_b200.install()
    try:
        for name in ("get_features", "pad_features"):
            assert getattr(jha, name) is getattr(ha, name)
        for name in ("extract_fbank_features", "_get_torchaudio_fbank"):
            routed = getattr(jha, name)
            # the shipped geometry (16 kHz, 80 bins) goes to the B200 function; anything else keeps the
            # reference's own behaviour through the reference's original function
            assert routed.__wrapped_b200__ is getattr(ha, name)
            assert routed.__reference_original__ is reference_modules["joeynmt.helpers_for_audio"][name]
            assert _params(routed) == _params(routed.__reference_original__)
        assert jda.CMVN is da.CMVN and jda.SpecAugment is da.SpecAugment
        assert jtk.CMVN is da.CMVN and jtk.SpecAugment is da.SpecAugment
        assert jtk.SpeechProcessor is SpeechProcessor and jtk.get_features is ha.get_features
        # the reference's own builder now hands out the B200 processor (tokenizers.py:611-619)
        assert sys.modules["joeynmt.tokenizers"].SpeechProcessor is SpeechProcessor
        # 8 kHz audio (frame geometry 200 / 80 / 256) is outside the B200 path: the call must end up in
        # the reference's CPU implementation and return what the unpatched reference returns
        import torch
        torch.manual_seed(0)
        wave8k = torch.rand(1, 4000) * 2 - 1
        with pytest.warns(UserWarning, match="outside the B200 path"):
            got = jha.extract_fbank_features(wave8k, 8000)
        want = reference_modules["joeynmt.helpers_for_audio"]["extract_fbank_features"](wave8k, 8000)
        assert got.shape == want.shape == (1 + (4000 - 200) // 80, 80) and (got == want).all()
        with pytest.warns(UserWarning):
            assert jha.extract_fbank_features(torch.rand(1, 8000), 16000, n_mel_bins=40).shape[1] == 40
    finally:
        for modname, mod in (("joeynmt.helpers_for_audio", jha), ("joeynmt.data_augmentation", jda),
                             ("joeynmt.tokenizers", jtk)):
            for k, v in reference_modules[modname].items():
                setattr(mod, k, v)

# coding: utf-8
"""
joeys2t_b200 — B200-native (sm_100a CUDA) implementation of JoeyS2T's audio front-end hot path:
Kaldi-compatible 80-bin log-mel fbank → utterance / global CMVN → SpecAugment, behind the
reference's unchanged Python signatures.

    joeys2t_b200.helpers_for_audio   ↔ joeynmt/helpers_for_audio.py
    joeys2t_b200.data_augmentation   ↔ joeynmt/data_augmentation.py
    joeys2t_b200.speech_processor    ↔ joeynmt/tokenizers.py:433-508 (SpeechProcessor)
    joeys2t_b200.frontend            batched device API over the C ABI (include/joeys2t_b200.h)
    joeys2t_b200.distributed         utterance sharding + the global-CMVN all-reduce

:func:`install` patches an importable ``joeynmt`` in place so datasets.SpeechDataset,
hub_interface.generate and the training / prediction loops pick the GPU path up unchanged.
"""
__version__ = "0.1.0"


def install():
    """Monkey-patch ``joeynmt`` (must be importable) with the B200 front-end.  See INTEGRATION.md."""
    import importlib
    from joeys2t_b200 import data_augmentation as da
    from joeys2t_b200 import helpers_for_audio as ha
    from joeys2t_b200 import speech_processor as sp

    jha = importlib.import_module("joeynmt.helpers_for_audio")
    for name in ("extract_fbank_features", "_get_torchaudio_fbank", "get_features", "pad_features"):
        setattr(jha, name, getattr(ha, name))
    jda = importlib.import_module("joeynmt.data_augmentation")
    jda.CMVN, jda.SpecAugment = da.CMVN, da.SpecAugment
    jtk = importlib.import_module("joeynmt.tokenizers")
    jtk.CMVN, jtk.SpecAugment = da.CMVN, da.SpecAugment
    jtk.get_features = ha.get_features
    jtk.SpeechProcessor = sp.SpeechProcessor
    return jha, jda, jtk

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


def _route_by_geometry(ours, theirs, rate_arg: int, bins_arg: int, bins_kw: str):
    """The hot path is specialised for the reference's shipped geometry (16 kHz, 80 bins: every config
    and fixture).  After :func:`install` a call with ANOTHER sample rate or bin count keeps the reference's
    own behaviour — it is handed to the reference's original function, untouched — instead of raising;
    calls with the shipped geometry never leave the GPU path (there is no fallback for them)."""
    import functools
    import warnings
    from joeys2t_b200 import tables

    @functools.wraps(ours)
    def routed(*args, **kwargs):
        rate = kwargs.get("sample_rate", args[rate_arg] if len(args) > rate_arg else tables.SAMPLE_RATE)
        bins = kwargs.get(bins_kw, args[bins_arg] if len(args) > bins_arg else tables.NUM_MEL_BINS)
        if int(rate) == tables.SAMPLE_RATE and int(bins) == tables.NUM_MEL_BINS:
            return ours(*args, **kwargs)
        warnings.warn(f"joeys2t_b200: sample_rate={rate}, bins={bins} is outside the B200 path (16 kHz, 80 "
                      "bins); using the reference's own CPU implementation for this call", stacklevel=2)
        return theirs(*args, **kwargs)

    routed.__wrapped_b200__ = ours
    return routed


def install():
    """Monkey-patch ``joeynmt`` (must be importable) with the B200 front-end.  See INTEGRATION.md."""
    import importlib
    from joeys2t_b200 import data_augmentation as da
    from joeys2t_b200 import helpers_for_audio as ha
    from joeys2t_b200 import speech_processor as sp

    jha = importlib.import_module("joeynmt.helpers_for_audio")
    originals = {n: getattr(jha, n) for n in ("extract_fbank_features", "_get_torchaudio_fbank")}
    for name in ("get_features", "pad_features"):
        setattr(jha, name, getattr(ha, name))
    # (waveform, sample_rate, output_path, n_mel_bins, overwrite) / (waveform, sample_rate, n_bins)
    for name, bins_arg, bins_kw in (("extract_fbank_features", 3, "n_mel_bins"), ("_get_torchaudio_fbank", 2, "n_bins")):
        orig = originals[name]
        orig = getattr(orig, "__reference_original__", orig)  # install() twice: keep the real original
        routed = _route_by_geometry(getattr(ha, name), orig, 1, bins_arg, bins_kw)
        routed.__reference_original__ = orig
        setattr(jha, name, routed)
    jda = importlib.import_module("joeynmt.data_augmentation")
    jda.CMVN, jda.SpecAugment = da.CMVN, da.SpecAugment
    jtk = importlib.import_module("joeynmt.tokenizers")
    jtk.CMVN, jtk.SpecAugment = da.CMVN, da.SpecAugment
    jtk.get_features = ha.get_features
    jtk.SpeechProcessor = sp.SpeechProcessor
    return jha, jda, jtk

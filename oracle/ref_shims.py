# coding: utf-8
"""
TEST INFRASTRUCTURE ONLY — never imported by the product path.

Import shims that let the *unmodified* reference (``/root/reference/joeynmt``)
run on CPU in the build container, so that

* ``oracle/make_golden.py`` can generate golden vectors from the reference itself, and
* ``tests/test_oracle.py`` can pin the numpy restatement in ``oracle/`` against the real thing, and
* the installed copy ``oracle/_ref`` (``oracle/build_ref.sh``: ``pip install --no-deps --target``, byte-
  identical to ``/root/reference/joeynmt``) can run on the GPU box, where ``/root/reference`` does not
  exist: ``bench.py --impl reference`` / the ``cpu_baseline`` leg time the reference's own
  ``extract_fbank_features`` -> ``CMVN``, and ``tests/test_gpu_reference_callers.py`` drives the
  reference's own dataset classes with ``joeys2t_b200.install()`` patched in.

Shims:

* ``torchaudio.sox_effects`` was removed from torchaudio 2.11 but is imported at
  ``joeynmt/helpers_for_audio.py:13`` → stub module whose ``apply_effects_tensor`` down-mixes
  (its result is discarded by ``helpers_for_audio.py:53-54`` anyway).
* ``torchaudio.load`` (``helpers_for_audio.py:115``) needs the absent ``torchcodec`` →
  replaced by a stdlib ``wave`` loader: int16 → float32 / 32768, shape (1, N).
* name-only stubs for packages the dataset stack imports but the hot path never calls:
  ``sacrebleu``, ``subword_nmt``, ``matplotlib``, ``editdistance``, ``plotly``.
"""
import importlib
import os
import sys
import types
import wave
from pathlib import Path

import numpy as np

_SOURCE_ROOT = Path("/root/reference")                     # the build container
_INSTALLED_ROOT = Path(__file__).resolve().parent / "_ref"  # oracle/build_ref.sh (travels to the GPU box)


def _has_reference(root: Path) -> bool:
    return (root / "joeynmt" / "helpers_for_audio.py").is_file()


# where ``import joeynmt`` resolves to: the mounted source tree if present, else the installed copy
# (JS2T_USE_INSTALLED_REF=1 forces the installed copy, to rehearse the GPU-box situation here)
REFERENCE_ROOT = _INSTALLED_ROOT if _has_reference(_INSTALLED_ROOT) and (
    os.environ.get("JS2T_USE_INSTALLED_REF") or not _has_reference(_SOURCE_ROOT)) else _SOURCE_ROOT


def reference_available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def speech_fixture_dir() -> Path:
    """test/data/speech of the reference (ten wavs, TSVs, vocabulary)."""
    if REFERENCE_ROOT == _SOURCE_ROOT:
        return REFERENCE_ROOT / "test" / "data" / "speech"
    return REFERENCE_ROOT / "test_data" / "speech"


def load_wav_int16(path) -> np.ndarray:
    """stdlib reader for the mono int16 fixtures; returns (N,) int16."""
    with wave.open(str(path), "rb") as w:
        assert w.getsampwidth() == 2, "fixtures are 16-bit PCM"
        n_ch = w.getnchannels()
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        sr = w.getframerate()
    if n_ch > 1:
        pcm = pcm.reshape(-1, n_ch)[:, 0]
    return pcm.copy(), sr


def _wave_torchaudio_load(path, *_args, **_kwargs):
    import torch
    pcm, sr = load_wav_int16(path)
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0).unsqueeze(0), sr


def _stub(name: str, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
        if not hasattr(mod, "__file__"):  # one of ours: allow late attributes
            mod.__dict__.update(attrs)
        return mod
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, mod)
    return mod


def install(full_stack: bool = False):
    """Install the shims and put the reference on ``sys.path``.

    :param full_stack: also stub the text-side packages needed to import
        ``joeynmt.tokenizers`` / ``joeynmt.datasets`` / ``joeynmt.data``.
    :returns: the imported ``joeynmt.helpers_for_audio`` module of the reference.
    """
    if not reference_available():
        raise RuntimeError("neither /root/reference nor oracle/_ref (oracle/build_ref.sh) holds the reference")
    import torchaudio

    def _apply_effects_tensor(waveform, sample_rate, effects):
        return waveform.mean(0, keepdim=True), sample_rate

    try:
        importlib.import_module("torchaudio.sox_effects")
    except Exception:  # pylint: disable=broad-except
        _stub("torchaudio.sox_effects", apply_effects_tensor=_apply_effects_tensor)
    torchaudio.load = _wave_torchaudio_load

    if full_stack:
        for name in ("sacrebleu", "subword_nmt", "matplotlib", "editdistance", "plotly"):
            try:
                importlib.import_module(name)
            except Exception:  # pylint: disable=broad-except
                _stub(name)
        if not hasattr(sys.modules["sacrebleu"], "__file__"):
            _stub("sacrebleu.metrics")
            _stub("sacrebleu.metrics.bleu", _get_tokenizer=lambda *_a, **_k: None)
            sys.modules["sacrebleu"].__dict__.setdefault("corpus_bleu", None)
            sys.modules["sacrebleu"].__dict__.setdefault("corpus_chrf", None)
        if not hasattr(sys.modules["subword_nmt"], "__file__"):
            _stub("subword_nmt.apply_bpe", BPE=object)
        if not hasattr(sys.modules["matplotlib"], "__file__"):
            _stub("matplotlib", use=lambda *_a, **_k: None, rcParams={})
            _stub("matplotlib.pyplot")
            _stub("matplotlib.backends")
            _stub("matplotlib.backends.backend_pdf", PdfPages=object)
            _stub("matplotlib.figure", Figure=object)
        if not hasattr(sys.modules["editdistance"], "__file__"):
            sys.modules["editdistance"].__dict__.setdefault("eval", None)
        if not hasattr(sys.modules["plotly"], "__file__"):
            _stub("plotly.express")

    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    return importlib.import_module("joeynmt.helpers_for_audio")

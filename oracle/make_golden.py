# coding: utf-8
"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Generates ``tests/golden/*.npz`` by running the **unmodified reference** from ``/root/reference``
(through ``oracle/ref_shims.py``) on its own fixture wavs.  Run in the build container only:

    python oracle/make_golden.py

The outputs are committed; the GPU box (which has no ``/root/reference``) reads only the .npz files.

What is produced
  fixtures_pcm.npz       the ten ``test/data/speech/wav/260-123440-{0..9}.wav`` clips as int16 PCM
                         (inputs; LibriSpeech excerpts, CC BY 4.0) + the TSV ``n_frames`` column
  ref_fbank.npz          ``get_features(wav)`` per clip — raw log-mel float32, reference call chain
                         helpers_for_audio.py:100-127 → :41-68 → torchaudio kaldi.fbank
  ref_tables.npz         torchaudio's povey window and 80x256 mel bank (bit patterns the product
                         tables must reproduce)
  ref_cmvn.npz           ``CMVN(norm_means, norm_vars)`` outputs (data_augmentation.py:96-109):
                         full arrays for clip 1 (every flag combo) and clip 4 (default flags; the
                         silence / floor cases), head/tail rows + float64 checksums for the rest
  ref_specaugment.npz    ``SpecAugment.__call__`` (data_augmentation.py:38-73) under
                         ``np.random.seed(s)``: full outputs for clip 1, replayed mask tables and
                         sha256 of the output bytes for all clips, incl. the degenerate branches
  ref_processor.npz      ``SpeechProcessor.__call__`` (tokenizers.py:458-494) through the real
                         ``load_data`` stack: train (CMVN→SpecAugment, seeded) and eval with
                         ``max_length=500`` truncation, before=True/False
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402
from oracle import fbank_numpy as O  # noqa: E402

GOLD = ROOT / "tests" / "golden"
SPEECH = ref_shims.REFERENCE_ROOT / "test" / "data" / "speech"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    helpers = ref_shims.install(full_stack=True)
    import torch
    import torchaudio.compliance.kaldi as K
    from joeynmt.data_augmentation import CMVN, SpecAugment
    from joeynmt.tokenizers import SpeechProcessor

    GOLD.mkdir(parents=True, exist_ok=True)
    ids = [f"260-123440-{i}" for i in range(10)]
    tsv_frames = {}
    for line in (SPEECH / "test.tsv").read_text().splitlines()[1:]:
        cols = line.split("\t")
        tsv_frames[cols[0]] = int(cols[2])

    # ---- inputs -----------------------------------------------------------------------
    pcm = {}
    for i, uid in enumerate(ids):
        x, sr = ref_shims.load_wav_int16(SPEECH / "wav" / f"{uid}.wav")
        assert sr == 16000
        pcm[f"pcm{i}"] = x
    np.savez_compressed(GOLD / "fixtures_pcm.npz",
                        n_frames=np.array([tsv_frames[u] for u in ids], np.int32), **pcm)

    # ---- raw log-mel through the reference's own call chain ---------------------------
    feats = []
    for uid in ids:
        f = helpers.get_features(SPEECH, f"wav/{uid}.wav")
        assert f.dtype == np.float32 and f.shape == (tsv_frames[uid], 80)
        feats.append(f)
    np.savez_compressed(GOLD / "ref_fbank.npz", **{f"fbank{i}": f for i, f in enumerate(feats)})

    # ---- tables -----------------------------------------------------------------------
    win = K._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32)
    banks = K.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)[0]
    np.savez_compressed(GOLD / "ref_tables.npz", povey400=win.numpy(), mel80x256=banks.numpy())

    # ---- CMVN -------------------------------------------------------------------------
    out = {}
    for nm in (True, False):
        for nv in (True, False):
            tag = f"m{int(nm)}v{int(nv)}"
            for i, f in enumerate(feats):
                y = CMVN(norm_means=nm, norm_vars=nv)(f)
                assert y.dtype == np.float32
                if i == 1 or (i == 4 and nm and nv):
                    out[f"{tag}_full{i}"] = y
                out[f"{tag}_head{i}"] = y[:4].copy()
                out[f"{tag}_tail{i}"] = y[-4:].copy()
                out[f"{tag}_sum{i}"] = np.array(
                    [y.astype(np.float64).sum(), (y.astype(np.float64)**2).sum()])
    # all-silent utterance (var = 0 → std = 1e-5 branch, data_augmentation.py:105)
    silent = helpers.extract_fbank_features(torch.zeros(1, 4000), 16000)
    out["silent_fbank"] = silent
    out["silent_cmvn"] = CMVN()(silent)
    np.savez_compressed(GOLD / "ref_cmvn.npz", **out)

    # ---- SpecAugment ------------------------------------------------------------------
    out = {}
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
    cm = [CMVN()(f) for f in feats]
    meta = []
    for cname, cfg in cfgs.items():
        for seed in (0, 1, 2345):
            for i, x in enumerate(cm):
                np.random.seed(seed)
                y = SpecAugment(**cfg)(x)
                # replay with the oracle's draw routine to record the table
                np.random.seed(seed)
                kw = {k: v for k, v in cfg.items() if k != "mask_value"}
                masks = O.draw_specaugment_masks(x.shape[0], x.shape[1], **kw)
                if masks is None:
                    fm, tm = np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32)
                    untouched = 1
                else:
                    fm = np.array(masks[0], np.int32).reshape(-1, 2)
                    tm = np.array(masks[1], np.int32).reshape(-1, 2)
                    untouched = 0
                key = f"{cname}_s{seed}_c{i}"
                out[key + "_fm"] = fm
                out[key + "_tm"] = tm
                out[key + "_untouched"] = np.array(untouched)
                out[key + "_maskvalue"] = np.array(
                    cfg.get("mask_value", x.mean()), dtype=np.float32)
                out[key + "_changed"] = np.packbits(y != x)
                meta.append((key, sha(y)))
                if i == 1:
                    out[key + "_full"] = y
    out["sha_keys"] = np.array([m[0] for m in meta])
    out["sha_vals"] = np.array([m[1] for m in meta])
    np.savez_compressed(GOLD / "ref_specaugment.npz", **out)

    # ---- SpeechProcessor (real class, real get_features) --------------------------------
    out = {}
    variants = {
        "before": dict(norm_means=True, norm_vars=True, before=True),
        "after": dict(norm_means=True, norm_vars=True, before=False),
    }
    sa = dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0)
    for vname, ccfg in variants.items():
        proc = SpeechProcessor(level="frame", num_freq=80, max_length=500, min_length=200,
                               specaugment=sa, cmvn=ccfg)
        proc.root_path = SPEECH
        for i, uid in enumerate(ids):
            for is_train in (True, False):
                np.random.seed(1000 + i)
                y = proc(f"wav/{uid}.wav", is_train=is_train)
                key = f"{vname}_{'train' if is_train else 'eval'}_c{i}"
                if y is None:
                    out[key + "_none"] = np.array(1)
                    continue
                out[key + "_none"] = np.array(0)
                out[key + "_shape"] = np.array(y.shape)
                out[key + "_sha"] = np.array(sha(y))
                if i in (1, 2, 3):
                    out[key + "_full"] = y.astype(np.float32)
    np.savez_compressed(GOLD / "ref_processor.npz", **out)

    for p in sorted(GOLD.glob("*.npz")):
        print(f"{p.name:24s} {p.stat().st_size/1e3:9.1f} kB")


if __name__ == "__main__":
    main()

# coding: utf-8
"""
TEST INFRASTRUCTURE ONLY.  Golden outputs of the small host-side helpers of the path, produced by the
*unmodified* reference (``/root/reference`` through ``oracle/ref_shims.py``):

    python oracle/make_golden_misc.py          # writes tests/golden/ref_misc.npz

  get_n_frames   helpers_for_audio.py:93-96 over a sweep of lengths x sample rates (the float
                 expression ``int(N / sr * 1000)`` rounds differently from ``int(1000 * N / sr)``)
  pad_features   helpers_for_audio.py:130-170 on seeded ragged lists, default and explicit arguments
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402

GOLD = ROOT / "tests" / "golden"
RATES = (8000, 16000, 22050, 44100, 48000)


def n_frames_lengths():
    """Every length up to 5 s of 16 kHz audio, then a stride-7 sweep up to 600 000 samples."""
    return np.concatenate([np.arange(400, 80001), np.arange(80001, 600001, 7)]).astype(np.int64)


def pad_cases():
    """(seed, embed_size, pad_index) of the seeded ragged lists."""
    return [(0, 80, 1), (1, 80, 1), (2, 80, 0), (3, 40, 1), (4, 80, -3), (5, 80, 1)]


def pad_input(seed, embed_size):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 7))
    return [rng.standard_normal((int(rng.integers(1, 60)), embed_size)).astype(np.float32) for _ in range(n)]


def main():
    helpers = ref_shims.install(full_stack=False)
    out = {}
    lengths = n_frames_lengths()
    for sr in RATES:
        out[f"n_frames_sr{sr}"] = np.array([helpers.get_n_frames(int(n), sr) for n in lengths], np.int32)
    for seed, embed, pad in pad_cases():
        feats, lens, third = helpers.pad_features(pad_input(seed, embed), embed_size=embed, pad_index=pad)
        assert third is None
        out[f"pad{seed}_features"] = feats
        out[f"pad{seed}_lengths"] = np.array(lens, np.int32)
    np.savez_compressed(GOLD / "ref_misc.npz", **out)
    print(f"ref_misc.npz {(GOLD / 'ref_misc.npz').stat().st_size / 1e3:.1f} kB")


if __name__ == "__main__":
    main()

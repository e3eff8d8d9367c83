# coding: utf-8
"""
TEST INFRASTRUCTURE ONLY.  Golden batches for the batched sampler / collate path,
produced by the *unmodified* reference: ``joeynmt.data.load_data`` → ``SpeechDataset.make_iter``
(``SentenceBatchSampler`` / ``TokenBatchSampler``, ``datasets.py:1143-1292``) → ``collate_fn``
(``datasets.py:186-242``: ``pad_features`` + ``torch.tensor(src).float()``) on the repo's 10 speech
fixtures, configuration of ``test/unit/test_data.py:185-214`` (max_length 500, char targets).

    python oracle/make_golden_batches.py        # needs /root/reference; writes tests/golden/ref_batches.npz

Recorded per (split, batch_type): the order in which the base sampler visits the indices, the index
lists of the batches, ``src`` shapes and lengths, a SHA-256 of every batch's ``src`` and two full
``src`` tensors; plus per item the frame count and the tokenised target length the samplers see.
"""
import hashlib
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref_shims.install(full_stack=True)
    os.chdir(ref_shims.REFERENCE_ROOT)  # the data config uses paths relative to the repo root
    import torch
    from joeynmt.data import load_data

    seed = 42
    data_cfg = {
        "train": "test/data/speech/test",
        "test": "test/data/speech/test",
        "src": {"lang": "en", "level": "frame", "num_freq": 80, "max_length": 500,
                "tokenizer_type": "speech"},
        "trg": {"lang": "en", "level": "char", "lowercase": True, "max_length": 50,
                "voc_file": "test/data/speech/char.txt"},
        "dataset_type": "speech",
        "special_symbols": SimpleNamespace(**{
            "unk_token": "<unk>", "pad_token": "<pad>", "bos_token": "<s>", "eos_token": "</s>",
            "sep_token": None, "unk_id": 0, "pad_id": 1, "bos_id": 2, "eos_id": 3, "sep_id": None,
            "lang_tags": []}),
    }
    _, trg_vocab, train_data, _, test_data = load_data(data_cfg, datasets=["train", "test"], task="S2T")
    out = {}
    # what the samplers see per item: frames of the processed features, tokens of the target
    for split, data in (("train", train_data), ("test", test_data)):
        n_frames, trg_len, dropped = [], [], []
        for idx in range(len(data)):
            _, src, trg = data[idx]
            n_frames.append(-1 if src is None else len(src))
            trg_len.append(-1 if trg is None else len(trg))
            dropped.append(src is None)
        out[f"{split}_item_frames"] = np.array(n_frames, np.int32)
        out[f"{split}_item_trg_len"] = np.array(trg_len, np.int32)
        out[f"{split}_item_dropped"] = np.array(dropped)
    out["tsv_n_frames"] = np.array(train_data.df["n_frames"].tolist(), np.int32)

    for split, data, shuffle in (("train", train_data, True), ("test", test_data, False)):
        for batch_type, batch_size in (("sentence", 2), ("sentence", 3), ("token", 600), ("token", 1500)):
            loader = data.make_iter(batch_size=batch_size, batch_type=batch_type, shuffle=shuffle,
                                    seed=seed, pad_index=trg_vocab.pad_index,
                                    eos_index=trg_vocab.eos_index, device=torch.device("cpu"),
                                    num_workers=0)
            key = f"{split}_{batch_type}{batch_size}"
            # visiting order of the base sampler for this epoch, then the batches of the same epoch
            loader.batch_sampler.set_seed(seed)
            order = list(iter(loader.batch_sampler.sampler))
            loader.batch_sampler.set_seed(seed)
            index_batches = [list(b) for b in loader.batch_sampler]
            loader.batch_sampler.set_seed(seed)
            np.random.seed(seed)
            shas, shapes, lens = [], [], []
            for bi, batch in enumerate(loader):
                src = batch.src.numpy()
                assert src.dtype == np.float32
                assert batch.indices.tolist() == index_batches[bi]
                shas.append(sha(src))
                shapes.append(src.shape)
                lens.append(batch.src_length.tolist())
                if batch_type == "sentence" and batch_size == 2 and bi == 2:
                    out[f"{key}_full2"] = src  # the batch test_data.py:251,270 asserts the shape of
            out[f"{key}_order"] = np.array(order, np.int32)
            out[f"{key}_batches"] = np.array([len(b) for b in index_batches], np.int32)
            out[f"{key}_indices"] = np.array([i for b in index_batches for i in b], np.int32)
            out[f"{key}_shapes"] = np.array(shapes, np.int32)
            out[f"{key}_lengths"] = np.array([t for ln in lens for t in ln], np.int32)
            out[f"{key}_sha"] = np.array(shas)
            print(key, "order", order, "batches", index_batches, "shapes", shapes)
    np.savez_compressed(GOLD / "ref_batches.npz", **out)
    print(f"ref_batches.npz {(GOLD / 'ref_batches.npz').stat().st_size / 1e3:.1f} kB")


if __name__ == "__main__":
    main()

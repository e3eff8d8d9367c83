# coding: utf-8
"""
On-disk feature store — SURVEY.md §8(f-3): the reference's ``npy``-in-``ZIP_STORED`` archive with a
``<zip name>:<byte offset>:<byte size>`` manifest and the TSV columns ``id, src, n_frames, trg``
(``scripts/audiodata_utils.py:45-98``; reader: ``joeynmt/helpers_for_audio.py:72-89,100-127``;
producer: ``scripts/prepare_librispeech.py:60-112``, ``prepare_mustc.py``, ``prepare_europarl.py``,
``prepare_openslr.py``).

``create_zip`` / ``get_zip_manifest`` / ``save_tsv`` / ``load_tsv`` keep the reference's names,
arguments and results.  :class:`ZipFeatureWriter` and :func:`extract_corpus` are the batched
replacement of the prep scripts' per-utterance loop (``_extract`` → ``np.save`` → ``create_zip`` →
``get_zip_manifest``): whole batches of utterances go through the fused GPU front-end and their
``.npy`` images are written straight into the archive, with the manifest entry of every member
known at write time.  A member's payload is byte-identical to what ``np.save`` writes, its offset is
what ``get_zip_manifest`` computes (``header_offset + 30 + len(filename)``, :52), so archives are
interchangeable with the reference's in both directions.
"""
import csv
import io
import zipfile
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from joeys2t_b200.helpers_for_audio import _is_npy_data


def _member_payload_span(archive, info: zipfile.ZipInfo) -> Tuple[int, int]:
    """(offset, size) of a STORED member's payload, from its LOCAL file header: 30 fixed bytes, then the
    name and the extra field whose lengths are the two little-endian uint16 at bytes 26-29.  The reference
    assumes ``header_offset + 30 + len(filename)`` (``audiodata_utils.py:52``), i.e. an empty extra field;
    that is what ``create_zip`` and :class:`ZipFeatureWriter` write, and it is asserted here."""
    archive.seek(info.header_offset)
    fixed = archive.read(30)
    if len(fixed) != 30 or fixed[:4] != b"PK\x03\x04":
        raise ValueError(f"{info.filename}: no local file header at byte {info.header_offset}")
    name_len = int.from_bytes(fixed[26:28], "little")
    extra_len = int.from_bytes(fixed[28:30], "little")
    if info.compress_type != zipfile.ZIP_STORED or extra_len != 0:
        raise ValueError(f"{info.filename}: the manifest format addresses uncompressed members without extra field")
    return info.header_offset + 30 + name_len, info.file_size


def get_zip_manifest(zip_path: Path, npy_root: Optional[Path] = None) -> Dict[str, str]:
    """Interface of ``scripts/audiodata_utils.py:45-63``: ``{utt_id: "<zip name>:<offset>:<size>"}`` for
    every member of an uncompressed archive — the strings ``get_features`` resolves
    (``joeynmt/helpers_for_audio.py:77-89``).  With ``npy_root`` every member is additionally compared with
    ``<npy_root>/<utt_id>.npy``."""
    zip_path = Path(zip_path)
    with zipfile.ZipFile(zip_path) as zf:
        members = [m for m in zf.infolist() if not m.is_dir()]
    manifest: Dict[str, str] = {}
    with open(zip_path, "rb") as archive:
        for member in members:
            utt_id = Path(member.filename).stem
            offset, size = _member_payload_span(archive, member)
            archive.seek(offset)
            payload = archive.read(size)
            if len(payload) != size or not _is_npy_data(payload):
                raise AssertionError(f"{zip_path.name}: member {member.filename} is not a .npy image")
            if npy_root is not None:
                on_disk = np.load(str(Path(npy_root) / f"{utt_id}.npy"))
                if not np.allclose(np.load(io.BytesIO(payload)), on_disk):
                    raise AssertionError(f"{utt_id}: archive member differs from {npy_root}/{utt_id}.npy")
            manifest[utt_id] = f"{zip_path.name}:{offset}:{size}"
    return manifest


def create_zip(data_root: Path, zip_path: Path) -> None:
    """Interface of ``scripts/audiodata_utils.py:66-73``: every ``*.npy`` of ``data_root`` into one
    uncompressed archive.  The payloads are copied as they are (no re-serialisation), through the same
    writer as the batched path, so the manifest of the result is known without re-reading it."""
    with ZipFeatureWriter(Path(zip_path)) as writer:
        for path in sorted(Path(data_root).glob("*.npy")):
            payload = path.read_bytes()
            if not _is_npy_data(payload):
                raise RuntimeError(f"{path}: not a .npy file")
            writer.add_npy_image(path.stem, payload)


def save_tsv(df, path: Path, header: bool = True) -> None:
    """``audiodata_utils.py:76-85``"""
    df.to_csv(path.as_posix(), sep="\t", header=header, index=False, encoding="utf-8",
              escapechar="\\", quoting=csv.QUOTE_NONE)


def load_tsv(path: Path):
    """``audiodata_utils.py:88-98``"""
    import pandas as pd  # pylint: disable=import-outside-toplevel
    return pd.read_csv(path.as_posix(), sep="\t", header=0, encoding="utf-8", escapechar="\\",
                       quoting=csv.QUOTE_NONE, na_filter=False)


def npy_bytes(features: np.ndarray) -> bytes:
    """The exact byte image ``np.save`` writes for ``features`` (magic ``\\x93NUMPY``)."""
    buf = io.BytesIO()
    np.save(buf, np.ascontiguousarray(features))
    return buf.getvalue()


class ZipFeatureWriter:
    """Append ``(utt_id, features)`` pairs to a ``ZIP_STORED`` archive and collect the manifest.

    Equivalent to ``np.save(feature_root / f"{id}.npy")`` for every utterance followed by
    ``create_zip`` and ``get_zip_manifest`` — without the intermediate files and without
    re-reading the archive.
    """

    def __init__(self, zip_path: Path, mode: str = "w"):
        self.zip_path = Path(zip_path)
        self._zf = zipfile.ZipFile(self.zip_path, mode, zipfile.ZIP_STORED, allowZip64=True)
        self.manifest: Dict[str, str] = {}
        self.n_frames: Dict[str, int] = {}

    def add(self, utt_id: str, features: np.ndarray) -> str:
        assert features.ndim == 2, "spectrogram must be a 2-D array."
        entry = self.add_npy_image(utt_id, npy_bytes(features))
        self.n_frames[utt_id] = int(features.shape[0])
        return entry

    def add_npy_image(self, utt_id: str, data: bytes) -> str:
        """Append an already serialised ``.npy`` image as member ``<utt_id>.npy``."""
        if utt_id in self.manifest:
            raise ValueError(f"duplicate utterance id {utt_id!r}")
        name = f"{utt_id}.npy"
        info = zipfile.ZipInfo(name)  # fixed timestamp: archives are reproducible
        info.compress_type = zipfile.ZIP_STORED
        self._zf.writestr(info, data)
        written = self._zf.getinfo(name)
        # local file header = 30 bytes + name (+ extra, empty here) — audiodata_utils.py:52
        offset = written.header_offset + 30 + len(name.encode("utf-8")) + len(written.extra)
        entry = f"{self.zip_path.name}:{offset}:{len(data)}"
        self.manifest[utt_id] = entry
        return entry

    def close(self) -> Dict[str, str]:
        self._zf.close()
        return self.manifest

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def extract_corpus(
    items: Iterable[Tuple[str, "np.ndarray"]],
    zip_path: Path,
    batch_utterances: int = 64,
    batch_seconds: float = 1000.0,
    sample_rate: int = 16000,
) -> Tuple[Dict[str, str], Dict[str, int], List[Tuple[str, str]]]:
    """GPU replacement of the prep scripts' extraction loop (``prepare_librispeech.py:74-107``).

    ``items`` yields ``(utt_id, waveform)`` — waveform as the scripts pass it to
    ``extract_fbank_features`` (float in [-1, 1), shape (N,) or (C, N)) or int16 PCM.  Utterances are
    grouped into batches of at most ``batch_utterances`` / ``batch_seconds`` of audio; each batch is
    one fused pass of the CUDA front-end (raw log-mel, no CMVN — the archive holds un-normalised
    features, :78-84) through :class:`joeys2t_b200.frontend.HostPipeline`, so the H2D copy and the
    kernels of batch ``i + 1`` run while batch ``i`` is serialised into the archive.

    Like the reference's ``_extract`` (:75-88), an utterance that cannot be processed (too short for
    one 25 ms frame) is reported and gets ``n_frames = 0`` instead of aborting the run.

    :returns: ``(manifest, n_frames, failed)`` — manifest ``{id: "name.zip:offset:size"}``,
        ``n_frames`` ``{id: T}`` (0 for failures), ``failed`` ``[(id, reason)]``.
    """
    import torch  # pylint: disable=import-outside-toplevel

    from joeys2t_b200 import frontend, tables  # pylint: disable=import-outside-toplevel

    if int(sample_rate) != tables.SAMPLE_RATE:
        raise ValueError(f"the front-end is specialised for {tables.SAMPLE_RATE} Hz audio")
    failed: List[Tuple[str, str]] = []
    n_frames: Dict[str, int] = {}
    max_samples = int(batch_seconds * sample_rate)
    n_slots = 2
    max_pcm_bytes = max_samples * 4 + 16 * batch_utterances
    max_rows = max_samples // tables.FRAME_SHIFT + batch_utterances
    pipe = frontend.HostPipeline(n_slots, max_pcm_bytes, max_rows)
    stage = [torch.empty(max_pcm_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(n_slots)]

    with ZipFeatureWriter(zip_path) as writer:
        pending = []  # (slot, ids, plan) in flight

        def drain(keep: int):
            while len(pending) > keep:
                slot, ids, plan = pending.pop(0)
                host = pipe.result(slot).numpy()
                off = 0
                for utt_id, t in zip(ids, plan.n_frames.tolist()):
                    writer.add(utt_id, host[off:off + t])
                    n_frames[utt_id] = t
                    off += t
                plan.close()

        def submit(ids: Sequence[str], waves: Sequence[np.ndarray]):
            if not ids:
                return
            drain(n_slots - 1)  # the slot about to be reused has been written out
            j = pipe._next  # pylint: disable=protected-access
            packed = frontend.PackedPCM(list(waves), host=stage[j])
            plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
            pending.append((pipe.submit(packed, plan), list(ids), plan))

        ids, waves, total = [], [], 0
        for utt_id, w in items:
            arr = np.asarray(w.detach().cpu().numpy() if hasattr(w, "detach") else w)
            n = int(arr.shape[-1])
            if tables.num_frames(n) <= 0:
                failed.append((utt_id, f"waveform of {n} samples is shorter than one 25 ms frame"))
                n_frames[utt_id] = 0
                continue
            if n > max_samples:
                raise ValueError(f"{utt_id}: {n / sample_rate:.0f} s of audio exceeds batch_seconds")
            if ids and (len(ids) >= batch_utterances or total + n > max_samples):
                submit(ids, waves)
                ids, waves, total = [], [], 0
            ids.append(utt_id)
            waves.append(arr)
            total += n
        submit(ids, waves)
        drain(0)
        pipe.synchronize()
        manifest = dict(writer.manifest)
    return manifest, n_frames, failed

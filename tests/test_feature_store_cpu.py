# coding: utf-8
"""
Feature store: npy-in-ZIP_STORED archive + ``name.zip:offset:size`` manifest.
CPU only: format compatibility with the reference's reader and manifest builder
(``scripts/audiodata_utils.py:45-73``, ``joeynmt/helpers_for_audio.py:72-89``) in both directions.
"""
import importlib.util
import io
import sys
import zipfile
from pathlib import Path

import numpy as np
import pytest

from joeys2t_b200 import feature_store as FS
from joeys2t_b200 import helpers_for_audio as HA
from oracle import ref_shims


def _corpus(seed=0, n=7):
    rng = np.random.default_rng(seed)
    # frame counts around the edge cases: 1 frame, odd sizes, a long one
    lens = [1, 2, 31, 32, 33, 172, 1248][:n]
    return {f"utt-{i:03d}_x": rng.standard_normal((t, 80)).astype(np.float32) for i, t in enumerate(lens)}


def test_writer_manifest_and_reader_round_trip(tmp_path):
    feats = _corpus()
    zip_path = tmp_path / "fbank80.zip"
    with FS.ZipFeatureWriter(zip_path) as w:
        for k, v in feats.items():
            w.add(k, v)
    manifest = w.manifest
    assert set(manifest) == set(feats)
    # the archive is a plain uncompressed zip of .npy members
    with zipfile.ZipFile(zip_path) as z:
        assert all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist())
        assert sorted(i.filename for i in z.infolist()) == sorted(f"{k}.npy" for k in feats)
    # manifest entries are what the manifest builder computes from the archive alone
    assert FS.get_zip_manifest(zip_path) == manifest
    # and what get_features reads back through "name.zip:offset:size" (helpers_for_audio.py:100-127)
    for k, v in feats.items():
        name, off, size = manifest[k].split(":")
        assert name == zip_path.name
        got = HA.get_features(tmp_path, manifest[k])
        assert got.dtype == np.float32 and np.array_equal(got, v)
        raw = zip_path.read_bytes()[int(off):int(off) + int(size)]
        assert raw == FS.npy_bytes(v)  # byte image of np.save
        assert w.n_frames[k] == v.shape[0]


def test_create_zip_from_npy_directory_matches_writer_payload(tmp_path):
    feats = _corpus(1, 4)
    root = tmp_path / "fbank80"
    root.mkdir()
    for k, v in feats.items():
        np.save(root / f"{k}.npy", v)
    FS.create_zip(root, root.with_suffix(".zip"))
    manifest = FS.get_zip_manifest(root.with_suffix(".zip"), npy_root=root)
    for k, v in feats.items():
        assert np.array_equal(HA.get_features(tmp_path, manifest[k]), v)
    # same payload bytes as the streaming writer produces
    with FS.ZipFeatureWriter(tmp_path / "w.zip") as w:
        for k, v in feats.items():
            w.add(k, v)
    for k in feats:
        _, o1, s1 = manifest[k].split(":")
        _, o2, s2 = w.manifest[k].split(":")
        a = root.with_suffix(".zip").read_bytes()[int(o1):int(o1) + int(s1)]
        b = (tmp_path / "w.zip").read_bytes()[int(o2):int(o2) + int(s2)]
        assert a == b


def test_reader_rejects_non_npy_payload_and_bad_paths(tmp_path):
    zip_path = tmp_path / "x.zip"
    with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_STORED) as z:
        z.writestr("a.txt", b"hello world")
    with pytest.raises(ValueError):
        HA.get_features(tmp_path, "x.zip:30:5")  # not an npy image (helpers_for_audio.py:84-88)
    with pytest.raises(FileNotFoundError):
        HA.get_features(tmp_path, "missing.zip:0:10")
    with pytest.raises(ValueError):
        HA.get_features(tmp_path, "x.zip:1")  # offset without size (helpers_for_audio.py:123-124)
    with FS.ZipFeatureWriter(tmp_path / "d.zip") as w:
        w.add("a", np.zeros((2, 80), np.float32))
        with pytest.raises(ValueError):
            w.add("a", np.zeros((2, 80), np.float32))


def test_tsv_round_trip(tmp_path):
    import pandas as pd
    df = pd.DataFrame({"id": ["a", "b"], "src": ["f.zip:38:100", "f.zip:200:64"], "n_frames": [3, 5],
                       "trg": ["Poor Alice!", "tab\\there"]})
    FS.save_tsv(df, tmp_path / "t.tsv")
    back = FS.load_tsv(tmp_path / "t.tsv")
    assert back["id"].tolist() == ["a", "b"] and back["n_frames"].tolist() == [3, 5]
    assert back["src"].tolist() == df["src"].tolist()
    assert (tmp_path / "t.tsv").read_text().splitlines()[0] == "id\tsrc\tn_frames\ttrg"


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference is not mounted")
def test_interchangeable_with_reference_tools(tmp_path):
    """Archives written here are read by the reference's own manifest builder and feature reader,
    and archives built by the reference's ``create_zip`` are read by ours."""
    ref_ha = ref_shims.install()
    spec = importlib.util.spec_from_file_location(
        "ref_audiodata_utils", ref_shims.REFERENCE_ROOT / "scripts" / "audiodata_utils.py")
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)

    feats = _corpus(2)
    zip_path = tmp_path / "fbank80.zip"
    with FS.ZipFeatureWriter(zip_path) as w:
        for k, v in feats.items():
            w.add(k, v)
    assert ref_utils.get_zip_manifest(zip_path) == w.manifest
    for k, v in feats.items():
        assert np.array_equal(ref_ha.get_features(tmp_path, w.manifest[k]), v)

    root = tmp_path / "npy"
    root.mkdir()
    for k, v in feats.items():
        np.save(root / f"{k}.npy", v)
    ref_utils.create_zip(root, tmp_path / "ref.zip")
    ref_manifest = ref_utils.get_zip_manifest(tmp_path / "ref.zip", npy_root=root)
    assert FS.get_zip_manifest(tmp_path / "ref.zip") == ref_manifest
    for k, v in feats.items():
        assert np.array_equal(HA.get_features(tmp_path, ref_manifest[k]), v)

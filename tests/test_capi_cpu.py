# coding: utf-8
"""
CPU-side checks of the drop-in boundary: the in-tree shared library loads without a GPU, exports
every symbol ``include/joeys2t_b200.h`` declares, its pure-host entry points answer, and the
product path refuses to run (loudly, no fallback) when no CUDA device is usable.
No compute call is made here.
"""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from joeys2t_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "joeys2t_b200.h"


@pytest.fixture(scope="module")
def lib():
    if _lib.is_stale():
        _lib.build()
    return _lib.load()


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(js2t_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert getattr(lib, name) is not None, name


def test_version_and_frame_count(lib):
    assert lib.js2t_version() == 100
    # TA:63-67 (snip_edges): 1 + (n - 400) // 160, 0 below one window
    for n, want in ((0, 0), (399, 0), (400, 1), (559, 1), (560, 2), (16000, 98), (240000, 1498)):
        assert lib.js2t_num_frames(n) == want


def test_argument_errors_set_last_error(lib):
    assert lib.js2t_ctx_create(0, None) == _lib.ERR_INVALID
    assert b"NULL" in lib.js2t_last_error()
    assert lib.js2t_plan_set_cmvn(None, 1, 1, 1, 1) == _lib.ERR_INVALID
    assert lib.js2t_fbank_execute(None, None, None, None) == _lib.ERR_INVALID
    assert lib.js2t_plan_total_frames(None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    """Without a CUDA device the C ABI reports JS2T_ERR_CUDA and the host API raises."""
    h = ctypes.c_void_p()
    assert lib.js2t_ctx_create(0, ctypes.byref(h)) == _lib.ERR_CUDA
    from joeys2t_b200 import frontend, helpers_for_audio
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        frontend.get_context(0)
    # the reference's error convention: failures surface as ValueError (helpers_for_audio.py:56-62)
    with pytest.raises(ValueError, match="no CPU fallback"):
        helpers_for_audio.extract_fbank_features(torch.zeros(1, 16000), 16000)


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under joeys2t_b200/ may import it."""
    for py in (ROOT / "joeys2t_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), py


def test_packed_pcm_alignment():
    from joeys2t_b200 import frontend
    waves = [np.arange(401, dtype=np.int16), np.zeros(403, np.float32), np.ones(999, np.int16)]
    p = frontend.PackedPCM(waves)
    assert (p.byte_off % 16 == 0).all() and p.is_f32.tolist() == [0, 1, 0]
    hv = p.host.numpy()
    assert np.array_equal(hv[:802].view(np.int16), waves[0])
    o = int(p.byte_off[2])
    assert np.array_equal(hv[o:o + 1998].view(np.int16), waves[2])


def test_packed_pcm_into_caller_staging():
    """PackedPCM packs into a caller-provided staging buffer (a tensor, or a callable that sizes one)
    exactly as into its own, and refuses a buffer that is too small."""
    import torch
    from joeys2t_b200 import frontend
    rng = np.random.default_rng(0)
    waves = [rng.integers(-100, 100, 401).astype(np.int16), (rng.random(777) - 0.5).astype(np.float32),
             rng.integers(-100, 100, (2, 1000)).astype(np.int16)]  # (C, N): channel 0 is taken (quirk Q1)
    own = frontend.PackedPCM(waves)
    asked = []

    def provide(nbytes):
        asked.append(nbytes)
        return torch.full((nbytes + 64,), 0xAB, dtype=torch.uint8)

    for host in (torch.zeros(own.nbytes + 100, dtype=torch.uint8), provide):
        p = frontend.PackedPCM(waves, host=host)
        assert p.nbytes == own.nbytes and np.array_equal(p.byte_off, own.byte_off)
        assert np.array_equal(p.n_samples, [401, 777, 1000]) and p.is_f32.tolist() == [0, 1, 0]
        for a, o in zip((waves[0], waves[1], waves[2][0]), p.byte_off):
            got = p.host.numpy()[o:o + a.nbytes]
            assert got.tobytes() == a.tobytes()
    assert asked == [own.nbytes]
    with pytest.raises(ValueError):
        frontend.PackedPCM(waves, host=torch.zeros(own.nbytes - 1, dtype=torch.uint8))


def test_reference_mel_bank_is_torchaudios(ref_tables):
    """js2t_reference_mel_bank: the bank the kernels were compiled with == torchaudio's get_mel_banks output
    (tests/golden/ref_tables.npz), bit for bit.  Pure host code."""
    lib = _lib.load()
    bank = np.zeros((80, 256), np.float32)
    assert lib.js2t_reference_mel_bank(bank.ctypes.data) == _lib.OK
    assert np.array_equal(bank.view(np.uint32), ref_tables["mel80x256"].view(np.uint32))


def test_pack_pcm_gathers_ragged_batches():
    """js2t_pack_pcm (host only): every utterance lands at its 16-byte aligned offset, for one thread and for
    the copy pool, int16 and float32 mixed, unaligned sources."""
    import ctypes
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for trial in range(6):
        arrs = []
        for _ in range(int(rng.integers(1, 24))):
            n = int(rng.integers(1, 300000))
            a = rng.integers(-30000, 30000, n + 3).astype(np.int16)[int(rng.integers(0, 3)):][:n]
            arrs.append(np.ascontiguousarray(a) if rng.uniform() < 0.5 else
                        np.ascontiguousarray(a.astype(np.float32) / np.float32(32768.0)))
        sizes = np.array([a.nbytes for a in arrs], np.int64)
        off = np.concatenate([[0], np.cumsum((sizes + 15) // 16 * 16)[:-1]]).astype(np.int64)
        total = int(off[-1] + sizes[-1])
        dst = np.full(total + 32, 0xAB, np.uint8)
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        rc = lib.js2t_pack_pcm(len(arrs), ptrs, sizes.ctypes.data, off.ctypes.data, dst.ctypes.data, total,
                               1 if trial % 2 else 0)
        assert rc == _lib.OK, lib.js2t_last_error()
        for a, o in zip(arrs, off):
            assert np.array_equal(dst[o:o + a.nbytes], a.view(np.uint8))
        assert (dst[total:] == 0xAB).all()
    # a destination that is too small is refused
    small = np.zeros(8, np.uint8)
    rc = lib.js2t_pack_pcm(len(arrs), ptrs, sizes.ctypes.data, off.ctypes.data, small.ctypes.data, 8, 1)
    assert rc == _lib.ERR_INVALID


def test_a_forked_child_is_told_what_to_do():
    """The reference's DataLoader workers are forked; a CUDA context does not survive fork, so a
    child that inherited the module's contexts must get a clear error instead of a driver failure."""
    import os

    from joeys2t_b200 import frontend

    old = frontend._owner_pid
    try:
        frontend._owner_pid = os.getpid() + 1  # as if the contexts had been created by the parent
        with pytest.raises(RuntimeError, match="do not survive fork"):
            frontend.get_context()
        frontend._owner_pid = os.getpid()
        frontend._check_not_forked()  # same process: fine
    finally:
        frontend._owner_pid = old

# coding: utf-8
"""PCM ingest, CPU only: the oracle's ``reformat_freq`` against the reference's
own function (body extracted from ``scripts/gradio_demo.py:35-45`` with ``ast`` — the module itself
cannot be imported: it needs gradio and downloads models at import time), and the summation order
numpy uses for ``mean(axis=1)`` that the CUDA kernel reproduces."""
import ast

import numpy as np
import pytest

from oracle import fbank_numpy as O
from oracle import ref_shims


def _signals():
    rng = np.random.default_rng(5)
    n = 3 * 4001
    t = np.arange(n)
    yield "speech-like int16", (3000 * np.sin(t / 37.0) * rng.random(n)).astype(np.int16)
    yield "full-scale int16", rng.integers(-32768, 32768, n).astype(np.int16)
    yield "negative peak larger than positive (overflowing cast)", \
        np.where(t % 7 == 0, -3000, 1000).astype(np.int16)
    yield "all non-positive (max -> 1)", (-np.abs(rng.integers(0, 200, n))).astype(np.int16)
    yield "zeros", np.zeros(n, np.int16)
    yield "float32 in [-1, 1) (max < 1 -> divisor 1)", (rng.random(n) * 2 - 1).astype(np.float32)
    yield "float32 int16-range", (rng.random(n) * 65535 - 32768).astype(np.float32)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference is not mounted")
def test_oracle_equals_reference_function():
    src = (ref_shims.REFERENCE_ROOT / "scripts" / "gradio_demo.py").read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "reformat_freq")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "gradio_demo.py", "exec"), ns)  # noqa: S102
    ref = ns["reformat_freq"]
    for name, y in _signals():
        with np.errstate(all="ignore"):
            a, sa = ref(48000, y.copy())
            b, sb = O.reformat_freq(48000, y.copy())
        assert sa == sb == 16000 and a.dtype == b.dtype == np.int16, name
        assert np.array_equal(a, b), name
        c, sc = ref(16000, y)
        assert sc == 16000 and c is y
    for fn_ in (ref, O.reformat_freq):
        with pytest.raises(ValueError):
            fn_(44100, np.zeros(30, np.int16))
        with pytest.raises(ValueError):
            fn_(48000, np.zeros(31, np.int16))  # reshape(-1, 3)


def test_numpy_mean_order_is_left_to_right():
    # the kernel computes ((a + b) + c) / 3 — this is what numpy's mean(axis=1) does for (M, 3)
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        a = (rng.standard_normal((20000, 3)) * 10.0**rng.integers(-6, 6, size=(20000, 3))).astype(dt)
        left = ((a[:, 0] + a[:, 1]) + a[:, 2]) / dt(3)
        assert np.array_equal(a.mean(axis=1), left)

# coding: utf-8
"""Decision gate for a tensor-core DFT — the NUMERICS half, on the CPU:

    python tools/sim_tc_dft.py

Emulates the 512-point real FFT of the fbank path as two radix-16 stages of small GEMMs on the 5th-gen
tensor cores (frames are the M dimension; stage 1: one [M x 32] x [32 x 32] GEMM per n2 with the povey
window, the DC-mean correction and the 16-point DFT folded into the constant matrix; stage 2: one per k1
with the W256 twiddles folded in), with SPLIT-PRECISION operands and fp32 accumulation:

    f16x2   operands as hi + lo fp16 pairs (3 MMAs per product: hi*hi + hi*lo + lo*hi), kind::f16
    tf32x3  the same with tf32 operands (10-bit mantissa, truncated like the hardware does), kind::tf32
    tf32x3rn  hi rounded to nearest first (cvt.rna), lo = exact remainder (truncated by the MMA)
    f16     single fp16 operands (1 MMA) — what "just use the tensor cores" would give
    bf16x3  three bf16 terms per operand (6 MMAs)

and compares the resulting log-mel with the reference goldens (tests/golden/ref_fbank.npz) on the ten
fixture clips — the same gate as the product kernel (max |err| <= 1e-3; the FP32 FFT kernel sits at
2-4e-4, the reference's own fp32-vs-fp64 noise is 4.5e-4).  Everything after stage 2 (real-input split,
power, mel, log) is fp32 like the product kernel.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from joeys2t_b200 import tables  # noqa: E402

F32 = np.float32


def trunc_tf32(x):
    return (np.asarray(x, F32).view(np.uint32) & np.uint32(0xFFFFE000)).view(F32)


def bf16_rn(x):
    u = np.asarray(x, F32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F32)


def split(x, mode):
    """list of operand terms (float32 arrays holding exactly representable values)"""
    x = np.asarray(x, F32)
    if mode == "f16":
        return [x.astype(np.float16).astype(F32)]
    if mode == "f16x2":
        hi = x.astype(np.float16).astype(F32)
        lo = (x - hi).astype(np.float16).astype(F32)
        return [hi, lo]
    if mode == "tf32x3":
        hi = trunc_tf32(x)
        lo = trunc_tf32(x - hi)
        return [hi, lo]
    if mode == "tf32x3rn":  # hi rounded to nearest (cvt.rna.tf32.f32), lo = x - hi exact, truncated by the MMA
        u = x.view(np.uint32).astype(np.uint64)
        hi = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(F32)
        lo = trunc_tf32(x - hi)
        return [hi, lo]
    if mode == "bf16x3":
        a = bf16_rn(x)
        b = bf16_rn(x - a)
        c = bf16_rn(x - a - b)
        return [a, b, c]
    if mode == "f32":
        return [x]
    raise ValueError(mode)


def mma(a_terms, b_terms, order):
    """sum of the term products whose combined order is <= `order`, each product fp32-accumulated"""
    acc = None
    for i, a in enumerate(a_terms):
        for j, b in enumerate(b_terms):
            if i + j <= order:
                p = (a.astype(np.float64) @ b.astype(np.float64)).astype(F32)  # exact products, fp32 result
                acc = p if acc is None else (acc + p).astype(F32)
    return acc


def stage_matrices():
    win = tables.povey_window().astype(np.float64)
    # stage 1, per n2: rows (n1, c) for n1 = 0..12 (26 rows) + 1 row for the DC mean; cols (k1, re|im)
    b1 = np.zeros((16, 27, 32))
    for n2 in range(16):
        for n1 in range(13):
            for c in range(2):
                j = 2 * (16 * n1 + n2) + c
                w = win[j] if j < 400 else 0.0
                for k1 in range(16):
                    th = 2 * np.pi * n1 * k1 / 16
                    # z = y_even + i y_odd ; X1 = sum z * exp(-i th)
                    if c == 0:
                        b1[n2, 2 * n1, 2 * k1] += w * np.cos(th)
                        b1[n2, 2 * n1, 2 * k1 + 1] += -w * np.sin(th)
                    else:
                        b1[n2, 2 * n1 + 1, 2 * k1] += w * np.sin(th)
                        b1[n2, 2 * n1 + 1, 2 * k1 + 1] += w * np.cos(th)
        # DC mean row: every sample contributes -(1 - 0.97) m
        b1[n2, 26] = -(1.0 - float(F32(0.97))) * b1[n2, :26].sum(0)
    # stage 2, per k1: rows (n2, re|im), cols (k2, re|im); Z[k1 + 16 k2] = sum_n2 X1[k1,n2] W256^(n2 k1) W16^(n2 k2)
    b2 = np.zeros((16, 32, 32))
    for k1 in range(16):
        for n2 in range(16):
            for k2 in range(16):
                th = 2 * np.pi * (n2 * k1 / 256 + n2 * k2 / 16)
                c, s = np.cos(th), np.sin(th)
                # (xr + i xi)(c - i s) = (xr c + xi s) + i(xi c - xr s)
                b2[k1, 2 * n2, 2 * k2] = c
                b2[k1, 2 * n2 + 1, 2 * k2] = s
                b2[k1, 2 * n2, 2 * k2 + 1] = -s
                b2[k1, 2 * n2 + 1, 2 * k2 + 1] = c
    return b1, b2


def fbank_tc(pcm, mode, b1, b2, order=1, in_scale=0.25, mid_scale=2.0**-4):
    x = pcm.astype(F32)
    T = 1 + (len(x) - 400) // 160
    idx = 160 * np.arange(T)[:, None] + np.arange(400)[None, :]
    fr = x[idx]
    # frame-independent pre-emphasis d[j] = x[j] - 0.97 x[j-1] (window[0] = 0 kills j = 0), DC mean per frame
    d = np.empty_like(fr)
    d[:, 1:] = fr[:, 1:] - F32(0.97) * fr[:, :-1]
    d[:, 0] = 0
    m = fr.astype(np.float64).sum(1).astype(F32) / F32(400.0)
    a_full = np.zeros((T, 16, 27), F32)
    for n2 in range(16):
        for n1 in range(13):
            for c in range(2):
                j = 2 * (16 * n1 + n2) + c
                if j < 400:
                    a_full[:, n2, 2 * n1 + c] = d[:, j]
        a_full[:, n2, 26] = m
    a_full *= F32(in_scale)
    x1 = np.zeros((T, 16, 32), F32)  # [n2][k1 re|im]
    for n2 in range(16):
        x1[:, n2] = mma(split(a_full[:, n2], mode), split((b1[n2] * mid_scale).astype(F32), mode), order)
    z = np.zeros((T, 256), np.complex64)
    for k1 in range(16):
        a2 = np.zeros((T, 32), F32)
        a2[:, 0::2] = x1[:, :, 2 * k1]
        a2[:, 1::2] = x1[:, :, 2 * k1 + 1]
        o = mma(split(a2, mode), split(b2[k1].astype(F32), mode), order)
        z[:, k1 + 16 * np.arange(16)] = o[:, 0::2] + 1j * o[:, 1::2]
    # fp32 tail like the product kernel: real split, power, mel, log
    k = np.arange(257)
    zk = z[:, k % 256]
    zc = np.conj(z[:, (256 - k) % 256])
    w = np.exp(-2j * np.pi * k / 512).astype(np.complex64)
    y = (F32(0.5) * (zk + zc) - F32(0.5) * 1j * w * (zk - zc)).astype(np.complex64)
    p = (y.real.astype(F32)**2 + y.imag.astype(F32)**2) * F32(1.0 / (in_scale * mid_scale)**2)
    mel = tables.mel_banks()
    e = p[:, :256] @ mel.T.astype(F32)
    return np.log(np.maximum(e, F32(1.1920929e-07))).astype(F32)


def main():
    gold = np.load(ROOT / "tests" / "golden" / "ref_fbank.npz")
    pcmz = np.load(ROOT / "tests" / "golden" / "fixtures_pcm.npz")
    b1, b2 = stage_matrices()
    print("max |log-mel - reference golden| per fixture clip (gate: 1e-3)")
    print(f"{'mode':22s} " + " ".join(f"clip{i:<6d}" for i in range(10)) + "   worst")
    for mode, order in (("f32", 0), ("f16x2", 1), ("f16x2", 2), ("tf32x3", 1), ("tf32x3", 2), ("tf32x3rn", 1),
                        ("tf32x3rn", 2), ("bf16x3", 2),
                        ("f16", 0)):
        errs = []
        for i in range(10):
            got = fbank_tc(pcmz[f"pcm{i}"], mode, b1, b2, order)
            errs.append(float(np.abs(got - gold[f"fbank{i}"]).max()))
        n_mma = sum(1 for a in range(3) for b in range(3)
                    if a + b <= order and a < len(split(np.zeros(1), mode)) and b < len(split(np.zeros(1), mode)))
        print(f"{mode + f' ({n_mma} MMA)':22s} " + " ".join(f"{e:10.2e}" for e in errs) + f"   {max(errs):.2e}")


if __name__ == "__main__":
    main()

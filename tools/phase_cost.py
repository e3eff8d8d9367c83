# coding: utf-8
"""Tuning aid (GPU box): time the persistent fbank kernel with individual phases skipped
(option "debug_skip"; results are wrong, only the timing is of interest).
   python tools/phase_cost.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 3
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("stats")
    sets.append((plan, packed.to_device(), plan.empty_output()))
frames = np.mean([s[0].total_frames for s in sets])


def timeit(skip, n=30, fused=False):
    for plan, _, _ in sets:
        plan.set_option("debug_skip", skip)
        plan.set_cmvn("utterance" if fused else "stats")
    for i in range(6):
        p, d, o = sets[i % R]
        p.execute(d, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        p, d, o = sets[i % R]
        p.execute(d, o)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


base = timeit(0)
print(f"frames/launch {frames:.0f}; full pipeline (fbank+stats kernel + finalize) {base:.1f} us")
names = {1: "staging", 2: "whole FFT phase", 4: "mel", 8: "store", 16: "butterflies only", 32: "exchange only",
         48: "butterflies+exchange", 2 | 4: "FFT+mel", 1 | 2 | 4 | 8: "everything (TMA + loop only)",
         2 | 4 | 8: "all but staging", 1 | 4 | 8: "all but FFT", 1 | 2 | 8: "all but mel"}
for skip, name in names.items():
    t = timeit(skip)
    print(f"skip {skip:2d} ({name:32s}): {t:7.1f} us   delta {base - t:7.1f} us  ({(base - t) / base * 100:5.1f} %)")

print("fused utterance CMVN (last arriver normalises the utterance):")
fb = timeit(0, fused=True)
print(f"fused full: {fb:.1f} us")
for skip, name in {64: "no normalisation", 128: "no fence", 192: "no normalisation, no fence"}.items():
    t = timeit(skip, fused=True)
    print(f"skip {skip:3d} ({name:28s}): {t:7.1f} us   delta {fb - t:7.1f} us")
for plan, _, _ in sets:
    plan.set_option("debug_skip", 0)
    plan.set_cmvn("utterance")
for i in range(6):
    p, d, o = sets[i % R]
    p.execute(d, o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    p, d, o = sets[i % R]
    p.execute(d, o)
e1.record()
torch.cuda.synchronize()
print(f"unfused utterance CMVN (3 kernels): {e0.elapsed_time(e1) / 30 * 1e3:.1f} us")

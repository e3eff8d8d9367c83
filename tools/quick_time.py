# coding: utf-8
"""Tuning aid (GPU box): device time per batch (config 2 shape) of the library selected by JS2T_LIB.
   JS2T_LIB=build/libjs2t_X.so python tools/quick_time.py [label]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 4
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    sets.append((plan, packed.to_device(), plan.empty_output()))
hours = np.mean([sum(len(w) for w in synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0))]) / 16000 / 3600


def timeit(mode, unfused, n=40, reps=3):
    best = 1e9
    for plan, _, _ in sets:
        plan.set_cmvn(mode)
    for _ in range(reps):
        for i in range(8):
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
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("JS2T_LIB", "product")
raw = timeit("none", 1)
unf = timeit("utterance", 1)
fus = timeit("utterance", 0)
print(f"{label:28s} raw fbank {raw:7.1f} us | CMVN unfused {unf:7.1f} us ({hours / unf * 1e6:6.0f} h/s) | "
      f"CMVN fused {fus:7.1f} us ({hours / fus * 1e6:6.0f} h/s)")

# coding: utf-8
"""Tuning aid (GPU box): raw fbank kernel time on the config-2 batch with the persistent grid capped
(option "max_ctas"): 1 CTA per SM vs 2 tells how latency-bound a single CTA is."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 4
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("none")
    sets.append((plan, packed.to_device(), plan.empty_output()))


def timeit(n=40):
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
    return e0.elapsed_time(e1) / n * 1e3


for ctas in (74, 148, 222, 296):
    for plan, _, _ in sets:
        plan.set_option("max_ctas", ctas)
    print(f"max_ctas {ctas:4d}: raw fbank {timeit():7.1f} us")

# coding: utf-8
"""Tuning aid (GPU box): raw fbank kernel time vs number of persistent CTAs (option "max_ctas")."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 3
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("none")
    sets.append((plan, packed.to_device(), plan.empty_output()))
for n in (296, 222, 148, 74):
    for plan, _, _ in sets:
        plan.set_option("max_ctas", n)
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
    t = e0.elapsed_time(e1) / 30 * 1e3
    print(f"max_ctas {n:4d}: {t:7.1f} us   CTA-us per tile {t * n / sets[0][0].total_frames * 32:6.2f}")

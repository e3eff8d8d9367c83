# coding: utf-8
"""Tuning aid (GPU box): fbank kernel time (library-side CUDA events) per kernel mode on one workload."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402
from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch  # noqa: E402

R = 3
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    sets.append((plan, packed.to_device(), plan.empty_output()))
mean, istd = np.full(80, 5.0), np.full(80, 0.25)


def run(label, setup, n=30):
    for plan, _, _ in sets:
        setup(plan)
        plan.enable_profiling(n // R + 1)
    for i in range(6):
        p, d, o = sets[i % R]
        p.execute(d, o)
    for plan, _, _ in sets:
        plan.enable_profiling(n // R + 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        p, d, o = sets[i % R]
        p.execute(d, o)
    e1.record()
    torch.cuda.synchronize()
    kt = np.concatenate([p.kernel_times_ms(n // R + 1) for p, _, _ in sets]).mean() * 1e3
    print(f"{label:44s} step {e0.elapsed_time(e1) / n * 1e3:7.1f} us   fbank kernel {kt:7.1f} us")


def masks(plan, value):
    np.random.seed(7)
    t, nf, nt = mask_tables_for_batch(SpecAugment(2, 27, 2, 100, 1.0), plan.n_frames)
    plan.set_masks(t, nf, nt, mask_value=value)


run("raw (mode 0, no stats)", lambda p: (p.set_masks(None), p.set_cmvn("none")))
run("stats only (mode 0 + finalize)", lambda p: (p.set_masks(None), p.set_cmvn("stats")))
run("utterance CMVN unfused", lambda p: (p.set_masks(None), p.set_cmvn("utterance")))
run("utterance CMVN in-kernel finalize (mode 1)", lambda p: (p.set_masks(None), p.set_cmvn("utterance")))
run("global CMVN epilogue (mode 2), no masks", lambda p: (p.set_masks(None), p.set_cmvn("global"), p.set_global_stats(mean, istd)))
run("global CMVN epilogue (mode 2), const-fill masks", lambda p: (p.set_cmvn("global"), p.set_global_stats(mean, istd), masks(p, 0.0)))
run("utterance CMVN unfused + masks", lambda p: (p.set_cmvn("utterance"), masks(p, None)))

# coding: utf-8
"""Timing probe (GPU box, -DJS2T_PROBE_STATIC=1 build): what would the SCHEDULE of "one thread-block cluster per
utterance" (SURVEY H3 option 1: log-mel of an utterance kept in the cluster's distributed shared memory, statistics
reduced over DSMEM, HBM written once) cost the fbank kernel?  The kernel itself is unchanged (it still writes raw
rows); only who processes which tile changes: a static schedule in which a group of `cs` CTAs takes one utterance
at a time and CTA r of the group its tiles r, r + cs, ...  A lower bound of that design's cost: no cluster barrier,
no DSMEM reduction, no normalisation out of shared memory, and the 30 KB per CTA it needs for the retained tiles
are not taken away.  Usage:  JS2T_LIB=build/libjs2t_static.so python tools/cluster_probe.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 4
batches = [synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0) for r in range(R)]
packs = [frontend.PackedPCM(w) for w in batches]
devs = [p.to_device() for p in packs]


def timeit(order, slots=False, n=40, reps=3):
    if order:
        os.environ["JS2T_PROBE_TILE_ORDER"] = str(order)
    else:
        os.environ.pop("JS2T_PROBE_TILE_ORDER", None)
    if slots:
        os.environ["JS2T_PROBE_CLUSTER_SLOTS"] = "1"
    else:
        os.environ.pop("JS2T_PROBE_CLUSTER_SLOTS", None)
    plans = []
    for p in packs:
        plan = frontend.Plan(p.n_samples, p.byte_off, p.is_f32)
        plan.set_cmvn("none")
        plans.append((plan, plan.empty_output()))
    best = 1e9
    for _ in range(reps):
        for i in range(8):
            plans[i % R][0].execute(devs[i % R], plans[i % R][1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            plans[i % R][0].execute(devs[i % R], plans[i % R][1])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


lib = os.environ.get("JS2T_LIB", "product")
tiles = [int(np.sum((1 + (np.asarray(p.n_samples) - 400) // 160 + 31) // 32)) for p in packs]
print(f"{lib}: config-2 batches, {np.mean(tiles):.0f} tiles per batch, {np.mean(tiles) / 256:.1f} per utterance; raw fbank kernel, us per launch")
if "static" not in lib:
    print(f"  dynamic tile claims (product)                                   {timeit(0):7.1f}")
else:
    print(f"  static round-robin of the product's processing order            {timeit(1):7.1f}")
    for cs in (16, 8, 4, 2):
        print(f"  static, one group of {cs:2d} CTAs per utterance                      {timeit(cs):7.1f}"
              f"   (every CTA of a group padded to the same number of tile slots per utterance: {timeit(cs, True):7.1f})")

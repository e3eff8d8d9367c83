# coding: utf-8
"""Tuning aid (GPU box): how many fbank CTAs are busy per SM while the side kernel runs next to it?
Per-tile %globaltimer stamps + %smid of the fbank kernel (option "debug_times"): busy CTA-time per SM / span.
   JS2T_LIB=build/libjs2t_X.so python tools/corun_occupancy.py [label]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
packed = frontend.PackedPCM(waves)
plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
plan.set_cmvn("none")
plan.set_option("debug_times", 1)
dev = packed.to_device()
out = plan.empty_output()
plan2 = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
plan2.set_cmvn("stats")
o2 = plan2.execute(dev, plan2.empty_output())
plan2.set_global_stats(np.zeros(80), np.ones(80))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def measure(with_side):
    torch.cuda.synchronize()
    for i in range(6):
        with torch.cuda.stream(sa):
            plan.execute(dev, out)
        if with_side:
            with torch.cuda.stream(sb):
                plan2.normalize(o2)
                plan2.normalize(o2)
    torch.cuda.synchronize()
    t = plan.debug_times().astype(np.int64)
    ok = t[:, 1] > 0
    t = t[ok]
    span = (t[:, 1].max() - t[:, 0].min()) / 1e3
    sm = t[:, 2]
    busy = np.array([(t[sm == s, 1] - t[sm == s, 0]).sum() / 1e3 for s in np.unique(sm)])
    conc = busy / span
    per_tile = (t[:, 1] - t[:, 0]) / 1e3
    print(f"  side kernel {'ON ' if with_side else 'off'}: span {span:7.1f} us, SMs seen {len(busy)}, busy fbank CTAs per SM: mean {conc.mean():.2f} "
          f"min {conc.min():.2f} p10 {np.percentile(conc, 10):.2f} p50 {np.percentile(conc, 50):.2f} max {conc.max():.2f}; "
          f"tile time mean {per_tile.mean():.2f} us p50 {np.percentile(per_tile, 50):.2f} p90 {np.percentile(per_tile, 90):.2f}")


print(sys.argv[1] if len(sys.argv) > 1 else os.environ.get("JS2T_LIB", "product"))
measure(False)
measure(True)

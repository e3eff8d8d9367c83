# coding: utf-8
"""Tuning aid (GPU box): does the apply kernel run NEXT TO the persistent fbank kernel?  Stream A: raw fbank
launches; stream B: apply-only launches (js2t_normalize_execute with global statistics) on other buffers.
Times A alone, B alone and both together (perfect overlap = max, no overlap = sum).
   JS2T_LIB=build/libjs2t_X.so python tools/corun_time.py [label]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

R = 4
fb, ap = [], []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("none")
    fb.append((plan, packed.to_device(), plan.empty_output()))
    plan2 = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan2.set_cmvn("stats")
    o2 = plan2.execute(fb[-1][1], plan2.empty_output())  # normalize_execute follows a statistics pass
    plan2.set_global_stats(np.zeros(80), np.ones(80))
    ap.append((plan2, o2))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()


def run(do_a, do_b, n=40):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        sa.wait_event(e0)
        sb.wait_event(e0)
        for i in range(n):
            if do_a:
                p, d, o = fb[i % R]
                with torch.cuda.stream(sa):
                    p.execute(d, o)
            if do_b:
                p2, o2 = ap[i % R]
                with torch.cuda.stream(sb):
                    p2.normalize(o2)
        for s in (sa, sb):
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


run(True, True, 8)
for p_, _, _ in fb:
    p_.enable_profiling(64)
label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("JS2T_LIB", "product")
a, b, ab = run(True, False), run(False, True), run(True, True)
ka = np.concatenate([p_.kernel_times_ms(64)[0:30] for p_, _, _ in fb]) * 1e3   # profiled calls 0..29: fbank alone
kt = np.concatenate([p_.kernel_times_ms(64)[30:60] for p_, _, _ in fb]) * 1e3  # calls 30..59: the co-run
print(f"{label:28s} per iteration: fbank alone {a:7.1f} us | apply alone {b:7.1f} us | both streams {ab:7.1f} us "
      f"(sum {a + b:6.1f}, max {max(a, b):6.1f}) | fbank kernel alone {ka.mean():6.1f} us, during the co-run: mean {kt.mean():6.1f} us, min {kt.min():6.1f}, max {kt.max():6.1f}")

# coding: utf-8
"""Tuning aid (GPU box): device time per config-2 step when consecutive steps alternate between TWO
streams (finalize/apply of step n can then run under the fbank kernel of step n+1, if they fit next to
its two resident CTAs), against the same steps on one stream.
   JS2T_LIB=build/libjs2t_X.so python tools/two_stream_time.py [label]"""
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
    plan.set_cmvn("utterance")
    sets.append((plan, packed.to_device(), plan.empty_output()))
hours = sum(len(w) for w in synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)) / 16000 / 3600


def timeit(n_streams, n=40, reps=3, prio=False, pipelined=None):
    for p_, _, _ in sets:
        p_.set_pipelined(n_streams > 1 if pipelined is None else pipelined)
    streams = [torch.cuda.Stream(priority=(-1 if (prio and i == 1) else 0)) for i in range(n_streams)]
    best = 1e9
    main = torch.cuda.current_stream()
    for _ in range(reps):
        for i in range(8):
            p, d, o = sets[i % R]
            with torch.cuda.stream(streams[i % n_streams]):
                p.execute(d, o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for i in range(n):
            p, d, o = sets[i % R]
            with torch.cuda.stream(streams[i % n_streams]):
                p.execute(d, o)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("JS2T_LIB", "product")
if len(sys.argv) > 2:  # timing probes: 0x100 no apply, 0x200 no finalize
    for p_, _, _ in sets:
        p_.set_option("debug_skip", int(sys.argv[2], 0))
one = timeit(1)
res = [timeit(k) for k in (2, 3, 4)]
np2 = timeit(2, pipelined=False)
print(f"{label:28s} utterance CMVN step: 1 stream {one:7.1f} us ({hours / one * 1e6:6.0f} h/s) | pipelined plans on 2 / 3 / 4 streams "
      + " / ".join(f"{t:6.1f}" for t in res) + f" us ({hours / min(res) * 1e6:6.0f} h/s) | 2 streams, plans not pipelined {np2:6.1f} us")

# coding: utf-8
"""Tuning aid (GPU box): completion time of every step of a pipelined run (plans on S streams), and the start /
end of every fbank kernel (library events) — to see where a pipelined region stalls.
   python tools/pipeline_trace.py [S] [steps]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
LIMIT = int(sys.argv[3]) if len(sys.argv) > 3 else 1
CTAS = int(sys.argv[4]) if len(sys.argv) > 4 else 0
SKIP = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
R = 4
sets = []
for r in range(R):
    waves = synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0)
    packed = frontend.PackedPCM(waves)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
    plan.set_cmvn("utterance")
    plan.set_pipelined(S > 1)
    plan.set_option("side_limit", LIMIT)
    plan.set_option("side_ctas", CTAS)
    plan.set_option("debug_skip", SKIP)
    sets.append((plan, packed.to_device(), [plan.empty_output() for _ in range(6)]))
streams = [torch.cuda.Stream() for _ in range(S)]
main = torch.cuda.current_stream()
print(f'side_limit {LIMIT} side_ctas {CTAS or "default"} skip {SKIP:#x}')
for rep in range(3):
    for i in range(8):
        p, d, o = sets[i % R]
        with torch.cuda.stream(streams[i % R % S]):
            p.execute(d, o[i % 6])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    done = []
    for i in range(N):
        p, d, o = sets[i % R]
        st = streams[i % R % S]
        with torch.cuda.stream(st):
            p.execute(d, o[i % 6])
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(st)
            done.append(ev)
    torch.cuda.synchronize()
    t = np.array([e0.elapsed_time(ev) * 1e3 for ev in done])
    order = np.sort(t)
    gaps = np.diff(np.concatenate([[0.0], order]))
    print(f"rep {rep}: {N} steps on {S} streams: total {order[-1]:8.1f} us = {order[-1] / N:6.1f} us per step; completion gaps "
          f"p50 {np.percentile(gaps, 50):6.1f} p90 {np.percentile(gaps, 90):6.1f} max {gaps.max():7.1f}")
    if "-v" in sys.argv:
        print("   completion times (us):", " ".join(f"{x:.0f}" for x in order[:24]), "...")

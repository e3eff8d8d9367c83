# coding: utf-8
"""Tuning aid: per-tile time stamps of the persistent fbank kernel (option "debug_times").
   python tools/tile_timeline.py [--unfused]   (GPU box)"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
packed = frontend.PackedPCM(waves)
plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
mode = sys.argv[sys.argv.index("--mode") + 1] if "--mode" in sys.argv else "utterance"
if mode == "global":
    plan.set_cmvn("global")
    plan.set_global_stats(np.full(80, 5.0), np.full(80, 0.25))
else:
    plan.set_cmvn(mode)
plan.set_option("debug_times", 1)
dev = packed.to_device()
out = plan.empty_output()
outs = [plan.empty_output() for _ in range(3)]
for i in range(7):
    plan.execute(dev, outs[i % 3])
torch.cuda.synchronize()
t = plan.debug_times().astype(np.int64)
n = t.shape[0]
G = 296
t0 = t[:, 0].min()
rel = (t - t0) / 1e3  # us
print("tiles", n, "kernel span us", (t.max() - t0) / 1e3)
compute = rel[:, 1] - rel[:, 0]
publish = rel[:, 2] - rel[:, 1]
norm = rel[:, 3] - rel[:, 2]
valid = t[:, 3] > 0
for name, v in (("compute", compute), ("publish", publish[valid]), ("wait+normalize", norm[valid])):
    if v.size == 0:
        continue
    print(f"{name:15s} mean {v.mean():7.2f} p50 {np.percentile(v,50):7.2f} p90 {np.percentile(v,90):7.2f} "
          f"p99 {np.percentile(v,99):7.2f} max {v.max():7.2f}")
rounds = n // G
print("round: start-time spread over CTAs (min / median / max, us) and mean wait")
for r in range(0, rounds, 3):
    s = rel[r * G:(r + 1) * G, 0]
    w = norm[r * G:(r + 1) * G]
    print(f"  r{r:2d}: {s.min():7.1f} {np.median(s):7.1f} {s.max():7.1f}   wait {w.mean():6.2f}")
# per-CTA end time
ends = np.array([rel[c::G, 3 if valid.any() else 1].max() for c in range(G)])
print("CTA end time: min %.1f median %.1f max %.1f" % (ends.min(), np.median(ends), ends.max()))
print("CTA end-time deciles:", np.round(np.percentile(ends, np.arange(0, 101, 10)), 1))
order = np.argsort(ends)
print("slowest CTAs:", order[-12:], np.round(ends[order[-12:]], 1))
print("fastest CTAs:", order[:12], np.round(ends[order[:12]], 1))
print("corr(end[c], end[c+148]) =", np.corrcoef(ends[:148], ends[148:])[0, 1])
per_cta_compute = np.array([compute[c::G].mean() for c in range(G)])
print("per-CTA mean compute: min %.2f median %.2f max %.2f" % (per_cta_compute.min(), np.median(per_cta_compute), per_cta_compute.max()))
gaps = np.array([np.diff(rel[c::G, 0]).mean() for c in range(G)])
print("per-CTA mean iteration period: min %.2f median %.2f max %.2f" % (gaps.min(), np.median(gaps), gaps.max()))
# is slowness persistent?  first-half vs second-half period per CTA
h = rounds // 2
a = np.array([np.diff(rel[c::G, 0])[:h].mean() for c in range(G)])
b = np.array([np.diff(rel[c::G, 0])[h:].mean() for c in range(G)])
print("corr(period first half, second half) =", np.corrcoef(a, b)[0, 1])

if "--unfused" in sys.argv:
    smid = t[:G, 2]
    cnt = np.bincount(smid, minlength=148)
    print("CTAs per SM histogram:", np.bincount(cnt))
    per_sm_end = {}
    for c in range(G):
        per_sm_end.setdefault(int(smid[c]), []).append((c, round(float(ends[c]), 1)))
    solo = [v for v in per_sm_end.values() if len(v) == 1]
    print("SMs with one CTA:", len(solo), solo[:8])
    duo = [v for v in per_sm_end.values() if len(v) == 2]
    print("SMs with two CTAs:", len(duo), duo[:8])
    tri = [v for v in per_sm_end.values() if len(v) > 2]
    print("SMs with >2 CTAs:", len(tri), tri[:4])

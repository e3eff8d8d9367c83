import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from joeys2t_b200 import frontend, synthetic
batches = [synthetic.pooled_batch(256, seed=1234 + r, lo=10.0, hi=15.0) for r in range(4)]
cm = dict(norm_means=True, norm_vars=True, before=True)
for i in range(6):
    frontend.fbank_cmvn_specaug_ragged(batches[i % 4], cmvn=cm, layout="ragged")
torch.cuda.synchronize()
print("---- timed ----", file=sys.stderr)
t0 = time.perf_counter()
for i in range(12):
    t1 = time.perf_counter()
    out, _ = frontend.fbank_cmvn_specaug_ragged(batches[i % 4], cmvn=cm, layout="ragged")
    print(f"call {i}: {1e3*(time.perf_counter()-t1):.2f} ms", file=sys.stderr)
torch.cuda.synchronize()
print(f"total {1e3*(time.perf_counter()-t0)/12:.2f} ms per batch (no D2H)", file=sys.stderr)

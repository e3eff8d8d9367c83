import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from joeys2t_b200 import frontend, synthetic
waves = synthetic.pooled_batch(16, seed=1, lo=10.0, hi=15.0)
for i in range(30):
    t0=time.perf_counter()
    out, nf = frontend.fbank_cmvn_specaug_ragged(waves, cmvn={}, layout="padded")
    t1=time.perf_counter()
    print(f"call {i}: {1e6*(t1-t0):.0f} us", file=sys.stderr)
torch.cuda.synchronize()

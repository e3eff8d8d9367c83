# coding: utf-8
"""Tuning aid (GPU box, -DJS2T_DBG=1 build): per-warp arrival times at each block-wide barrier.
   JS2T_LIB=build/libjs2t_dbg.so python tools/warp_phases.py [--mode stats|none]"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import _lib, frontend, synthetic  # noqa: E402

waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
packed = frontend.PackedPCM(waves)
plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
plan.set_cmvn(sys.argv[sys.argv.index("--mode") + 1] if "--mode" in sys.argv else "stats")
plan.set_option("debug_times", 1)
dev = packed.to_device()
outs = [plan.empty_output() for _ in range(3)]
for i in range(6):
    plan.execute(dev, outs[i % 3])
torch.cuda.synchronize()
n = int(np.sum((plan.n_frames + 31) // 32))
buf = np.zeros(68 * n, np.uint64)
_lib.check(_lib.load().js2t_plan_debug_times(plan._h, buf.ctypes.data, buf.size))
w = buf[4 * n:].reshape(n, 8, 8).astype(np.int64)  # [tile][warp][slot]
w[:, :, 2] = w[:, :, 1]  # slot 1 = PCM of the tile has arrived (mbarrier); slot 2 is not stamped
ok = (w[:, :, :6] > 0).all(axis=(1, 2))
w = w[ok]
print(f"{ok.sum()} of {n} tiles with complete stamps")
names = ["top->PCM ready", "(unused)", "PCM->fft done", "fft->mel done", "mel->store done"]
# arrival of each warp at barrier s relative to the release of barrier s-1 (= max arrival over warps at s-1)
for s in (1, 3, 4, 5):
    rel = w[:, :, s] - w[:, :, s - 1].max(axis=1, keepdims=True)
    dur = rel / 1e3
    crit = dur.max(axis=1)
    print(f"{names[s - 1]:16s} phase (release->last arrival) mean {crit.mean():6.3f} us | per-warp mean arrival "
          + " ".join(f"{dur[:, k].mean():5.2f}" for k in range(8))
          + f" | mean wait of a warp at the barrier {(crit[:, None] - dur).mean():5.3f} us")
tot = (w[:, :, 5].max(axis=1) - w[:, :, 0].max(axis=1)) / 1e3
print(f"tile (top barrier release -> end barrier last arrival) mean {tot.mean():.3f} us")

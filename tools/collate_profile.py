# coding: utf-8
"""Tuning aid (GPU box): where a per-batch call of the public path spends its time — host wall time of each
step of frontend.fbank_cmvn_specaug_ragged for a 20 000-frame token batch, and the device time of its H2D copy
and kernels (CUDA events)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic, tables  # noqa: E402
from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch  # noqa: E402

waves = synthetic.pooled_batch(16, seed=1, lo=10.0, hi=15.0)
sa = SpecAugment(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=1.0)
T = {}


def tick(name, t0):
    T.setdefault(name, []).append(time.perf_counter() - t0)
    return time.perf_counter()


dev_ms = []
for it in range(40):
    t0 = time.perf_counter()
    n_frames = [tables.num_frames(len(w)) for w in waves]
    table, nf, nt = mask_tables_for_batch(sa, n_frames)
    t0 = tick("draw masks", t0)
    packed = frontend.PackedPCM(waves)
    t0 = tick("PackedPCM (own pinned alloc)", t0)
    plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32, layout="padded")
    t0 = tick("Plan()", t0)
    plan.set_cmvn("utterance", True, True, True)
    plan.set_masks(table, nf, nt, None)
    t0 = tick("set_cmvn + set_masks", t0)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    dev = packed.host[:packed.nbytes].to("cuda", non_blocking=True)
    e[1].record()
    t0 = tick("H2D enqueue", t0)
    out = plan.execute(dev)
    e[2].record()
    t0 = tick("execute enqueue", t0)
    torch.cuda.synchronize()
    t0 = tick("synchronize", t0)
    dev_ms.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    plan.close()
    t0 = tick("plan.close", t0)
print(f"batch: {len(waves)} utterances, {sum(n_frames)} frames, {packed.nbytes / 1e6:.1f} MB PCM")
for k, v in T.items():
    print(f"{k:32s} {np.median(v[5:]) * 1e6:8.1f} us (median)")
d = np.array(dev_ms[5:])
print(f"device: H2D {np.median(d[:, 0]) * 1e3:.1f} us, kernels {np.median(d[:, 1]) * 1e3:.1f} us")

# the one-call entry point, asynchronous return
for _ in range(10):
    frontend.fbank_cmvn_specaug_ragged(waves, cmvn={}, masks=table, n_fmask=nf, n_tmask=nt, layout="padded")
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 200
for _ in range(N):
    frontend.fbank_cmvn_specaug_ragged(waves, cmvn={}, masks=table, n_fmask=nf, n_tmask=nt, layout="padded")
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
audio_h = sum(len(w) for w in waves) / 16000 / 3600
print(f"fbank_cmvn_specaug_ragged: {(t1 - t0) / N * 1e6:.0f} us per call to enqueue, {(t2 - t0) / N * 1e6:.0f} us per call incl. final "
      f"sync = {audio_h * N / (t2 - t0):.0f} audio-h/s")

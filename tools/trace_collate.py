# coding: utf-8
"""Tuning aid (GPU box): host wall time per phase of the collate path (SpeechBatchCollator -> process_batch ->
js2t_batch_fbank) on the token batches of tools/bench_next_rows.py.  Run with JS2T_BATCH_TRACE=1 to get the C
phases on stderr; this script prints the Python-side split."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic, tables  # noqa: E402
from joeys2t_b200.batching import FrameCountBatchSampler, SpeechBatchCollator  # noqa: E402
from joeys2t_b200.speech_processor import SpeechProcessor  # noqa: E402

waves = synthetic.pooled_batch(256, seed=1, lo=10.0, hi=15.0)
n_frames = np.array([tables.num_frames(len(w)) for w in waves])
proc = SpeechProcessor(level="frame", num_freq=80, normalize=False, max_length=-1, min_length=1,
                       specaugment=dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=1.0),
                       cmvn=dict(norm_means=True, norm_vars=True, before=True))
sampler = FrameCountBatchSampler(range(256), 20000, "token", n_frames=n_frames, max_length=3000, is_train=True)
collate = SpeechBatchCollator(proc, lambda i: waves[i], is_train=True)
batches = [b for b in sampler]
np.random.seed(1)
for _ in range(3):
    for b in batches:
        collate(b)
torch.cuda.synchronize()
print("---- timed ----", file=sys.stderr)
per = []
t00 = time.perf_counter()
for rep in range(3):
    for b in batches:
        t0 = time.perf_counter()
        collate(b)
        per.append(time.perf_counter() - t0)
t_enq = time.perf_counter() - t00
torch.cuda.synchronize()
t_all = time.perf_counter() - t00
per = np.array(per) * 1e6
print(f"collate call: median {np.median(per):.0f} us, mean {per.mean():.0f} us, max {per.max():.0f} us; "
      f"enqueue loop {t_enq * 1e3:.2f} ms, with final sync {t_all * 1e3:.2f} ms for {len(per)} batches "
      f"({t_all / len(per) * 1e6:.0f} us per batch)")
# the same batches through the frontend call alone (no sampler / collator / mask draws)
lists = [[waves[i] for i in b] for b in batches]
for ws in lists:
    frontend.fbank_cmvn_specaug_ragged(ws, cmvn={}, layout="padded")
torch.cuda.synchronize()
t00 = time.perf_counter()
for rep in range(3):
    for ws in lists:
        frontend.fbank_cmvn_specaug_ragged(ws, cmvn={}, layout="padded")
t_enq = time.perf_counter() - t00
torch.cuda.synchronize()
t_all = time.perf_counter() - t00
print(f"frontend call alone: enqueue loop {t_enq / 54 * 1e6:.0f} us per batch, with final sync {t_all / 54 * 1e6:.0f} us per batch")
# device-side floor: the H2D transfers of these batches back to back from pinned memory
pin = [torch.empty(sum(w.nbytes for w in ws), dtype=torch.uint8).pin_memory() for ws in lists]
dev = [torch.empty_like(p, device="cuda") for p in pin]
torch.cuda.synchronize()
t00 = time.perf_counter()
for rep in range(3):
    for p, d in zip(pin, dev):
        d.copy_(p, non_blocking=True)
torch.cuda.synchronize()
t_all = time.perf_counter() - t00
mb = sum(p.numel() for p in pin) / len(pin) / 1e6
print(f"raw pinned H2D of the same batches ({mb:.1f} MB each): {t_all / 54 * 1e6:.0f} us per batch = {mb / (t_all / 54) / 1e3:.1f} GB/s")

# coding: utf-8
"""Tuning aid (GPU box): pinned host -> device copy time against the copy size (CUDA events, warm), i.e. the floor
of the per-batch public path, whose batches are a few MB each."""
import torch

sizes_mb = [0.5, 1, 2, 4, 5.4, 8, 16, 32, 64, 102]
s = torch.cuda.Stream()
for mb in sizes_mb:
    n = int(mb * 1e6)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(s):
        for _ in range(5):
            dev.copy_(host, non_blocking=True)
        s.synchronize()
        reps = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        e1.record(s)
        s.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{mb:6.1f} MB: {us:8.1f} us per copy = {n / us / 1e3:6.1f} GB/s")

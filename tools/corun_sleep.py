# coding: utf-8
"""Tuning aid (GPU box): fbank kernel time while a generic side kernel (tools/microbench_corun_side.cu) is
resident on every SM.  Usage: python tools/corun_sleep.py"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic  # noqa: E402

lib = ctypes.CDLL(str(Path(__file__).resolve().parent.parent / "build" / "libcorun_side.so"))
lib.launch_side.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                            ctypes.c_longlong, ctypes.c_void_p]
waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
packed = frontend.PackedPCM(waves)
plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
plan.set_cmvn("none")
dev = packed.to_device()
out = plan.empty_output()
buf = torch.zeros(64 * 1024 * 1024 + 64, dtype=torch.float32, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(5):
    plan.execute(dev, out)
torch.cuda.synchronize()


def run(label, grid, threads, smem, mode, each_us=0):
    torch.cuda.synchronize()
    n = 10
    if grid > 0:
        with torch.cuda.stream(sb):
            if each_us == 0:
                lib.launch_side(grid, threads, smem, mode, int(n * 400e3), buf.data_ptr(), 16 * 1024 * 1024, sb.cuda_stream)
            else:  # many short side kernels back to back
                for _ in range(int(n * 300 / each_us)):
                    lib.launch_side(grid, threads, smem, mode, int(each_us * 1e3), buf.data_ptr(), 16 * 1024 * 1024, sb.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sa):
        plan.execute(dev, out)  # the side kernel is resident by now
        e0.record(sa)
        for _ in range(n):
            plan.execute(dev, out)
        e1.record(sa)
    torch.cuda.synchronize()
    moved = int(buf[64 * 1024 * 1024:].view(torch.int64)[0].item()) if grid > 0 else 0
    dur = n * 400.0 if each_us == 0 else each_us
    print(f"{label:60s} fbank {e0.elapsed_time(e1) / n * 1e3:7.1f} us   side kernel: {moved / 1e6 / dur * 1e6 / 1e6:6.2f} TB/s each way (elements touched)")


run("alone", 0, 0, 0, 0)
run("side: 148 x 128 thr, 0 smem, 8 loads in flight + stores", 148, 128, 0, 4)
run("side: 148 x 128 thr, 0 smem, 8 loads in flight, no stores", 148, 128, 0, 5)
run("side: 148 x 32 thr, 0 smem, 8 loads in flight + stores", 148, 32, 0, 4)
run("side: 148 x 128 thr, 0 smem, 1 load in flight + store", 148, 128, 0, 2)
# the library's own apply kernel (whatever variant the build defaults to) as the side kernel
plan2 = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32)
plan2.set_cmvn("stats")
o2 = plan2.execute(dev, plan2.empty_output())
plan2.set_global_stats(np.zeros(80), np.ones(80))
torch.cuda.synchronize()


def run_apply(label, k):
    torch.cuda.synchronize()
    n = 10
    with torch.cuda.stream(sb):
        for _ in range(k):
            plan2.normalize(o2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sa):
        plan.execute(dev, out)
        e0.record(sa)
        for _ in range(n):
            plan.execute(dev, out)
        e1.record(sa)
    t_b = torch.cuda.Event(enable_timing=True)
    t_b.record(sb)
    torch.cuda.synchronize()
    print(f"{label:60s} fbank {e0.elapsed_time(e1) / n * 1e3:7.1f} us   (stream B ended {e0.elapsed_time(t_b) * 1e3:8.1f} us after the timed region began; it lasted {e0.elapsed_time(e1) * 1e3:8.1f} us)")


run_apply("side: library apply kernel x 30", 30)
run_apply("side: library apply kernel x 12", 12)
run("alone again", 0, 0, 0, 0)

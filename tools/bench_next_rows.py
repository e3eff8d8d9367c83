# coding: utf-8
"""Measurement of the rows either side of the hot path (SURVEY.md §8 f-2 ... f-4) on one B200:
   python tools/bench_next_rows.py > profiles/<round>_next_rows.txt
 * f-4 ingest kernels: achieved GB/s of the 48 kHz -> 16 kHz conversion against the HBM peak, next to
       the numpy expression of the reference on one host core;
 * f-2 collate path: host PCM of a token-sized batch -> (B, Tmax, 80) on the device, per batch;
 * f-3 feature store: batched extraction of a corpus into the npy-in-zip archive (tmpfs)."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from joeys2t_b200 import feature_store, frontend, synthetic  # noqa: E402
from joeys2t_b200.batching import FrameCountBatchSampler, SpeechBatchCollator  # noqa: E402
from joeys2t_b200.speech_processor import SpeechProcessor  # noqa: E402
from joeys2t_b200 import tables  # noqa: E402

peaks = ROOT / "MEASURED_PEAKS.json"
hbm = json.loads(peaks.read_text()).get("hbm_gbs", 6650.0) if peaks.is_file() else 6650.0
src_peak = "MEASURED_PEAKS.json" if peaks.is_file() else "fallback of B200_PROFILING.md"


def ev_time(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n  # ms


print(f"# next rows, measured on {torch.cuda.get_device_name(0)}; HBM peak {hbm:.0f} GB/s ({src_peak})")

# ---- f-4 ------------------------------------------------------------------------------------------
print("\n## f-4  reformat_freq 48 kHz -> 16 kHz (scripts/gradio_demo.py:35-45), PCM resident in HBM")
rng = np.random.default_rng(0)
for dtype, secs in ((np.int16, 3600), (np.float32, 1800)):
    n = 48000 * secs
    host = (rng.integers(-20000, 20000, n)).astype(dtype)
    # two inputs larger than L2 in total so that consecutive launches do not hit in cache
    devs = [torch.from_numpy(host).cuda(), torch.from_numpy(host[::-1].copy()).cuda()]
    outs = [torch.empty(n // 3, dtype=torch.int16, device="cuda") for _ in devs]
    it = [0]

    def step():
        i = it[0] & 1
        frontend.reformat_48k_to_16k(devs[i], outs[i])
        it[0] += 1

    ms = ev_time(step)
    item = np.dtype(dtype).itemsize
    alg = n * item * 2 + (n // 3) * 2  # the maximum needs its own pass: 2 reads of the input + 1 write
    t0 = time.perf_counter()
    sample = host[: 48000 * 60]
    with np.errstate(all="ignore"):  # the reference's numpy expression (scripts/gradio_demo.py:42-43), timed on one core
        ref = ((sample / max(np.max(sample), 1)) * 32767).reshape((-1, 3)).mean(axis=1).astype("int16")
    cpu_s = time.perf_counter() - t0
    got = frontend.reformat_48k_to_16k(torch.from_numpy(sample).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    print(f"{np.dtype(dtype).name:8s} {secs / 3600:.1f} audio-h per launch: {ms * 1e3:8.1f} us  "
          f"{alg / ms / 1e6:7.0f} GB/s algorithmic = {alg / ms / 1e6 / hbm * 100:4.1f} % of HBM peak "
          f"({secs / 3600 / (ms / 1e3):9.0f} audio-h/s)  |  numpy on one host core: "
          f"{60 / 3600 / cpu_s:6.2f} audio-h/s (1-minute sample, bit-identical output)")
    del devs, outs

# ---- f-2 ------------------------------------------------------------------------------------------
print("\n## f-2  sampler + collator: host int16 PCM -> (B, Tmax, 80) fp32 on the device (utterance CMVN + SpecAugment)")
waves = synthetic.pooled_batch(256, seed=1234, lo=10.0, hi=15.0)
n_frames = np.array([tables.num_frames(len(w)) for w in waves])
proc = SpeechProcessor(level="frame", num_freq=80, max_length=3000,
                       specaugment=dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=1.0),
                       cmvn=dict(norm_means=True, norm_vars=True, before=True))
sampler = FrameCountBatchSampler(range(256), 20000, "token", n_frames=n_frames, max_length=3000, is_train=True)
collate = SpeechBatchCollator(proc, lambda i: waves[i], is_train=True)
t0 = time.perf_counter()
batches = [b for b in sampler]
t_sampler = time.perf_counter() - t0
np.random.seed(1)
for b in batches[:2]:
    collate(b)
torch.cuda.synchronize()
for b in batches:  # every staging slot and pool buffer has been allocated once
    collate(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
audio = 0.0
for rep in range(3):
    for b in batches:
        src, ln, _ = collate(b)
        audio += sum(len(waves[i]) for i in b) / 16000
torch.cuda.synchronize()
dt = (time.perf_counter() - t0)
batches_timed = 3 * len(batches)
# host-side gather alone (js2t_pack_pcm, pinned destination): pool of copy threads vs one thread
import ctypes  # noqa: E402
from joeys2t_b200 import _lib  # noqa: E402
b0 = batches[0]
arrs = [waves[i] for i in b0]
sizes = np.array([a.nbytes for a in arrs], np.int64)
offs = np.concatenate([[0], np.cumsum((sizes + 15) // 16 * 16)[:-1]]).astype(np.int64)
dstbuf = torch.empty(int(offs[-1] + sizes[-1] + 16), dtype=torch.uint8, pin_memory=True)
ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
pack_us = {}
for thr in (1, 0):
    for _ in range(5):
        _lib.load().js2t_pack_pcm(len(arrs), ptrs, sizes.ctypes.data, offs.ctypes.data, dstbuf.data_ptr(), dstbuf.numel(), thr)
    t0 = time.perf_counter()
    for _ in range(50):
        _lib.load().js2t_pack_pcm(len(arrs), ptrs, sizes.ctypes.data, offs.ctypes.data, dstbuf.data_ptr(), dstbuf.numel(), thr)
    pack_us[thr] = (time.perf_counter() - t0) / 50 * 1e6
print(f"js2t_pack_pcm of one batch ({sizes.sum() / 1e6:.1f} MB into pinned memory): {pack_us[1]:.0f} us on one thread, "
      f"{pack_us[0]:.0f} us with the copy pool")
print(f"token batches of 20000 frames (librispeech_100h.yaml:83-85): {len(batches)} batches from 256 utterances; "
      f"sampler {t_sampler * 1e3:.2f} ms total (no audio touched); collate {dt / batches_timed * 1e3:.2f} ms per batch wall "
      f"= {audio / 3600 / dt:7.1f} audio-h/s through the per-batch Python path (pack + H2D + 3 kernels)")

import cProfile  # noqa: E402
import pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
for b in batches:
    collate(b)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

# ---- f-3 ------------------------------------------------------------------------------------------
print("\n## f-3  extract_corpus: 256 utterances (0.89 audio-h) -> fbank80.zip (npy-in-ZIP_STORED) on tmpfs")
with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
    items = [(f"utt{i:05d}", w) for i, w in enumerate(waves)]
    feature_store.extract_corpus(items[:8], Path(d) / "warm.zip")
    t0 = time.perf_counter()
    manifest, frames, failed = feature_store.extract_corpus(items, Path(d) / "fbank80.zip")
    dt = time.perf_counter() - t0
    size = (Path(d) / "fbank80.zip").stat().st_size
    hours = sum(len(w) for w in waves) / 16000 / 3600
    print(f"{dt * 1e3:.0f} ms wall for {hours:.3f} audio-h -> {size / 1e6:.0f} MB archive: {hours / dt:7.1f} audio-h/s "
          f"(GPU fbank + D2H + npy serialisation + zip write; {len(failed)} failed)")

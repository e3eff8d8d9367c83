# coding: utf-8
"""Tuning aid (GPU box): wall time of the reference-facing per-item calls (one utterance per call, host
numpy in / out) — extract_fbank_features, SpeechProcessor.__call__ on a wav file — and where it goes."""
import cProfile
import pstats
import sys
import time
import wave
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import helpers_for_audio as HA  # noqa: E402
from joeys2t_b200 import synthetic  # noqa: E402
from joeys2t_b200.speech_processor import SpeechProcessor  # noqa: E402

waves = synthetic.pooled_batch(32, seed=1, lo=10.0, hi=15.0)
tens = [torch.from_numpy(w.astype(np.float32) / 32768.0)[None] for w in waves]
for t in tens[:4]:
    HA.extract_fbank_features(t, 16000)
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in tens:
    HA.extract_fbank_features(t, 16000)
dt = (time.perf_counter() - t0) / len(tens)
audio = np.mean([len(w) for w in waves]) / 16000
print(f"extract_fbank_features: {dt * 1e3:.3f} ms per {audio:.1f} s utterance = {audio / 3600 / dt:.2f} audio-h/s per calling thread")

tmp = Path("/dev/shm/js2t_lat")
tmp.mkdir(exist_ok=True)
for i, w in enumerate(waves):
    with wave.open(str(tmp / f"u{i}.wav"), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(w.tobytes())
proc = SpeechProcessor(level="frame", num_freq=80, max_length=3000,
                       specaugment=dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=40, time_mask_p=1.0),
                       cmvn=dict(norm_means=True, norm_vars=True, before=True))
proc.root_path = tmp
for i in range(4):
    proc(f"u{i}.wav", is_train=True)
t0 = time.perf_counter()
for i in range(len(waves)):
    proc(f"u{i}.wav", is_train=True)
dt = (time.perf_counter() - t0) / len(waves)
print(f"SpeechProcessor.__call__ (wav -> CMVN -> SpecAugment): {dt * 1e3:.3f} ms per utterance = {audio / 3600 / dt:.2f} audio-h/s")
pr = cProfile.Profile()
pr.enable()
for i in range(len(waves)):
    proc(f"u{i}.wav", is_train=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# coding: utf-8
"""Tuning aid (GPU box): where a per-item extract_fbank_features call spends its wall time."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from joeys2t_b200 import frontend, helpers_for_audio as H, synthetic
w16 = synthetic.pooled_batch(1, seed=3, lo=12.9, hi=12.9)[0]
wf = torch.from_numpy(w16.astype(np.float32) / 32768.0)[None]   # (1, N) float like torchaudio.load
for name, w in (("float (1,N) tensor", wf), ("int16 (N,) array", w16)):
    for _ in range(20):
        H.extract_fbank_features(w, 16000)
    T = {"enqueue": [], "d2h": [], "total": []}
    for _ in range(200):
        t0 = time.perf_counter()
        dev, _ = frontend.fbank_cmvn_specaug_ragged([w])
        t1 = time.perf_counter()
        f = dev.cpu().numpy()
        t2 = time.perf_counter()
        T["enqueue"].append(t1 - t0); T["d2h"].append(t2 - t1)
    for _ in range(200):
        t0 = time.perf_counter()
        H.extract_fbank_features(w, 16000)
        T["total"].append(time.perf_counter() - t0)
    print(f"{name:20s}: enqueue {np.median(T['enqueue'])*1e6:6.1f} us | .cpu().numpy() {np.median(T['d2h'])*1e6:6.1f} us | extract_fbank_features {np.median(T['total'])*1e6:6.1f} us")

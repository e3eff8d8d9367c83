import sys, time, ctypes
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import frontend, synthetic, _lib
waves = synthetic.pooled_batch(16, seed=1, lo=10.0, hi=15.0)
n = sum((w.nbytes + 15)//16*16 for w in waves)
def timed(label, make_host, pre=None, reps=20):
    ms=[]
    for _ in range(reps):
        host = make_host()
        if pre: keep = pre()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dev = host[:n].to("cuda", non_blocking=True); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(f"{label:60s} H2D {np.median(ms[3:])*1e3:8.1f} us  ({n/np.median(ms[3:])/1e6:.1f} GB/s)")
fixed = torch.empty(n, dtype=torch.uint8, pin_memory=True)
timed("fixed pinned buffer", lambda: fixed)
timed("fresh torch.empty(pin_memory=True) each time", lambda: torch.empty(n, dtype=torch.uint8, pin_memory=True))
def packed(): return frontend.PackedPCM(waves).host
timed("PackedPCM (fresh pinned + js2t_pack_pcm pool)", packed)
def packed_fixed(): return frontend.PackedPCM(waves, host=fixed).host
timed("PackedPCM into the fixed pinned buffer", packed_fixed)
p = frontend.PackedPCM(waves)
def mkplan():
    pl = frontend.Plan(p.n_samples, p.byte_off, p.is_f32, layout="padded"); return pl
timed("fixed pinned buffer, a Plan created right before the copy", lambda: fixed, pre=mkplan)
lib=_lib.load()
def pack1():
    arrs=[np.ascontiguousarray(w) for w in waves]
    sizes=np.array([a.nbytes for a in arrs],np.int64); off=np.concatenate([[0],np.cumsum((sizes+15)//16*16)[:-1]]).astype(np.int64)
    ptrs=(ctypes.c_void_p*len(arrs))(*[a.ctypes.data for a in arrs])
    lib.js2t_pack_pcm(len(arrs),ptrs,sizes.ctypes.data,off.ctypes.data,fixed.data_ptr(),fixed.numel(),1); return fixed
timed("single-thread pack into the fixed pinned buffer", pack1)

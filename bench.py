# coding: utf-8
"""
bench.py — headline benchmark of the B200 audio front-end (BASELINE.json):

    metric   log-mel+CMVN audio-hours/sec
    workload configs[1]: LibriSpeech-100h-shaped synthetic batch, 256 x 10-15 s 16 kHz int16
             utterances, 80 bins, utterance CMVN (librispeech_100h.yaml), one batch per step

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1: RANK / LOCAL_RANK / WORLD_SIZE from the environment).
Every rank processes its own batches (utterances are independent: no data-path collective, weak
scaling).  A "step" is one pass of the hot path (fbank -> utterance CMVN) over one batch.

* ``value``     whole-job audio-hours/sec with PCM resident in HBM, CUDA-event timed, max over ranks
* ``e2e``       same metric through the public host API: pinned host PCM -> H2D -> kernels -> D2H of
                the features into pinned host memory, every step inside the timed region
* ``roofline``  algorithmic HBM bytes of the dominant kernel / its CUDA-event duration (events are
                recorded inside the library on the launching stream) against MEASURED_PEAKS.json
* ``cpu_baseline``  the oracle port (numpy restatement of the reference path) on all host cores
* ``--impl reference``  times that CPU port alone and prints the same JSON line

Between timed steps the working set rotates over ``--rotate`` distinct batches so that inputs and
outputs (~205 MB per batch) exceed the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "log-mel+CMVN audio-hours/sec"
UNIT = "audio-hours/s"
SR = 16000
BYTES_PER_FRAME_I16 = 640  # 160 int16 samples in + 80 float32 out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--rotate", type=int, default=4, help="distinct batches cycled through (L2 flush)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cmvn-path", default="auto", choices=["auto", "fused", "unfused"],
                    help="utterance CMVN inside the persistent fbank kernel, or the three-kernel path")
    return ap.parse_args()


def workload_name(utts):
    return (f"librispeech_100h-shaped synthetic batch: {utts} x U(10,15) s 16 kHz int16 utterances, "
            "80-bin Kaldi fbank + utterance CMVN")


# ----------------------------------------------------------------------------------------------
# CPU leg: the oracle port on all host cores (the reference's own pattern is one process per core,
# scripts/prepare_mustc.py:53,118 num_proc=16)
# ----------------------------------------------------------------------------------------------
def _cpu_init():
    from oracle import ref_cpu_path
    ref_cpu_path.single_thread()


def _cpu_worker(w):
    from oracle import ref_cpu_path
    return ref_cpu_path.fbank_cmvn(w).shape[0]


def cpu_kind():
    from oracle import ref_cpu_path
    return ref_cpu_path.KIND


def cpu_port_throughput(waves, cores, budget_s=12.0):
    """audio-hours/sec of the reference CPU path over ``waves`` with a pool of ``cores`` processes,
    repeated until about ``budget_s`` seconds of wall time have been measured."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    hours = sum(len(w) for w in waves) / SR / 3600.0
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_worker, waves[:cores])  # warm the workers (imports, FFT plans)
        passes, t0 = 0, time.perf_counter()
        while True:
            pool.map(_cpu_worker, waves, chunksize=1)
            passes += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s:
                break
    return hours * passes / dt, dt, passes


def cpu_sample(waves, cores):
    """Bounded sample of the workload: a whole batch when it is small enough."""
    n = min(len(waves), max(cores * 8, 64))
    return waves[:n]


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # pylint: disable=broad-except
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    """``--impl reference``: the reference's CPU implementation of the path.  The reference is pure
    Python around an un-vendored torchaudio; /root/reference does not travel to the GPU box, but
    torchaudio (which holds all of the arithmetic) is part of the image, so oracle/ref_cpu_path.py
    calls the real ``torchaudio.compliance.kaldi.fbank`` with the reference's arguments plus the
    restated joeynmt glue and CMVN (falling back to the numpy restatement if torchaudio is missing),
    one process per host core like the reference's own prep scripts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from joeys2t_b200 import synthetic
    cores = os.cpu_count() or 1
    waves = synthetic.pooled_batch(args.utts, seed=1234, lo=10.0, hi=15.0)
    sample = cpu_sample(waves, cores)
    hours = sum(len(w) for w in sample) / SR / 3600.0
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, sample[:cores])
        # keep the whole run within a few minutes
        steps = args.steps
        t_all = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, sample, chunksize=1)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > 150:
                break
    dt = float(np.mean(times))
    value = hours / dt
    sample_desc = (f"{len(sample)} of the {len(waves)} utterances of one batch ({hours:.3f} audio-h) per step, "
                   f"{cores} processes x 1 thread; {cpu_kind()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.utts), "cpu_model": cpu_model()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_model():
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from joeys2t_b200 import _lib, distributed, frontend, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    rank, local_rank, world = distributed.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if _lib.is_stale():
        if local_rank == 0:
            _lib.build()
        if world > 1:
            dist.barrier()

    # ---- synthetic workload: `rotate` distinct batches per rank, resident in HBM ----------------
    R = max(1, args.rotate)
    batches, plans, pcm_dev, outs, packs = [], [], [], [], []
    for r in range(R):
        waves = synthetic.pooled_batch(args.utts, seed=1234 + 1000 * rank + r, lo=10.0, hi=15.0)
        packed = frontend.PackedPCM(waves)
        plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32, device=local_rank)
        plan.set_cmvn("utterance", True, True, True)
        plan.set_option("fused_cmvn", int(args.cmvn_path == "fused"))
        batches.append(waves)
        packs.append(packed)
        plans.append(plan)
        pcm_dev.append(packed.to_device(dev))
        outs.append(plan.empty_output())
    torch.cuda.synchronize()
    hours_per_step = [sum(len(w) for w in b) / SR / 3600.0 for b in batches]
    frames_per_step = [p.total_frames for p in plans]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(args.warmup):
        plans[i % R].execute(pcm_dev[i % R], outs[i % R])
    for p in plans:
        p.enable_profiling((args.steps + R - 1) // R)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        plans[i % R].execute(pcm_dev[i % R], outs[i % R])
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms_total = ev0.elapsed_time(ev1)
    hours_done = sum(hours_per_step[i % R] for i in range(args.steps))
    frames_done = sum(frames_per_step[i % R] for i in range(args.steps))
    kt = np.concatenate([p.kernel_times_ms((args.steps + R - 1) // R) for p in plans])
    kernel_ms = float(kt.mean())

    # ---- end to end: pinned host PCM -> device -> features -> pinned host -------------------------
    e2e = None
    if not args.no_e2e:
        # public streaming API: H2D of batch i+1 || kernels of batch i || D2H of batch i-1
        pipe = frontend.HostPipeline(n_slots=min(3, max(2, R)), max_pcm_bytes=max(p.nbytes for p in packs),
                                     max_out_rows=max(p.out_rows for p in plans), device=local_rank)

        def e2e_step(i):
            j = i % R
            slot = pipe.submit(packs[j], plans[j])      # pinned host PCM -> device -> pinned host features
            return slot

        for i in range(max(3, args.warmup)):
            e2e_step(i)
        pipe.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.s_in.wait_event(e0)
        last = 0
        for i in range(args.steps):
            last = e2e_step(i)
        checksum = float(pipe.result(last)[0, 0])       # the host really has the last batch's features
        for s_ in (pipe.s_in, pipe.s_compute, pipe.s_out):
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        assert np.isfinite(checksum)
        e2e = (hours_done, e2e_ms, int(np.mean([p.nbytes for p in packs])),
               int(np.mean([o.numel() * 4 for o in outs])))

    # ---- reduce over ranks: max time, summed work ---------------------------------------------------
    t = torch.tensor([ms_total, e2e[1] if e2e else 0.0], dtype=torch.float64, device=dev)
    w = torch.tensor([hours_done, float(frames_done)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_ms_max = t.tolist()
    hours_all, frames_all = w.tolist()

    if rank == 0:
        value = hours_all / (ms_total_max * 1e-3)
        peak, peak_src = measured_peaks()
        frames_launch = float(np.mean(frames_per_step))
        algo_bytes = frames_launch * BYTES_PER_FRAME_I16
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.utts),
                "frames_per_step_per_gpu": frames_launch,
                "audio_hours_per_step_per_gpu": float(np.mean(hours_per_step)),
                "pcm": "int16, resident in HBM", "cmvn": "utterance (norm_means, norm_vars, before)",
                "l2": f"rotating over {R} distinct batches per GPU "
                      f"(~{R * (np.mean([p.nbytes for p in packs]) + np.mean([o.numel()*4 for o in outs])) / 1e6:.0f} MB "
                      "of inputs+outputs, larger than the 126 MB L2)",
                "parallelism": f"utterance-sharded x{world}, no data-path collective",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "kernel": "fbank_tile_kernel", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": algo_bytes,
                "peak_source": peak_src,
                "kernel_share_of_step": kernel_ms / (ms_total / args.steps),
                "note": "the kernel is FP32-issue bound, not HBM bound; see DESIGN.md and profiles/",
            },
            "clocks": clk,
            "gpu_launches": (1 if args.cmvn_path == "fused" else 3) * args.steps,
        }
        if e2e:
            line["e2e"] = {"value": hours_all / (e2e_ms_max * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[2], "d2h_bytes_per_step": e2e[3]}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sample = cpu_sample(batches[0], cores)
            v, dt, passes = cpu_port_throughput(sample, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{len(sample)} of {len(batches[0])} utterances of one batch x {passes} passes, "
                          f"{dt:.1f} s wall on {cores} processes x 1 thread ({cpu_model()}); {cpu_kind()}"}
        print(json.dumps(line), flush=True)
    for p in plans:
        p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

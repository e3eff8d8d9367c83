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
                (``frontend.HostPipeline``: the caller hands over one packed pinned buffer per batch);
                ``e2e.from_pageable_arrays`` = the same from separate pageable numpy arrays, one per
                utterance, through the one public per-batch call (gather, upload, kernels, read-back),
                nothing packed or planned outside the timed region
* ``roofline``  algorithmic HBM bytes of the dominant kernel / its CUDA-event duration (events are
                recorded inside the library on the launching stream) against MEASURED_PEAKS.json.
                The step's three kernels are chained by programmatic dependent launch, which an
                event between two kernels would undo; the events are therefore recorded in a second
                region of the same ``steps`` steps right after the timed one, and that region's own
                (slower) step time is reported next to them
* ``roofline_fp32``  the same kernel against the FP32 pipe, from the ncu counters of THIS build
                (profiles/fbank_ncu_metrics.json, keyed by a hash of the kernel sources; null when
                the committed capture belongs to another build)
* ``cpu_baseline``  the reference's CPU path on all host cores: the UNMODIFIED reference
                (joeynmt.helpers_for_audio.extract_fbank_features -> joeynmt.data_augmentation.CMVN,
                installed into oracle/_ref by oracle/build_ref.sh) when available — kind
                "reference" — else torchaudio's kaldi.fbank + restated glue / CMVN — kind "port"
* ``--impl reference``  times that CPU path alone, the whole batch per step, same JSON line
* ``cfg5``      (N > 1 only) a short corpus sweep with global CMVN whose statistics are unknown: the
                path's one collective (161 x fp64 all-reduce through the C ABI's ncclAllReduce), with
                on-box checks that every rank ends up with the same statistics and that they equal a
                rank-0 float64 recomputation from all ranks' per-utterance statistics

Between timed steps the working set rotates over enough batch buffers (inputs + outputs ~205 MB per
batch) to exceed 4 GB, far beyond the 126 MB L2: ``--rotate`` distinct synthetic batches, replicated
into distinct device buffers (the kernels are data-independent; what matters for L2 is the address).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=0, help="utterances per batch (0 = the workload's own size)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config; cfg2 is the headline, the others are extra measurement rows")
    ap.add_argument("--sweep-hours", type=float, default=1000.0, help="cfg5: corpus size in audio-hours")
    ap.add_argument("--rotate", type=int, default=4, help="distinct synthetic batches generated per rank")
    ap.add_argument("--working-set-gb", type=float, default=4.2,
                    help="device buffers cycled through between timed steps (inputs + outputs), >> L2")
    ap.add_argument("--cfg5-leg-hours", type=float, default=48.0,
                    help="N > 1: size of the attached config-5 sweep in audio-hours (0 disables)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# BASELINE.json configs; `utts` = utterances per batch
WORKLOADS = {
    "cfg1": dict(utts=10, cmvn="utterance", name="test/data fixture clips (10 wavs, 1.7-14.7 s, int16) + utterance CMVN"),
    "cfg2": dict(utts=256, cmvn="utterance",
                 name="librispeech_100h-shaped synthetic batch: {utts} x U(10,15) s 16 kHz int16 utterances, "
                      "80-bin Kaldi fbank + utterance CMVN (generator: joeys2t_b200.synthetic.pooled_batch, "
                      "seed 1234 + 1000 rank + batch — SURVEY 8d's envelope x pink-noise 'speech' cut from one "
                      "90 s pool with per-utterance gain; the kernels are data-independent)"),
    "cfg3": dict(utts=512, cmvn="utterance", masks=True,
                 name="mustc_st-shaped ragged batch: {utts} x lognormal(median 5 s, 1-30 s) int16, utterance CMVN + "
                      "SpecAugment (2 freq masks F=27, 2 time masks T=100)"),
    "cfg3g": dict(utts=512, cmvn="global", masks=True,
                  name="mustc_st-shaped ragged batch: {utts} x lognormal(median 5 s, 1-30 s) int16, global CMVN "
                       "(statistics known, normalised in the fbank epilogue) + SpecAugment, constant fill"),
    "cfg4": dict(utts=64, cmvn="utterance",
                 name="long-form ragged batch: {utts} x U(30,60) s, even utterances int16 / odd float32, utterance CMVN"),
    "cfg5": dict(utts=256, cmvn="sweep",
                 name="{hours:.0f}-hour synthetic corpus sweep (librispeech-shaped int16 batches cycled), global CMVN with "
                      "unknown statistics: pass 1 statistics -> all-reduce of 161 fp64 -> pass 2 normalised fbank"),
}


def bench_config(args):
    """`config` of the JSON line — identical in both arms (the driver compares them)."""
    w = WORKLOADS[args.workload]
    return {"workload": workload_name(args), "baseline_config": args.workload,
            "utterances_per_step": args.utts or w["utts"], "cmvn": w["cmvn"]}


def workload_name(args):
    w = WORKLOADS[args.workload]
    return w["name"].format(utts=args.utts or w["utts"], hours=args.sweep_hours)


def make_batch(args, seed):
    """Waveforms of one batch of the selected workload (host, numpy)."""
    from joeys2t_b200 import synthetic
    w = WORKLOADS[args.workload]
    n = args.utts or w["utts"]
    if args.workload == "cfg1":
        z = np.load(ROOT / "tests" / "golden" / "fixtures_pcm.npz")
        return [z[f"pcm{i}"] for i in range(10)]
    if args.workload in ("cfg3", "cfg3g"):
        rng = np.random.default_rng(seed)
        dur = np.clip(np.exp(rng.normal(np.log(5.0), 0.7, size=n)), 1.0, 30.0)
        pool = synthetic.pooled_batch(8, seed=seed, lo=30.0, hi=30.0)
        return [pool[i % 8][:int(d * SR)].copy() for i, d in enumerate(dur)]
    if args.workload == "cfg4":
        base = synthetic.pooled_batch(n, seed=seed, lo=30.0, hi=60.0)
        return [x if i % 2 == 0 else x.astype(np.float32) / np.float32(32768.0) for i, x in enumerate(base)]
    return synthetic.pooled_batch(n, seed=seed, lo=10.0, hi=15.0)


def algorithmic_bytes(packed, plan):
    """160 new samples read per frame (2 or 4 bytes each) + 80 float32 written."""
    per_in = np.where(packed.is_f32 != 0, 640, 320).astype(np.int64)
    return int((plan.n_frames.astype(np.int64) * (per_in + 320)).sum())


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


def cpu_kind_tag():
    """"reference" (unmodified joeynmt from /root/reference or oracle/_ref) or "port"."""
    from oracle import ref_cpu_path
    return ref_cpu_path.KIND_TAG


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
    """The whole batch: one step of the reference arm is the same 256 utterances as one GPU step
    (0.887 audio-h take ~0.15 s on 16 cores, so nothing needs to be sub-sampled)."""
    del cores
    return waves


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    NVML is polled from a thread every ~2 ms (nvidia-smi -lms cannot resolve a region of tens of
    milliseconds); falls back to one nvidia-smi query if NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.smax = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # pylint: disable=broad-except
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # pylint: disable=broad-except
                break
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self.sm:
                return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(np.min(self.sm)),
                        "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(self.sm),
                        "how": "NVML polled every ~2 ms during the timed region"}
        try:
            out = subprocess.run(
                ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                 str(self.index)], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1,
                    "how": "single nvidia-smi query after the timed region (NVML unavailable)"}
        except Exception:  # pylint: disable=broad-except
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


def build_metrics():
    """ncu counters of the dominant kernel for THE BUILD BEING MEASURED: profiles/fbank_ncu_metrics.json is
    written by tools/ncu_summary.py --json from one `ncu --set full` capture and carries a hash of the kernel
    sources; a capture of another build is not attached (returns None and the reason)."""
    from joeys2t_b200 import _lib
    p = ROOT / "profiles" / "fbank_ncu_metrics.json"
    try:
        m = json.loads(p.read_text())
    except Exception:  # pylint: disable=broad-except
        return None, "profiles/fbank_ncu_metrics.json missing"
    sha = _lib.kernel_source_sha16()
    if m.get("source_sha16") != sha:
        return None, f"profiles/fbank_ncu_metrics.json belongs to build {m.get('source_sha16')}, this is {sha}"
    return m, f"profiles/fbank_ncu_metrics.json ({m.get('report')}, build {sha})"


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
    waves = make_batch(args, 1234)
    sample = cpu_sample(waves, cores)
    hours = sum(len(w) for w in sample) / SR / 3600.0
    import multiprocessing as mp
    # one thread per worker process (the workers import torch after the fork and inherit this):
    # without it every worker spins up its own OpenMP / MKL pool and the cores are oversubscribed
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, sample, chunksize=1)
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
    sample_desc = (f"all {len(sample)} utterances of one batch ({hours:.3f} audio-h) per step, "
                   f"{cores} processes x 1 thread ({cpu_model()}); {cpu_kind()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind_tag(),
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

    from joeys2t_b200 import _lib, distributed, frontend

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    rank, local_rank, world = distributed.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host thread (and the pinned buffers it first-touches) next to the GPU: the e2e path is PCIe-bound
    orig_affinity = os.sched_getaffinity(0)
    numa_bound = frontend.bind_host_thread_to_gpu(local_rank) if not os.environ.get("JS2T_NO_NUMA_BIND") else False
    if _lib.is_stale():
        if local_rank == 0:
            _lib.build()
        if world > 1:
            dist.barrier()

    if args.workload == "cfg5":
        res = sweep_measure(args, rank, local_rank, world, dev, args.sweep_hours,
                            recompute=bool(os.environ.get("JS2T_SWEEP_RECOMPUTE")), comm=None)
        if rank == 0:
            print(json.dumps(sweep_line(args, world, res)), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- synthetic workload: `rotate` distinct batches per rank, resident in HBM ----------------
    from joeys2t_b200.data_augmentation import SpecAugment, mask_tables_for_batch
    wl = WORKLOADS[args.workload]
    R = max(1, args.rotate)
    batches, plans, pcm_dev, outs, packs = [], [], [], [], []
    for r in range(R):
        waves = make_batch(args, 1234 + 1000 * rank + r)
        packed = frontend.PackedPCM(waves)
        plan = frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32, device=local_rank)
        if wl["cmvn"] == "global":
            # statistics "known up front": taken from one statistics pass outside the timed region
            plan.set_cmvn("stats")
            dev_pcm = packed.to_device(dev)
            tmp = plan.execute(dev_pcm)
            acc = distributed.new_accumulator(dev)
            plan.accumulate_global(acc)
            mean, istd = distributed.stats_to_mean_istd(distributed.allreduce_global_stats(acc))
            del tmp
            plan.set_cmvn("global", True, True, True)
            plan.set_global_stats(mean, istd)
        else:
            plan.set_cmvn("utterance", True, True, True)
        if wl.get("masks"):
            np.random.seed(2345 + r)
            table, nf, nt = mask_tables_for_batch(
                SpecAugment(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
                plan.n_frames)
            plan.set_masks(table, nf, nt, mask_value=0.0 if wl["cmvn"] == "global" else None)
        batches.append(waves)
        packs.append(packed)
        plans.append(plan)
        pcm_dev.append(packed.to_device(dev))
        outs.append(plan.empty_output())
    torch.cuda.synchronize()
    hours_per_step = [sum(len(w) for w in b) / SR / 3600.0 for b in batches]
    frames_per_step = [p.total_frames for p in plans]
    # L2 (126 MB) must not help between timed steps: every step uses its own input AND output buffers out of
    # a ring whose total size exceeds --working-set-gb (>= 4 GB).  The R distinct synthetic batches
    # are replicated into distinct device buffers — the kernels are data-independent, for the caches only
    # the addresses matter — and each buffer keeps the plan (geometry) of the batch it holds.
    per_batch = float(np.mean([p.nbytes for p in packs]) + np.mean([o.numel() * 4 for o in outs]))
    n_buf = max(R, int(np.ceil(args.working_set_gb * 1e9 / per_batch)))
    n_buf = min(n_buf, max(R, int(0.5 * torch.cuda.mem_get_info(dev)[0] / per_batch)))
    for k in range(R, n_buf):
        pcm_dev.append(pcm_dev[k % R].clone())
        outs.append(torch.empty_like(outs[k % R]))
    torch.cuda.synchronize()
    working_set = sum(t.numel() for t in pcm_dev) + 4 * sum(o.numel() for o in outs)

    def step(i):
        plans[i % n_buf % R].execute(pcm_dev[i % n_buf], outs[i % n_buf])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(i)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    # (A) the timed region of `value`: exactly `steps` steps, nothing but the product's launches on the
    # stream.  The three kernels of a step are chained by programmatic dependent launch; an event
    # recorded between two kernels would serialise them again, so the per-kernel events of the roofline
    # line are taken in a second, identical region (B) right after, whose own step time is reported too.
    # The device spins for a moment in front of the region (a one-thread clock loop, torch.cuda._sleep) while
    # the host enqueues the steps behind it: a 20-step region is 4 ms of device time, and on a box where N
    # ranks, their NCCL / sampling threads and the driver share the host cores, one descheduled launching
    # thread otherwise shows up as a step time twice the real one (seen once at N = 8: 0.43 ms against 0.21 ms
    # in the 60-ms region of the same run).  The events still bracket exactly `steps` steps on the device.
    head_start_ms = min(6.0, 0.04 * args.steps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(head_start_ms * 1e-3 * 1.9e9))  # pylint: disable=protected-access
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    # (B) same steps again with CUDA events around the fbank kernel (recorded by the library on the
    # launching stream)
    for p in plans:
        p.enable_profiling((args.steps + R - 1) // R + 1)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(args.steps):
        step(i)
    ev3.record()
    barrier()
    clk = clocks.stop()
    ms_total_events = ev2.elapsed_time(ev3)
    hours_done = sum(hours_per_step[i % n_buf % R] for i in range(args.steps))
    frames_done = sum(frames_per_step[i % n_buf % R] for i in range(args.steps))
    kt = np.concatenate([p.kernel_times_ms((args.steps + R - 1) // R + 1) for p in plans])
    kernel_ms = float(kt.mean())
    for p in plans:
        p.enable_profiling(0)
    # (C) a longer region (>= 50 ms of device time, same loop) — the driver's --steps 20 region is ~4 ms,
    # this one shows that the short region is representative
    long_steps = max(args.steps, int(np.ceil(60.0 / max(ms_total / args.steps, 1e-3))))
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev4.record()
    for i in range(long_steps):
        step(i)
    ev5.record()
    barrier()
    ms_long = ev4.elapsed_time(ev5)
    hours_long = sum(hours_per_step[i % n_buf % R] for i in range(long_steps))

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
        e2e_hours = sum(hours_per_step[i % R] for i in range(args.steps))
        e2e = (e2e_hours, e2e_ms, int(np.mean([p.nbytes for p in packs])),
               int(np.mean([plans[j].out_rows * 80 * 4 for j in range(R)])))

    # ---- end to end once more, from what a per-batch caller really holds: separate pageable numpy arrays, one
    # per utterance, through the ONE public call (frontend.fbank_cmvn_specaug_ragged -> js2t_batch_fbank: gather
    # into a pinned staging slot, one upload, kernels), features read back into pinned host memory.  Nothing is
    # packed or planned outside the timed region.  (utterance-CMVN workloads only; the statistics of the global
    # ones come from a pass outside the region)
    e2e_call = None
    if not args.no_e2e and wl["cmvn"] == "utterance" and not wl.get("masks"):
        rows_max = max(p.out_rows for p in plans)
        host_out = [torch.empty((rows_max, 80), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        s_out = torch.cuda.Stream(device=dev)
        cm = dict(norm_means=True, norm_vars=True, before=True)

        def call_step(i):
            feats, _ = frontend.fbank_cmvn_specaug_ragged(batches[i % R], cmvn=cm, layout="ragged")
            k = i % 2
            done[k].synchronize()                       # the copy that last used this host buffer has finished
            s_out.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s_out):
                host_out[k][:feats.shape[0]].copy_(feats, non_blocking=True)
                feats.record_stream(s_out)
                done[k].record(s_out)
            return k

        for i in range(max(3, args.warmup)):
            call_step(i)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        last = 0
        for i in range(args.steps):
            last = call_step(i)
        done[last].synchronize()
        torch.cuda.current_stream(dev).wait_stream(s_out)
        c1.record()
        barrier()
        assert np.isfinite(float(host_out[last][0, 0]))
        e2e_call = (sum(hours_per_step[i % R] for i in range(args.steps)), c0.elapsed_time(c1))

    # ---- reduce over ranks: max time, summed work ---------------------------------------------------
    t = torch.tensor([ms_total, e2e[1] if e2e else 0.0, ms_long, e2e_call[1] if e2e_call else 0.0],
                     dtype=torch.float64, device=dev)
    w = torch.tensor([hours_done, float(frames_done), e2e[0] if e2e else 0.0, hours_long,
                      e2e_call[0] if e2e_call else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    ms_total_max, e2e_ms_max, ms_long_max, call_ms_max = t.tolist()
    hours_all, frames_all, e2e_hours_all, hours_long_all, call_hours_all = w.tolist()

    # ---- N > 1: the path's one collective, in front of the driver ------------------------------------
    cfg5 = None
    if world > 1 and args.cfg5_leg_hours > 0:
        cfg5 = cfg5_leg(args, rank, local_rank, world, dev)

    if rank == 0:
        value = hours_all / (ms_total_max * 1e-3)
        peak, peak_src = measured_peaks()
        frames_launch = float(np.mean(frames_per_step))
        algo_bytes = float(np.mean([algorithmic_bytes(pk, pl) for pk, pl in zip(packs, plans)]))
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        metrics, metrics_src = build_metrics() if args.workload == "cfg2" else (None, "captured on cfg2 only")
        sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args),
            "details": {
                "frames_per_step_per_gpu": frames_launch,
                "audio_hours_per_step_per_gpu": float(np.mean(hours_per_step)),
                "pcm": "resident in HBM", "cmvn": wl["cmvn"] + " (norm_means, norm_vars, before)",
                "enqueue_head_start_ms": head_start_ms,
                "cmvn_path": "fbank epilogue" if wl["cmvn"] == "global" else "fbank+stats, finalize, apply kernels",
                "l2": f"every step uses its own input and output buffers out of a ring of {n_buf} "
                      f"({working_set / 1e9:.2f} GB of inputs+outputs per GPU, >> the 126 MB L2; {R} distinct "
                      "synthetic batches replicated into distinct device buffers)",
                "working_set_bytes_per_gpu": int(working_set),
                "parallelism": f"utterance-sharded x{world}, no data-path collective in this workload",
                "long_region": {"steps": long_steps, "ms": ms_long_max,
                                "value": hours_long_all / (ms_long_max * 1e-3),
                                "note": "same loop over >= 60 ms of device time (the timed region of `value` "
                                        "is the contract's K steps)"},
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": (metrics or {}).get("dram_bytes_per_launch"),
                "traffic_source": metrics_src,
                "kernel": "fbank_tile_kernel", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": algo_bytes,
                "peak_source": peak_src,
                "kernel_share_of_step": kernel_ms / (ms_total_events / args.steps),
                "ms_per_step_with_kernel_events": ms_total_events / args.steps,
                "working_set_bytes": int(working_set),
                "how": "kernel_ms = mean of CUDA-event pairs recorded by the library around every fbank "
                       "launch on the launching stream, in a second region of the same `steps` steps right "
                       "after the timed one (events between kernels defeat the programmatic dependent "
                       "launch that chains the step's three kernels, so that region's steps are slower)",
            },
            "roofline_fp32": fp32_roofline(metrics, metrics_src, frames_launch, kernel_ms, sm_mhz, n_sm),
            "clocks": clk,
            "gpu_launches": ((2 if wl.get("masks") else 1) if wl["cmvn"] == "global" else 3) * args.steps,
        }
        if e2e:
            line["e2e"] = {"value": e2e_hours_all / (e2e_ms_max * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[2], "d2h_bytes_per_step": e2e[3],
                           "gbs_per_direction_per_gpu": max(e2e[2], e2e[3]) * args.steps / (e2e_ms_max * 1e-3) / 1e9,
                           "gbs_per_direction_all_gpus": max(e2e[2], e2e[3]) * args.steps * world / (e2e_ms_max * 1e-3) / 1e9,
                           "host_ceiling_gbs": host_ceiling(world),
                           "host_thread_bound_to_gpu_numa_node": bool(numa_bound)}
            if e2e_call:
                line["e2e"]["from_pageable_arrays"] = {
                    "value": call_hours_all / (call_ms_max * 1e-3), "unit": UNIT,
                    "how": "the same batches as separate pageable numpy arrays (one per utterance) through the one "
                           "public call frontend.fbank_cmvn_specaug_ragged -> js2t_batch_fbank (gather into a pinned "
                           "staging slot on the library's copy threads, one upload, three kernels), features copied "
                           "into pinned host memory; nothing packed or planned outside the timed region"}
        if cfg5 is not None:
            line["cfg5"] = cfg5
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, orig_affinity)  # the CPU baseline gets every core back
            cores = os.cpu_count() or 1
            sample = cpu_sample(batches[0], cores)
            v, dt, passes = cpu_port_throughput(sample, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind_tag(),
                "sample": f"all {len(sample)} utterances of one batch x {passes} passes, "
                          f"{dt:.1f} s wall on {cores} processes x 1 thread ({cpu_model()}); {cpu_kind()}"}
        print(json.dumps(line), flush=True)
    for p in plans:
        p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def fp32_roofline(metrics, metrics_src, frames_launch, kernel_ms, sm_mhz, n_sm):
    """The FP32 side of the roofline (the expected binding roof).  `achieved` = FP32 lane operations
    per second from the LIVE kernel time of this run x the per-frame lane-operation count of this build's
    SASS (ncu opcode mix); `peak` = n_sm x 128 lanes x the SM clock sampled during the run.  fma_pipe_frac /
    issue_frac are ncu's own pipe-busy and issue-slot fractions of the same build (cold-cache capture)."""
    out = {"source": metrics_src, "fma_pipe_frac": None, "issue_frac": None, "lane_ops_per_frame": None,
           "achieved": None, "peak": n_sm * 128 * sm_mhz * 1e6 / 1e12, "unit": "T lane-op/s", "frac": None}
    if metrics:
        lane = metrics.get("lane_ops_per_frame")
        out.update(fma_pipe_frac=metrics.get("fma_pipe_frac"), issue_frac=metrics.get("issue_frac"),
                   lsu_wavefront_frac=metrics.get("lsu_wavefront_frac"), lane_ops_per_frame=lane,
                   warp_inst_per_frame=metrics.get("warp_inst_per_frame"))
        if lane:
            out["achieved"] = lane * frames_launch / (kernel_ms * 1e-3) / 1e12
            out["frac"] = out["achieved"] / out["peak"]
    return out


def host_ceiling(n_gpus):
    """Raw pinned-memory copy ceiling of the 8 x B200 box for ``n_gpus`` GPUs copying at once, both directions
    concurrently (tools/h2d_probe.py, committed as profiles/h2d_probe.json; GB/s per direction, summed over
    the GPUs): what bounds the host-resident (e2e) path as GPUs are added — the host side of the box, not the
    kernels and not one GPU's PCIe link."""
    p = ROOT / "profiles" / "h2d_probe.json"
    try:
        rows = json.loads(p.read_text())["rows"]
        row = max((r for r in rows if r["gpus"] <= n_gpus), key=lambda r: r["gpus"])
        return {"gpus": row["gpus"], "h2d_gbs": row["both:h2d_gbs"], "d2h_gbs": row["both:d2h_gbs"],
                "h2d_alone_gbs": row["h2d:h2d_gbs"], "d2h_alone_gbs": row["d2h:d2h_gbs"],
                "source": "profiles/h2d_probe.json (tools/h2d_probe.py on the 8 x B200 box, H2D and D2H concurrently)"}
    except Exception:  # pylint: disable=broad-except
        return None


def cfg5_leg(args, rank, local_rank, world, dev):
    """N > 1: a short config-5 sweep (global CMVN with unknown statistics) attached to the headline line, so
    that the path's one exchange step — the all-reduce of 161 float64 — runs under the driver's own
    scaling run, with on-box checks of what it produced.  Both pass-2 strategies are run."""
    import torch
    import torch.distributed as dist

    from joeys2t_b200 import distributed
    impl = "js2t_global_stats_allreduce (C ABI -> ncclAllReduce on a communicator from js2t_nccl_comm_create)"
    comm = None
    try:
        comm = distributed.NcclCommunicator(local_rank)
    except Exception as err:  # pylint: disable=broad-except
        impl = f"torch.distributed.all_reduce (C-ABI communicator unavailable: {err})"
    flag = torch.tensor([1.0 if comm is not None else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if flag.item() < 1.0 and comm is not None:  # all ranks or none
        comm.close()
        comm = None
        impl = "torch.distributed.all_reduce (C-ABI communicator unavailable on another rank)"
    sub = argparse.Namespace(**vars(args))
    sub.workload, sub.rotate = "cfg5", min(args.rotate, 2)
    out = {"workload": WORKLOADS["cfg5"]["name"].format(hours=args.cfg5_leg_hours), "allreduce_impl": impl}
    for name, recompute in (("retain", False), ("recompute", True)):
        res = sweep_measure(sub, rank, local_rank, world, dev, args.cfg5_leg_hours, recompute=recompute, comm=comm)
        if rank == 0:
            peak, _ = measured_peaks()
            out[name] = {"value": res["value"], "unit": UNIT, "ms": res["ms"], "frac": res["achieved"] / peak,
                         "bytes_per_frame": res["bytes_per_frame"], "corpus_hours": res["hours"],
                         "frames": res["frames"]}
            out["stats_check"] = {**out.get("stats_check", {}), name: res["check"]}
            out["allreduce_us"] = res["allreduce_us"]
            out["ranks"] = res["ranks_seen"]
    if rank == 0:
        out["value"] = out["retain"]["value"]
        out["frac"] = out["retain"]["frac"]
        out["stats_check"]["ok"] = all(v.get("ok") for v in out["stats_check"].values() if isinstance(v, dict))
    if comm is not None:
        comm.close()
    return out if rank == 0 else None


def sweep_measure(args, rank, local_rank, world, dev, hours_total, recompute, comm):
    """cfg5: corpus sweep sharded over the ranks with global CMVN whose statistics are NOT known:
    pass 1 = fbank + per-utterance statistics, accumulated on the device in fp64; one all-reduce of
    161 float64 over NCCL (the path's only collective); pass 2 normalises.  Two strategies for pass 2:

    * retain (default when this rank's share of the features fits in HBM — 115 MB per audio-hour, so
      1000 h over 8 GPUs is 14 GB each): pass 1 keeps the raw log-mel of the whole share on the device
      and pass 2 normalises it in place with the apply kernel.  1 280 algorithmic bytes per frame, but
      the fbank kernel — bound by FP32 issue / latency, not HBM — runs once instead of twice;
    * recompute (or the share does not fit): pass 2 runs the fbank kernel again with the normalisation
      fused into its epilogue: 960 bytes per frame, twice the arithmetic.

    The resident PCM batches are cycled until hours_total / world have been processed by this rank.
    After the timed sweep the statistics are checked ON THE BOX: the frame count in acc[160] equals the
    number of frames all ranks processed; mean / 1/std are bit-identical on every rank; and they equal a
    rank-0 float64 recomputation from every rank's per-utterance statistics."""
    import torch
    import torch.distributed as dist

    from joeys2t_b200 import distributed, frontend
    R = max(1, args.rotate)
    plans, pcm_dev, outs, hours = [], [], [], []
    for r in range(R):
        waves = make_batch(args, 4567 + 1000 * rank + r)
        packed = frontend.PackedPCM(waves)
        plans.append(frontend.Plan(packed.n_samples, packed.byte_off, packed.is_f32, device=local_rank))
        pcm_dev.append(packed.to_device(dev))
        outs.append(plans[-1].empty_output())
        hours.append(sum(len(w) for w in waves) / SR / 3600.0)
    n_steps = max(1, int(np.ceil(hours_total / world / float(np.mean(hours)))))
    rows = [p.out_rows for p in plans]
    share_bytes = sum(rows[i % R] for i in range(n_steps)) * 80 * 4
    free_bytes = torch.cuda.mem_get_info(dev)[0]
    retain = not recompute and share_bytes < 0.8 * free_bytes
    store, views = None, []
    if retain:
        store = torch.empty(share_bytes // 4, dtype=torch.float32, device=dev)
        off = 0
        for i in range(n_steps):
            n = rows[i % R] * 80
            views.append(store[off:off + n].view(rows[i % R], 80))
            off += n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def sweep(n, keep_stats=None):
        acc = distributed.new_accumulator(dev)
        for p in plans:
            p.set_cmvn("stats")
        for i in range(n):                                   # pass 1: raw log-mel + statistics
            plans[i % R].execute(pcm_dev[i % R], views[i] if retain else outs[i % R])
            plans[i % R].accumulate_global(acc)
            if keep_stats is not None:
                keep_stats.append(plans[i % R].utt_stats())
        local = acc.clone() if keep_stats is not None else None
        distributed.allreduce_global_stats(acc, comm=comm)   # the one collective (161 x fp64)
        for p in plans:
            p.finalize_global(acc)
        if retain:
            for i in range(n):                               # pass 2: normalise the retained rows in place
                plans[i % R].normalize(views[i])
        else:
            for p in plans:
                p.set_cmvn("global", True, True, True)
            for i in range(n):                               # pass 2: fbank again, normalised in the epilogue
                plans[i % R].execute(pcm_dev[i % R], outs[i % R])
        return acc, local

    sweep(min(n_steps, max(args.warmup, 3)))
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    acc, _ = sweep(n_steps)
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    hours_done = sum(hours[i % R] for i in range(n_steps))
    frames_done = sum(plans[i % R].total_frames for i in range(n_steps))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    w = torch.tensor([hours_done, float(frames_done), 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)

    # ---- on-box checks of the exchange step (untimed): a second, instrumented sweep ------------------
    stats = []
    acc2, local = sweep(n_steps, keep_stats=stats)
    torch.cuda.synchronize()
    mine = torch.stack(stats).sum(dim=(0, 1))                       # (160,) fp64 from the per-utterance statistics
    mine = torch.cat([mine, torch.tensor([float(frames_done)], dtype=torch.float64, device=dev)])
    gm = plans[0].global_mean_istd()                  # (160,) float32 mean | 1/std on this rank
    if world > 1:
        all_acc = [torch.empty_like(acc2) for _ in range(world)]
        all_local = [torch.empty_like(mine) for _ in range(world)]
        all_gm = [torch.empty_like(gm) for _ in range(world)]
        dist.all_gather(all_acc, acc2)
        dist.all_gather(all_local, mine)
        dist.all_gather(all_gm, gm)
    else:
        all_acc, all_local, all_gm = [acc2], [mine], [gm]
    # timing of the collective itself (CUDA events around 20 back-to-back all-reduces of a scratch vector)
    scratch = distributed.new_accumulator(dev)
    distributed.allreduce_global_stats(scratch, comm=comm)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(20):
        distributed.allreduce_global_stats(scratch, comm=comm)
    a1.record()
    torch.cuda.synchronize()
    allreduce_us = a0.elapsed_time(a1) * 1e3 / 20
    res = None
    if rank == 0:
        ms_max = float(t[0])
        hours_all, frames_all, ranks_seen = w.tolist()
        total = torch.stack(all_local).sum(0).cpu().numpy()             # rank-0 fp64 recomputation
        got = acc2.cpu().numpy()
        rel = float(np.max(np.abs(got[:160] - total[:160]) / np.maximum(np.abs(total[:160]), 1e-300)))
        n = total[160]
        mu = total[:80] / n
        istd = 1.0 / np.sqrt(np.maximum(total[80:160] / n - mu**2, 1e-10))
        want = np.concatenate([mu, istd]).astype(np.float32)
        mine_gm = all_gm[0].cpu().numpy()
        check = {
            "frames_in_acc160": float(got[160]), "frames_processed_all_ranks": float(frames_all),
            "frame_count_exact": bool(got[160] == frames_all == n),
            "acc_identical_on_all_ranks": bool(all(torch.equal(a, all_acc[0]) for a in all_acc)),
            "mean_istd_bit_identical_on_all_ranks": bool(all(torch.equal(g, all_gm[0]) for g in all_gm)),
            "max_rel_err_vs_rank0_fp64_recomputation": rel,
            "mean_istd_max_abs_err_vs_recomputation": float(np.max(np.abs(mine_gm - want))),
        }
        check["ok"] = bool(check["frame_count_exact"] and check["acc_identical_on_all_ranks"] and
                           check["mean_istd_bit_identical_on_all_ranks"] and rel <= 1e-12 and
                           check["mean_istd_max_abs_err_vs_recomputation"] <= 1e-6)
        # recompute: 960 B/frame (PCM read twice, features written once; the raw rows pass 1 also writes
        # are not algorithmic traffic); retain: 1 280 B/frame (PCM in, raw out, raw in, normalised out)
        bytes_per_frame = 1280.0 if retain else 960.0
        res = {"value": hours_all / (ms_max * 1e-3), "ms": ms_max, "hours": hours_all, "frames": frames_all,
               "retain": retain, "share_bytes": share_bytes, "bytes_per_frame": bytes_per_frame,
               "achieved": frames_all / world * bytes_per_frame / (ms_max * 1e-3) / 1e9, "n_steps": n_steps,
               "R": R, "n_plans": len(plans), "clk": clk, "check": check, "allreduce_us": allreduce_us,
               "ranks_seen": int(ranks_seen), "acc160": float(acc[160]),
               "allreduce_impl": "C ABI ncclAllReduce" if comm is not None else "torch.distributed"}
    for p in plans:
        p.close()
    del store
    return res


def sweep_line(args, world, res):
    peak, peak_src = measured_peaks()
    retain, n_steps = res["retain"], res["n_steps"]
    return {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
        "steps": 2 * n_steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms"] / (2 * n_steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "baseline_config": "cfg5",
                   "corpus_hours": res["hours"], "frames": res["frames"], "passes": 2,
                   "pass2": ("in-place normalisation of the raw log-mel retained in HBM "
                             f"({res['share_bytes'] / 1e9:.1f} GB per GPU)") if retain
                            else "fbank recomputed with the normalisation fused into its epilogue",
                   "collective": "one all-reduce (SUM) of 161 float64 between the passes",
                   "allreduce_impl": res["allreduce_impl"], "allreduce_us": res["allreduce_us"],
                   "global_frames_in_statistics": res["acc160"], "stats_check": res["check"],
                   "l2": f"cycling {res['R']} resident batches per GPU (~{res['R'] * 204} MB of PCM + features)",
                   "parallelism": f"utterance-sharded x{world}"},
        "roofline": {"bound": "hbm", "achieved": res["achieved"], "peak": peak, "unit": "GB/s",
                     "frac": res["achieved"] / peak, "traffic": None, "peak_source": peak_src,
                     "note": f"per GPU; {res['bytes_per_frame']:.0f} algorithmic bytes per frame for the two-pass "
                             "global CMVN"},
        "clocks": res["clk"], "gpu_launches": 4 * n_steps + res["n_plans"],
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

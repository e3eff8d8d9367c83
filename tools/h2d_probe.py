# coding: utf-8
"""Raw host<->device copy ceiling of the box, for the end-to-end (host-resident) number of bench.py:

    python tools/h2d_probe.py [--mb 256] [--reps 8] [--json profiles/h2d_probe.json]

For k = 1, 2, 4, ... GPUs of the box at once: pinned host buffers, one H2D and one D2H stream per GPU,
`reps` back-to-back cudaMemcpyAsync of `mb` MB each, H2D alone, D2H alone, and both directions
concurrently (PCIe is full duplex).  Device time per GPU from CUDA events, wall time around the whole
group; the aggregate is bytes / max over GPUs.  No kernels, no library code: this is what a perfect
pipeline could reach, and the e2e path of bench.py (102 MB each way per step) is read against it.
"""
import argparse
import json
import time

import torch


def probe(devs, mb, reps, mode):
    n = mb << 20
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            bufs.append(dict(
                h_in=torch.empty(n, dtype=torch.uint8, pin_memory=True),
                h_out=torch.empty(n, dtype=torch.uint8, pin_memory=True),
                d_in=torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"),
                d_out=torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"),
                s_in=torch.cuda.Stream(d), s_out=torch.cuda.Stream(d),
                ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]))
    for d in devs:
        torch.cuda.synchronize(d)

    def run(r):
        for d, b in zip(devs, bufs):
            with torch.cuda.device(d):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(b["s_in"]):
                        b["ev"][0].record()
                        for _ in range(r):
                            b["d_in"].copy_(b["h_in"], non_blocking=True)
                        b["ev"][1].record()
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(b["s_out"]):
                        b["ev"][2].record()
                        for _ in range(r):
                            b["h_out"].copy_(b["d_out"], non_blocking=True)
                        b["ev"][3].record()
        for d in devs:
            torch.cuda.synchronize(d)

    run(2)
    t0 = time.perf_counter()
    run(reps)
    wall = time.perf_counter() - t0
    per_dir = {}
    if mode in ("h2d", "both"):
        per_dir["h2d"] = max(b["ev"][0].elapsed_time(b["ev"][1]) for b in bufs) * 1e-3
    if mode in ("d2h", "both"):
        per_dir["d2h"] = max(b["ev"][2].elapsed_time(b["ev"][3]) for b in bufs) * 1e-3
    total = len(devs) * reps * n
    return {k: total / v / 1e9 for k, v in per_dir.items()}, total * len(per_dir) / wall / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    rows = []
    k = 1
    while k <= n_dev:
        devs = list(range(k))
        row = {"gpus": k}
        for mode in ("h2d", "d2h", "both"):
            per_dir, wall_gbs = probe(devs, args.mb, args.reps, mode)
            for d, v in per_dir.items():
                row[f"{mode}:{d}_gbs"] = round(v, 2)
            row[f"{mode}:wall_gbs"] = round(wall_gbs, 2)
        rows.append(row)
        print(row, flush=True)
        k *= 2
    out = {"how": f"tools/h2d_probe.py: {args.reps} x {args.mb} MB pinned cudaMemcpyAsync per GPU and direction, "
                  "aggregate GB/s over the GPUs used at once (CUDA events, max over GPUs); 'both' = H2D and D2H "
                  "concurrently, per direction", "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if args.json:
        with open(args.json, "w", encoding="utf8") as f:
            json.dump(out, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()

# coding: utf-8
"""Warp instructions per frame and stall-sample share by source-line range of fbank_kernels.cu,
from an ncu --set full --import-source on report.  Usage:
   python tools/ncu_phase_hist.py REPORT FRAMES  name:lo-hi [name:lo-hi ...]"""
import csv
import io
import subprocess
import sys

rep, F = sys.argv[1], float(sys.argv[2])
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                      capture_output=True, text=True).stdout
hdr, rows, seen = None, [], 0


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in csv.reader(io.StringIO(text)):
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        seen += 1
        continue
    if hdr is None or seen != 1 or not r[0].strip().isdigit():
        continue
    rows.append((int(r[0]), num(r[hdr.index("Instructions Executed")]), num(r[hdr.index("# Samples")])))
tot = sum(r[1] for r in rows)
ts = sum(r[2] for r in rows)
print(f"total {tot / F:.1f} warp instructions per frame, {ts} stall samples")
for spec in sys.argv[3:]:
    name, rg = spec.split(":")
    a, b = map(int, rg.split("-"))
    n = sum(r[1] for r in rows if a <= r[0] <= b)
    s = sum(r[2] for r in rows if a <= r[0] <= b)
    print(f"{name:36s} lines {a:4d}-{b:4d}: {n / F:7.1f} instr/frame ({100 * n / tot:5.1f} %)   stall samples {100 * s / max(ts, 1):5.1f} %")

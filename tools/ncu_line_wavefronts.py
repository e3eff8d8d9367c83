# coding: utf-8
"""Shared-memory wavefronts (and instructions, stall samples) per CUDA source line of the first captured
kernel of an `ncu --set full --import-source on` report:
    python tools/ncu_line_wavefronts.py gpurun_out/x.ncu-rep --frames N [--top 40]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 1.0
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                      capture_output=True, text=True, check=True).stdout
hdr, rows, seen = None, [], 0
for r in csv.reader(io.StringIO(text)):
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        seen += 1
        continue
    if hdr is None or seen != 1 or not r[0].strip().isdigit():
        continue
    rows.append(r)
ci = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[ci[name]] or 0)
    except (ValueError, KeyError):
        return 0.0


tot_w = sum(num(r, "L1 Wavefronts Shared") for r in rows)
tot_i = sum(num(r, "Instructions Executed") for r in rows)
print(f"# {rep}: {tot_w / frames:.1f} shared-memory wavefronts and {tot_i / frames:.1f} warp instructions per frame")
print(f"{'line':>5s} {'wavefronts/frame':>17s} {'ideal':>8s} {'inst/frame':>11s}  source")
for r in sorted(rows, key=lambda r: -num(r, "L1 Wavefronts Shared"))[:top]:
    print(f"{r[0]:>5s} {num(r, 'L1 Wavefronts Shared') / frames:17.2f} {num(r, 'L1 Wavefronts Shared Ideal') / frames:8.2f} "
          f"{num(r, 'Instructions Executed') / frames:11.2f}  {r[1].strip()[:100]}")

# coding: utf-8
"""Dynamic instruction mix of a kernel from an ``ncu --set full --import-source on`` report:

    python tools/ncu_source_hist.py gpurun_out/prof.ncu-rep [--frames N] [--launch K] [--lines]

Aggregates the SASS page's "Instructions Executed" and stall samples by opcode (and, with
--lines, the CUDA-C source page by line).  With --frames prints warp instructions per frame.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def page(rep, which):
    args = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if which == "sass":
        args += ["--print-source", "sass"]
    else:
        args += ["--print-source", "cuda"]
    return subprocess.run(args, capture_output=True, text=True, check=True).stdout


def split_kernels(text):
    """The source page prints one table per launch, each introduced by a "Kernel Name" row."""
    out, cur = [], None
    for row in csv.reader(io.StringIO(text)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    return out


def main():
    rep = sys.argv[1]
    frames = float(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else None
    launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    k = split_kernels(page(rep, "sass"))[launch]
    col = {h: i for i, h in enumerate(k["hdr"])}
    by_op = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in k["rows"]:
        src = r[col["Source"]].strip()
        src = re.sub(r"^@!?U?P\w+\s+", "", src)
        op = src.split()[0].rstrip(";") if src else "?"
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
        n = int(r[col["Instructions Executed"]] or 0)
        s = int(r[col["# Samples"]] or 0)
        by_op[op][0] += n
        by_op[op][1] += s
        by_op[op][2] += 1
        tot_i += n
        tot_s += s
    print(f"# {k['name']}  launch {launch}: {tot_i} warp instructions, {tot_s} stall samples")
    scale = 1.0 / frames if frames else 1.0
    unit = "per frame" if frames else "total"
    print(f"{'opcode':12s} {'static':>7s} {'executed ' + unit:>22s} {'% inst':>8s} {'% samples':>10s}")
    for op, (n, s, c) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{op:12s} {c:7d} {n * scale:22.2f} {100.0 * n / max(tot_i, 1):8.2f} {100.0 * s / max(tot_s, 1):10.2f}")
    if "--lines" in sys.argv:
        # cuda,sass view: rows with a line number carry the metrics aggregated over that line's SASS
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
            if hdr is None or seen != launch + 1 or not r[0].strip().isdigit():
                continue
            ci = hdr.index("Instructions Executed")
            cs = hdr.index("# Samples")
            try:
                rows.append((int(r[ci] or 0), int(r[cs] or 0), r[1].strip()[:105], r[0]))
            except ValueError:
                continue
        tl = sum(t[1] for t in rows)
        print("\n# hottest source lines (line, warp instructions " + unit + ", % of stall samples)")
        for n, sm, src, ln in sorted(rows, key=lambda t: -t[1])[:70]:
            print(f"{ln:>6s} {n * scale:12.2f} {100.0 * sm / max(tl, 1):7.2f}%  {src}")


if __name__ == "__main__":
    main()

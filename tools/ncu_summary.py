# coding: utf-8
"""Curated text summary of an ``ncu --set full`` report (read here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--frames N] [--json OUT.json] > profiles/<name>.txt

``--json`` additionally writes the per-build counters bench.py attaches to its line (``roofline.traffic``,
``roofline_fp32``): DRAM bytes, FMA-pipe / issue / LSU fractions and per-frame instruction counts of the
FIRST captured launch, keyed by the hash of the kernel sources (joeys2t_b200._lib.kernel_source_sha16).

Prints, per captured launch: duration, DRAM traffic, issue / pipe utilisation, shared-memory
wavefronts and bank conflicts, occupancy, warp-stall breakdown, and (with --frames) per-frame
instruction and wavefront counts.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__cycles_active.avg",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum",
    "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum",
    "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    frames = None
    if "--frames" in sys.argv:
        frames = float(sys.argv[sys.argv.index("--frames") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of {rep}  ({len(data)} launches)")
    for k, r in enumerate(data):
        print(f"\n## launch {k}: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        vals = {}
        for name in WANT:
            if name in col:
                vals[name] = r[col[name]]
                print(f"{name:75s} {r[col[name]]:>18s} {units[col[name]]}")
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
               h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        if stalls:
            print("-- warp stall reasons (warps stalled per issue-active cycle), top 10")
            for v, h in sorted(stalls, reverse=True)[:10]:
                print(f"{h:75s} {v:18.3f}")
        def num(n):
            try:
                return float(vals[n].replace(",", ""))
            except (KeyError, ValueError):
                return None

        if k == 0 and "--json" in sys.argv:
            from joeys2t_b200 import _lib

            def scaled(n):
                v = num(n)
                if v is None:
                    return None
                u = units[col[n]]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

            def sass(op):  # thread-level (lane) FP32 instruction counts, when the section was collected
                n = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
                return float(r[col[n]].replace(",", "")) if n in col and r[col[n]] not in ("", "n/a") else None

            rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
            fp = [sass(o) for o in ("fadd", "fmul", "ffma")]
            out = {
                "source_sha16": _lib.kernel_source_sha16(),
                "kernel": r[col["Kernel Name"]], "report": Path(rep).name,
                "how": "ncu --set full --clock-control none, first captured launch (cold caches, serialised)",
                "frames_per_launch": frames,
                "gpu_time_us": num("gpu__time_duration.sum"),
                "dram_bytes_read": rd, "dram_bytes_write": wr,
                "dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
                "fma_pipe_frac": None if num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") is None
                else num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
                "issue_frac": None if num("smsp__issue_active.avg.pct_of_peak_sustained_active") is None
                else num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                "lsu_wavefront_frac": None if num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") is None
                else num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0,
                "registers_per_thread": num("launch__registers_per_thread"),
                "warps_active_per_scheduler": num("smsp__warps_active.avg.per_cycle_active"),
            }
            if frames:
                wi = num("smsp__inst_executed.sum")
                out["warp_inst_per_frame"] = None if wi is None else wi / frames
                # FP32 lane operations from the SASS opcode mix of the same launch (the sass_thread_inst
                # counters do not see the packed FADD2 / FMUL2 / FFMA2): a packed instruction is two
                # FMA-pipe lane operations per active thread
                try:
                    import re
                    from ncu_source_hist import page, split_kernels
                    kk = split_kernels(page(rep, "sass"))[0]
                    cc = {h: i for i, h in enumerate(kk["hdr"])}
                    fp1 = fp2 = 0
                    for row in kk["rows"]:
                        src = re.sub(r"^@!?U?P\w+\s+", "", row[cc["Source"]].strip())
                        op = src.split()[0].rstrip(";").split(".")[0] if src else "?"
                        n_exec = int(row[cc["Instructions Executed"]] or 0)
                        if op in ("FADD", "FMUL", "FFMA"):
                            fp1 += n_exec
                        elif op in ("FADD2", "FMUL2", "FFMA2"):
                            fp2 += n_exec
                    act = num("smsp__thread_inst_executed_per_inst_executed.ratio") or 32.0
                    out["fp32_warp_inst_per_frame"] = {"scalar": fp1 / frames, "packed_x2": fp2 / frames}
                    out["lane_ops_per_frame"] = (fp1 + 2 * fp2) * act / frames
                    out["lane_ops_how"] = ("(FADD+FMUL+FFMA + 2 x (FADD2+FMUL2+FFMA2)) warp instructions of the SASS "
                                           f"page x {act:.2f} active threads per instruction")
                except Exception as err:  # pylint: disable=broad-except
                    out["lane_ops_per_frame"] = None
                    out["lane_ops_how"] = f"unavailable: {err}"
            Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1) + "\n")

        if frames:
            print(f"-- per frame ({frames:.0f} frames per launch)")
            for n in ("smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
                      "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
                      "sm__inst_executed_pipe_xu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                      "dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = num(n)
                if v is not None:
                    scale = 1e6 if n.startswith("dram") and "Mbyte" in units[col[n]] else 1.0
                    print(f"{n:75s} {v * scale / frames:18.2f}")


if __name__ == "__main__":
    main()

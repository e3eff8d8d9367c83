# coding: utf-8
"""Curated text summary of an ``ncu --set full`` report (read here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--frames N] > profiles/<name>.txt

Prints, per captured launch: duration, DRAM traffic, issue / pipe utilisation, shared-memory
wavefronts and bank conflicts, occupancy, warp-stall breakdown, and (with --frames) per-frame
instruction and wavefront counts.
"""
import csv
import io
import subprocess
import sys

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
        if frames:
            def num(n):
                try:
                    return float(vals[n].replace(",", ""))
                except (KeyError, ValueError):
                    return None
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

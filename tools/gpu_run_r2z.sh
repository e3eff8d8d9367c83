#!/bin/bash
# round 2, final run of a build: full GPU suite, ncu capture of THIS build (-> per-build counters the bench line
# attaches), reference arm, headline bench, launch list of the bench command, the other BASELINE configs, next rows.
#   tools/gpu_run_r2z.sh [TAG]      (outputs gpurun_out/TAG_*; TAG defaults to r2z)
set -x
cd "${GRAFT_REPO_ROOT:-.}"
T=${1:-r2z}
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
# ncu --set full of the fbank kernel of this build; summary + per-build counters (bench.py reads profiles/)
timeout 300 bash tools/ncu_one.sh ${T}_fbank
timeout 120 python tools/ncu_summary.py gpurun_out/${T}_fbank.ncu-rep --frames 318883 --json gpurun_out/${T}_fbank_ncu_metrics.json \
    > gpurun_out/${T}_fbank_tile_kernel_ncu_full.txt
cp gpurun_out/${T}_fbank_ncu_metrics.json profiles/fbank_ncu_metrics.json
timeout 120 python tools/ncu_source_hist.py gpurun_out/${T}_fbank.ncu-rep 318883 > gpurun_out/${T}_fbank_tile_kernel_opcode_mix.txt 2>&1
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_cfg2_steps20.json 2> gpurun_out/${T}_bench_cfg2_steps20.err
timeout 400 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
tail -3 gpurun_out/${T}_bench_cfg2.err
cut -c1-400 gpurun_out/${T}_bench_cfg2.json
# launch list of the bench command (serialised, cold: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_launches.log 2>&1
for w in cfg1 cfg3 cfg3g cfg4; do
  timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
timeout 400 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg5_n1.json 2> gpurun_out/${T}_bench_cfg5_n1.err
timeout 300 python tools/bench_next_rows.py > gpurun_out/${T}_next_rows.txt 2>&1
python - <<P
import json
for w in ("cfg2_steps20","cfg2","cfg1","cfg3","cfg3g","cfg4","cfg5_n1","reference_arm"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, round(d["value"],1), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("traffic"))
    except Exception as e:
        print(w, "FAILED", e)
P

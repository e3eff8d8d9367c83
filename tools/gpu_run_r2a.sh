#!/bin/bash
# round 2, run A: full GPU test suite, headline bench, reference arm, ncu baseline of this build, copy probe
set -x
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
tail -3 gpurun_out/r2a_bench_cfg2.err
cut -c1-600 gpurun_out/r2a_bench_cfg2.json
timeout 300 bash tools/ncu_one.sh r2a_fbank
timeout 120 python tools/h2d_probe.py --json gpurun_out/r2a_h2d_probe_n1.json

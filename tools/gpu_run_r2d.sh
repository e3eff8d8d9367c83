#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "INFO\|TRG\]" | tail -15 > gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
timeout 300 python tools/bench_next_rows.py > gpurun_out/r2d_next_rows.txt 2>&1
grep -A60 "f-2" gpurun_out/r2d_next_rows.txt | head -70
timeout 200 python tools/per_item_latency.py 2>&1 | head -40 > gpurun_out/r2d_per_item.txt
head -3 gpurun_out/r2d_per_item.txt

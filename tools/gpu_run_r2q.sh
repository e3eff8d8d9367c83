#!/bin/bash
# round 2, session 3: full GPU suite with the one-call per-batch entry point, rows either side of the path
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_next_rows.py > gpurun_out/r2q_next_rows.txt 2>&1
tail -40 gpurun_out/r2q_next_rows.txt

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 20 > gpurun_out/corun_clk.csv &
SMI=$!
JS2T_LIB=build/libjs2t_side40.so python tools/corun_time.py side40
kill $SMI
sort gpurun_out/corun_clk.csv | uniq -c | sort -rn | head -20

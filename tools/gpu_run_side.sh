#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
T=$1; shift
if [ "$T" = "test" ]; then JS2T_LIB=build/libjs2t_$1.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for V in "$@"; do
JS2T_LIB=build/libjs2t_$V.so python tools/corun_time.py $V
JS2T_LIB=build/libjs2t_$V.so python tools/two_stream_time.py $V
done

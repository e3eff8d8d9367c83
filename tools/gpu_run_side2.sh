#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for V in "$@"; do
JS2T_LIB=build/libjs2t_$V.so python tools/corun_time.py $V
JS2T_LIB=build/libjs2t_$V.so python tools/two_stream_time.py $V
done

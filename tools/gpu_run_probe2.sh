#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
python tools/two_stream_time.py product
for v in "$@"; do JS2T_LIB=build/libjs2t_$v.so python tools/two_stream_time.py $v; done

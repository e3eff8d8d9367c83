#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/quick_time.py product
python tools/two_stream_time.py product

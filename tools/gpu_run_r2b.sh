#!/bin/bash
# round 2, run B (2 GPUs): remaining GPU tests, N=2 headline bench with the cfg5 leg (C-ABI NCCL all-reduce), copy probe
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_cfg2_n2.json 2> gpurun_out/r2b_bench_cfg2_n2.err
tail -5 gpurun_out/r2b_bench_cfg2_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench_cfg2_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'])
print(json.dumps(d.get('cfg5'), indent=1))
P
timeout 200 python tools/h2d_probe.py --json gpurun_out/r2b_h2d_probe_n2.json

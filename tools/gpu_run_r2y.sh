#!/bin/bash
# round 2, final build on the 8-GPU box: headline bench at N = 8 and N = 2 with the cfg5 leg (C-ABI NCCL all-reduce + checks)
set -x
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L | wc -l; nproc
for n in 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2y_bench_cfg2_n$n.json 2> gpurun_out/r2y_bench_cfg2_n$n.err
  tail -3 gpurun_out/r2y_bench_cfg2_n$n.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2y_bench_cfg2_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['gbs_per_direction_per_gpu'])
c=d.get('cfg5') or {}
print({k:c.get(k) for k in ('value','frac','allreduce_us','ranks','allreduce_impl')}, (c.get('stats_check') or {}).get('ok'))
P
done

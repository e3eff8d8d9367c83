#!/bin/bash
# round 2, 8-GPU run: raw copy ceiling at 1/2/4/8 GPUs, headline bench at N = 8 and N = 4 with the cfg5 leg
set -x
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m | head -12
timeout 300 python tools/h2d_probe.py --json gpurun_out/r2n_h2d_probe.json
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2n_bench_cfg2_n$n.json 2> gpurun_out/r2n_bench_cfg2_n$n.err
  tail -3 gpurun_out/r2n_bench_cfg2_n$n.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2n_bench_cfg2_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['e2e']['value'], d['e2e']['gbs_per_direction_per_gpu'])
c=d.get('cfg5') or {}
print({k:c.get(k) for k in ('value','frac','allreduce_us','ranks','allreduce_impl')}, (c.get('stats_check') or {}).get('ok'))
P
done

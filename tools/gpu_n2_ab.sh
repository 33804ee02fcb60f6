#!/bin/bash
# 2 GPUs: the paired bench under variants of the routing (environment switches), one JSON line each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "FQG_ROUTE_AFTER=0" "FQG_ROUTE_AFTER=1"; do
  echo "== $v"
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 2> gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['ms_per_step'], 2), 'ms', round(d['value']), 'GB/s', {k: round(v, 2) for k, v in d['roofline']['all_kernels_ms_per_step'].items() if v}, d['parity']['all_ok'])
"
done

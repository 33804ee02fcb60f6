#!/bin/bash
# NP (default 2) ranks, the routing rounds narrated on stderr with host time stamps (FQG_DEBUG_ROUTE), short flag patience
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NP=${NP:-2}
FQG_FLAG_PATIENCE_MS=${PATIENCE:-300} FQG_DEBUG_ROUTE=1 FQG_GLOO_TIMEOUT_S=40 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NP --steps 1 --warmup 1 "$@" > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep "route\]" gpurun_out/bench_n2.err | tail -${LINES_OUT:-90}; grep -v "route\]" gpurun_out/bench_n2.err | grep -E "Error|error|rank[01]\]:" | head -20; cut -c1-1200 gpurun_out/bench_n2.json

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "one_gpu" 2>&1 | tail -8
FQG_DEBUG_ROUTE=1 FQG_GLOO_TIMEOUT_S=40 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 "$@" > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep "route\]" gpurun_out/bench_n2.err | tail -40; grep -v "route\]" gpurun_out/bench_n2.err | grep -E "Error|error|rank[01]\]:" | head -20; cut -c1-1500 gpurun_out/bench_n2.json

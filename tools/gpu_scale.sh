#!/bin/bash
# N GPUs ($1): the paired bench; extra flags after the count
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
FQG_GLOO_TIMEOUT_S=90 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep -E "Error|rank[0-9]\]:" gpurun_out/bench_n$N.err | head -8; cut -c1-3500 gpurun_out/bench_n$N.json

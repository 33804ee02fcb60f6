#!/bin/bash
# two-GPU check: sharded parity over NCCL, the sharded bench (pipelined and one-exchange routing), a short one-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
[ -z "$SKIPTEST" ] && timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_n$N$1.err | grep '^{' > gpurun_out/bench_n$N$1.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_n$N$1.err | tail -4 | cut -c1-300; python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N$1.json")); print("N=$N$1", round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms", {k: round(v,2) for k,v in d["roofline"]["all_kernels_ms_per_step"].items() if v})
PY
}
FQG_DEBUG_ROUTE=1 run ""
[ -n "$STORES" ] && FQG_P2P_STORES=1 run _stores
[ -n "$A2A" ] && FQG_P2P=0 run _a2a
[ -n "$QUICK" ] && exit 0
FQG_NO_PIPELINE=1 run _oneexchange
timeout 200 python bench.py --no-e2e --cpu-sample-reads 200000 2> gpurun_out/bench_n1q.err | grep '^{' > gpurun_out/bench_n1q.json; tail -3 gpurun_out/bench_n1q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n1q.json")); print("N=1", round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms", {k: round(v,2) for k,v in d["roofline"]["all_kernels_ms_per_step"].items() if v}, d["roofline"].get("alone"))
PY

#!/bin/bash
# sharded bench at N GPUs under a few tuning settings (env assignments, one per argument after N)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
i=0
for setting in "$@"; do
  i=$((i+1))
  echo "== $setting"
  env $setting FQG_DEBUG_ROUTE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2981$i bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/sweep_$i.err | grep '^{' > gpurun_out/sweep_$i.json
  grep "host ms\|kernel ms" gpurun_out/sweep_$i.err | cut -c1-260
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$i.json")); print(round(d["value"],1), "GB/s", round(d["ms_per_step"],2), "ms", {k: round(v,2) for k,v in d["roofline"]["all_kernels_ms_per_step"].items() if v})
PY
done

#!/bin/bash
# N-GPU check (default 2): sharded parity over NCCL + CUDA IPC, then the sharded bench under the routing variants of dist.py
#   bash tools/gpu_n2.sh [N]        (QUICK=1: the default routing only)
cd "$(dirname "$0")/.."
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8
if [ -n "$QUICK" ]; then
  bash tools/gpu_n2_sweep.sh "$N" "FQG_P2P=1"
else
  bash tools/gpu_n2_sweep.sh "$N" "FQG_P2P=1" "FQG_P2P_STORES=1" "FQG_P2P=0" "FQG_NO_PIPELINE=1"
fi

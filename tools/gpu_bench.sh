#!/bin/bash
# GPU: the streaming / parity tests that touch the clean-data pass, then bench lines (arguments: extra bench.py flags)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_streaming.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json

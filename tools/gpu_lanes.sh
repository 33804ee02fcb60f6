#!/bin/bash
# GPU check after a change of the clean-data pass: its targeted tests, the streaming tests, then the whole suite, then the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_streaming.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json

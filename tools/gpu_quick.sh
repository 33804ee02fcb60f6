#!/bin/bash
# quick GPU check of the clean-data pass: targeted tests first, compute-sanitizer on one of them, then the rest
set -x
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lanes" 2>&1 | tail -30

#!/bin/bash
# ncu capture of the dominant kernel + launch list (per B200_PROFILING.md)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=${1:-fq_lanes_kernel}
python tools/prof_lanes.py 5900000 3 | tail -2
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K python tools/prof_lanes.py 5900000 2 > gpurun_out/ncu_$K.log 2>&1
tail -3 gpurun_out/ncu_$K.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/prof_lanes.py 5900000 2 index > /dev/null 2>&1
ls -la gpurun_out

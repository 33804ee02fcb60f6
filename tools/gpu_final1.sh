#!/bin/bash
# one GPU, end of a round: the whole GPU suite, the default bench, the long-read bench, the ncu capture of the clean-data pass (index
# mode: names hashed and copied to the arena, as in the paired job) and the launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json
timeout 600 python bench.py --workload longread > gpurun_out/bench_longread_n1.json 2> gpurun_out/bench_longread_n1.err; tail -3 gpurun_out/bench_longread_n1.err; cut -c1-400 gpurun_out/bench_longread_n1.json
timeout 600 python bench.py --workload illumina_se --no-extras > gpurun_out/bench_se_n1.json 2> gpurun_out/bench_se_n1.err; cut -c1-300 gpurun_out/bench_se_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_lanes_kernel -s 3 -c 1 -f -o gpurun_out/prof_lanes_index python tools/prof_lanes.py 5900000 2 index > gpurun_out/ncu_lanes_index.log 2>&1; tail -2 gpurun_out/ncu_lanes_index.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-e2e > gpurun_out/bench_under_ncu.json 2> /dev/null
wc -l gpurun_out/launches_bench.csv

#!/bin/bash
# ncu --set full captures of the hash kernels (one launch each) + this round's launch list of a paired job
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in engine shard pack; do python tools/prof_hash.py $w 2>&1 | tail -3; done
ncu --set full --clock-control none --import-source on -k regex:'fq_index_insert_kernel|fq_mate_claim_kernel' -s 2 -c 2 -f -o gpurun_out/prof_hash_engine python tools/prof_hash.py engine > gpurun_out/ncu_hash_engine.log 2>&1; tail -2 gpurun_out/ncu_hash_engine.log
ncu --set full --clock-control none --import-source on -k regex:'fq_shard_insert_slots_kernel|fq_shard_claim_slots_kernel' -s 2 -c 2 -f -o gpurun_out/prof_hash_shard python tools/prof_hash.py shard > gpurun_out/ncu_hash_shard.log 2>&1; tail -2 gpurun_out/ncu_hash_shard.log
ncu --set full --clock-control none --import-source on -k regex:'fq_names_pack_slots_kernel' -s 1 -c 1 -f -o gpurun_out/prof_hash_pack python tools/prof_hash.py pack > gpurun_out/ncu_hash_pack.log 2>&1; tail -2 gpurun_out/ncu_hash_pack.log
ncu --set full --clock-control none --import-source on -k regex:'fq_lanes_kernel' -s 3 -c 1 -f -o gpurun_out/prof_lanes_index python tools/prof_lanes.py 5900000 2 index > gpurun_out/ncu_lanes_index.log 2>&1; tail -2 gpurun_out/ncu_lanes_index.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_pair.csv python tools/prof_hash.py engine > /dev/null 2>&1
ls -la gpurun_out | tail -12

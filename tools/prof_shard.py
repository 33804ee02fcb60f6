"""World-1 run of the sharded orchestration (prescan, names to owners, owner-side insert) on 6 M synthetic reads: for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_utils_b200 as fq
from fastq_utils_b200 import dist as fqdist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_900_000
rb = fq.illumina_record_bytes()
t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
fq.synth_illumina(t, 0, n, seed=42, mate=1, perm_window=0, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
run = fqdist.ShardedFastqInfo(fq.MODE_INDEX, device=0, n_hint=n)
for _ in range(2):
    res = run.run_device(t.data_ptr(), n * rb, name="a.fq")
print(res["transcript"][0], res["transcript"][2][-60:])

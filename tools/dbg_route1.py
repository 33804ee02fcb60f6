"""one GPU, world of one: a routed one-file and a routed two-file job with the pass's result words printed (FQG_DEBUG=1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FQG_MAX_CHUNK_BYTES"] = str(24 << 20)
os.environ["FQG_DEBUG"] = "1"
os.environ["FQG_DEBUG_ROUTE"] = "1"
import torch
import fastq_utils_b200 as fq
from fastq_utils_b200 import dist as fqdist
rb = fq.illumina_record_bytes()
n = 300 * 1024
st = torch.cuda.current_stream().cuda_stream
f1 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
f2 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
fq.synth_illumina(f1, 0, n, seed=43, mate=1, stream=st)
fq.synth_illumina(f2, 0, n, seed=43, mate=2, perm_window=1024, stream=st)
torch.cuda.synchronize()
run = fqdist.ShardedFastqInfo(fq.MODE_INDEX, device=0, n_hint=n)
res = run.run_device(f1.data_ptr(), n * rb, name="a.fq")
print("one file: reruns", run.exact_reruns, "rounds", run.rounds_done, "plan", run._plan, file=sys.stderr)
run = fqdist.ShardedFastqInfo(fq.MODE_INDEX_PAIR, device=0, n_hint=n)
res = run.run_device(f1.data_ptr(), n * rb, name="a.fq", ptr2=f2.data_ptr(), nbytes2=n * rb, name2="b.fq")
print("two files: reruns", run.exact_reruns, "rounds", run.rounds_done, file=sys.stderr)

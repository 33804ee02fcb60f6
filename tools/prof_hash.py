"""Driver for ncu captures of the hash kernels: one paired job (6 M pairs, ~4.3 GB) through
  engine   the one-GPU engine: fq_index_insert_kernel (file 1), fq_mate_claim_kernel (file 2)
  shard    the sharded path with a world of one, names written by the clean-data pass: fq_shard_insert_slots_kernel, fq_shard_claim_slots_kernel
  pack     the same with FQG_ROUTE_IN_PASS=0: fq_names_pack_slots_kernel fills the regions"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "engine"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5_898_240
if which == "pack":
    os.environ["FQG_ROUTE_IN_PASS"] = "0"
import torch
import fastq_utils_b200 as fq
from fastq_utils_b200 import dist as fqdist
rb = fq.illumina_record_bytes()
st = torch.cuda.current_stream().cuda_stream
f1 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
f2 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
fq.synth_illumina(f1, 0, n, seed=43, mate=1, stream=st)
fq.synth_illumina(f2, 0, n, seed=43, mate=2, perm_window=1024, stream=st)
torch.cuda.synchronize()
if which == "engine":
    h = fq.FastqInfo(fq.MODE_INDEX_PAIR, index_capacity_hint=n)
    for _ in range(2):
        h.reset()
        h.feed_device(0, f1.data_ptr(), n * rb, last=True)
        h.feed_device(1, f2.data_ptr(), n * rb, last=True)
        rep = h.finish()
        assert rep.error.code == 0 and rep.n_index_entries == n and rep.n_index_left == 0
    ks = h.kernel_stats()
else:
    run = fqdist.ShardedFastqInfo(fq.MODE_INDEX_PAIR, device=0, n_hint=n)
    for _ in range(2):
        res = run.run_device(f1.data_ptr(), n * rb, name="a.fq", ptr2=f2.data_ptr(), nbytes2=n * rb, name2="b.fq")
        assert res["event_key"] == (1 << 64) - 1 and run.exact_reruns == 0, (res["event_key"], run.exact_reruns)
    ks = run.shard.kernel_stats()
    ks["other"] = run.ctx.kernel_stats()["other"]
for k in ("index", "mate", "other"):
    if ks[k]["launches"]:
        print(f"{which} {k}: {ks[k]['ms'] / ks[k]['launches']:.4f} ms/launch x{ks[k]['launches']}  {ks[k]['items'] / ks[k]['ms'] / 1e3:.1f} M items/s")

"""Short driver for ncu: 6 M synthetic Illumina records (2.15 GB, one chunk) through the -r loop, a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_utils_b200 as fq

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_900_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = fq.MODE_INDEX if len(sys.argv) > 3 and sys.argv[3] == "index" else fq.MODE_SINGLE
rb = fq.illumina_record_bytes()
t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
fq.synth_illumina(t, 0, n, seed=42, mate=1, perm_window=0, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
h = fq.FastqInfo(mode, index_capacity_hint=n)
for _ in range(reps):
    h.reset()
    h.feed_device(0, t.data_ptr(), n * rb, last=True)
    rep = h.finish()
    assert rep.error.code == 0, rep.error.code
ks = h.kernel_stats()
pc = h.path_counts()
for k in ("lanes", "tile", "records", "scan", "index"):
    if ks[k]["launches"]:
        print(f"{k}: {ks[k]['ms'] / ks[k]['launches']:.4f} ms/launch x{ks[k]['launches']}" + (f"  {ks[k]['bytes'] / ks[k]['ms'] / 1e6:.1f} GB/s" if ks[k]["bytes"] else ""))
print("paths", pc)

"""A/B of the clean-data pass under FQG_LANES_TUNE values: 17.7 M records (3 chunks, 6.4 GB) through -r (the pass alone) and through the
default mode (index kernel beside it); prints ms per launch of the pass for every value."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_utils_b200 as fq
vals = sys.argv[1:] or ["0", "2"]
n = 17_700_000
rb = fq.illumina_record_bytes()
t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
for s in range(0, n, 5_900_000):
    fq.synth_illumina(t[s * rb:], s, min(5_900_000, n - s), seed=42, mate=1, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
for mode, nm in ((fq.MODE_SINGLE, "alone (-r)"), (fq.MODE_INDEX, "in job (index)")):
    for v in vals * 2:
        os.environ["FQG_LANES_TUNE"] = v
        h = fq.FastqInfo(mode, index_capacity_hint=n)
        for i in range(4):
            if i == 1:
                h.kernel_stats(reset=True)
            h.reset()
            h.feed_device(0, t.data_ptr(), n * rb, last=True)
            rep = h.finish()
            assert rep.error.code == 0
        ks = h.kernel_stats()["lanes"]
        print(f"{nm:16s} tune={v}: {ks['ms'] / ks['launches']:.4f} ms/launch  {ks['bytes'] / ks['ms'] / 1e6:.0f} GB/s")
        h.close()

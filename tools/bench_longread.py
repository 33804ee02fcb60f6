"""BASELINE.json config 5 on one GPU: synthetic long reads (1-100 kb, log-normal, seed 7) through the -r loop, input resident in HBM.
Prints one JSON line (GB/s, reads/s, kernel times, which path validated the chunks) and checks the statistics against the lengths."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_utils_b200 as fq

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 12.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(7)
hdr = fq.lib().fqg_synth_long_header_bytes()
nrec = int(gb * 1e9 / (2 * 13_000 + hdr))  # mean of the clipped log-normal is about 13 kb
lens = torch.exp(torch.randn(nrec, generator=g) + 8.987).clamp(1000, 100000).to(torch.int64)  # ln 8000 = 8.987
off = torch.zeros(nrec + 1, dtype=torch.int64)
off[1:] = torch.cumsum(hdr + 2 * lens + 4, 0)
nb = int(off[-1])
t = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
offd = off.cuda()
fq.synth_longreads(t, offd, 0, nrec, seed=7, stream=st)  # record r of the call lands at offsets[r] - offsets[0]
torch.cuda.synchronize()
h = fq.FastqInfo(fq.MODE_SINGLE)
def step():
    h.reset()
    h.feed_device(0, t.data_ptr(), nb, last=True)
    return h.finish()
for _ in range(2):
    rep = step()
h.kernel_stats(reset=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    rep = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
assert rep.error.code == 0 and rep.file[0].num_rds == nrec, (rep.error.code, rep.file[0].num_rds)
assert (rep.file[0].min_rl, rep.file[0].max_rl) == (int(lens.min()) + 1, int(lens.max()) + 1)
srt = torch.sort(lens).values
ks = h.kernel_stats()
print(json.dumps({"workload": "longread_skew_seed7", "mode": "-r", "bytes": nb, "records": nrec, "GBps": nb / dt / 1e9, "reads_per_s": nrec / dt, "ms_per_step": dt * 1e3,
                  "median_rl": rep.median_rl, "paths": h.path_counts(), "kernel_ms_per_step": {k: v["ms"] / steps for k, v in ks.items() if v["launches"]},
                  "lanes_GBps": ks["lanes"]["bytes"] / ks["lanes"]["ms"] / 1e6 if ks["lanes"]["ms"] else None}))

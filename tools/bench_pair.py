"""BASELINE.json config 4 at one-GPU scale: paired Illumina files (mates permuted inside windows of 1024) through the default
two-file mode (index loop over file 1, mate loop over file 2), inputs resident in HBM.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_utils_b200 as fq

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rb = fq.illumina_record_bytes()
st = torch.cuda.current_stream().cuda_stream
files = []
for mate, perm in ((1, 0), (2, 1024)):
    t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
    for s in range(0, n, 8_192_000):
        k = min(8_192_000, n - s)
        fq.synth_illumina(t[s * rb:], s, k, seed=43, mate=mate, perm_window=perm, stream=st)
    files.append(t)
torch.cuda.synchronize()
h = fq.FastqInfo(fq.MODE_INDEX_PAIR, index_capacity_hint=n)
def step():
    h.reset()
    h.feed_device(0, files[0].data_ptr(), n * rb, last=True)
    h.feed_device(1, files[1].data_ptr(), n * rb, last=True)
    return h.finish()
for _ in range(2):
    rep = step()
h.kernel_stats(reset=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps):
    rep = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
assert rep.error.code == 0 and rep.n_index_entries == n, (rep.error.code, rep.n_index_entries)
ks = h.kernel_stats()
print(json.dumps({"workload": f"illumina_pe_{n // 1_000_000}M_pairs_2x150", "mode": "default, two files (index + mate loop)", "bytes": 2 * n * rb, "pairs": n,
                  "GBps": 2 * n * rb / dt / 1e9, "reads_per_s": 2 * n / dt, "ms_per_step": dt * 1e3, "paths": h.path_counts(),
                  "kernel_ms_per_step": {k: round(v["ms"] / steps, 3) for k, v in ks.items() if v["launches"]}}))

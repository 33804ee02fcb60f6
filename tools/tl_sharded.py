"""Development aid: the sharded two-file job with a world of one (same kernels and stream choreography as N ranks, no peers),
timed per phase, with the launch timeline of the last job (FQG_TIMELINE) summarised per round.
usage: python tools/tl_sharded.py [pairs] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fastq_utils_b200 as fq  # noqa: E402
from fastq_utils_b200 import dist as fqdist  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152 * 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tl = "gpurun_out/timeline.txt"
os.makedirs("gpurun_out", exist_ok=True)
rb = fq.illumina_record_bytes()
st = torch.cuda.current_stream().cuda_stream
f1 = torch.empty(n * rb + 64, dtype=torch.uint8, device="cuda")
f2 = torch.empty(n * rb + 64, dtype=torch.uint8, device="cuda")
fq.synth_illumina(f1, 0, n, seed=43, mate=1, stream=st)
fq.synth_illumina(f2, 0, n, seed=43, mate=2, perm_window=1024, stream=st)
torch.cuda.synchronize()
run = fqdist.ShardedFastqInfo(fq.MODE_INDEX_PAIR, device=0, n_hint=n)
for k in range(steps):
    if k == steps - 1:
        run.ctx.kernel_stats(); run.shard.kernel_stats()  # collect what is pending
        os.environ["FQG_TIMELINE"] = tl
        if os.path.exists(tl + ".0"):
            os.remove(tl + ".0")
    run.phase_ms = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = run.run_device(f1.data_ptr(), n * rb, name="a.fq", ptr2=f2.data_ptr(), nbytes2=n * rb, name2="b.fq")
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    print(f"step {k}: {ms:.2f} ms  rc {res['transcript'][0]} reruns {run.exact_reruns} rounds {run.rounds_done} phases", {a: round(b, 2) for a, b in run.phase_ms.items()}, flush=True)
run.ctx.kernel_stats(); run.shard.kernel_stats()
os.environ.pop("FQG_TIMELINE")
names = list(fq.KERNEL_CLASSES)
rows = []
for line in open(tl + ".0"):
    if line.startswith("#"):
        continue
    c, s, a, b = line.split()
    rows.append((float(a), float(b), names[int(c)], s))
rows.sort()
t00 = rows[0][0]
for a, b, c, s in rows:
    print(f"{a - t00:8.3f} {b - t00:8.3f} {b - a:7.3f}  {c:8s} {s}")

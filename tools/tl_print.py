"""Print a launch timeline written with FQG_TIMELINE (file per device): python tools/tl_print.py gpurun_out/timeline.txt.0 [last_ms]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastq_utils_b200.api import KERNEL_CLASSES  # noqa: E402
rows = []
for line in open(sys.argv[1]):
    if line.startswith("#"):
        continue
    c, s, a, b = line.split()
    rows.append((float(a), float(b), list(KERNEL_CLASSES)[int(c)], s))
rows.sort()
last = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
t1 = rows[-1][1]
rows = [r for r in rows if r[0] >= t1 - last]
t0, prev_end = rows[0][0], rows[0][0]
for a, b, c, s in rows:
    print(f"{a - t0:8.3f} {b - t0:8.3f} {b - a:7.3f}  gap {a - prev_end:6.3f}  {c:8s} {s}")
    prev_end = max(prev_end, b)

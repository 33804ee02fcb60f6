"""Attribute the per-instruction counts of an ncu source page to source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]

ncu's CSV source page lists SASS instructions in address order; nvdisasm -g on the cubin inside libfastq_gpu.so lists
the same instructions with '//## File "...", line N' markers (needs -lineinfo).  Matching by position gives executed warp
instructions and stall samples per source line (inlined callees are attributed to the innermost line)."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
mangled = kern  # substring of the mangled name selecting the function in the cubin ("name@mangled" to give both)
if "@" in kern:
    kern, mangled = kern.split("@", 1)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "fastq_utils_b200", "libfastq_gpu.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, active, where = None, False, ("?", 0)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = mangled in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            where = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((where, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# find the block of the wanted kernel
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and kern in r[1]:
        start = i
        break
hdr = rows[start + 1]
body = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        body.append(r)
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
print(f"disasm {len(lines)} ncu {len(body)}")
n = min(len(lines), len(body))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for k in range(n):
    w = lines[k][0]
    a = agg[w]
    a[0] += int(body[k][ci]); a[1] += int(body[k][cs]); a[2] += int(body[k][ct])
    tot[0] += int(body[k][ci]); tot[1] += int(body[k][cs]); tot[2] += int(body[k][ct])
print(f"total warp inst {tot[0]}  thread/warp {tot[2] / max(1, tot[0]):.1f}  samples {tot[1]}")
src_cache = {}
def src(w):
    f, l = w
    for d in ("fastq_utils_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < l <= len(src_cache[p]):
                return src_cache[p][l - 1].strip()[:110]
    return ""
for w, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / tot[0]:5.1f}% inst {100 * a[1] / max(1, tot[1]):5.1f}% stall  lanes {a[2] / max(1, a[0]):4.1f}  {w[0]}:{w[1]:4d}  {src(w)}")

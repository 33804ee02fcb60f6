"""Summarise an ncu report of fq_lanes_kernel by kernel phase (uses tools/ncu_lines.py) and print the headline counters."""
import csv, os, re, subprocess, sys
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "fq_lanes_kernel"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kname = kern.split("@")[0]
vals = [r for r in rows[2:] if any(kname in x for x in r)][0]
want = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h):
        try:
            if "stalled" in h and float(vals[i].replace(",", "")) < 0.1: continue
        except Exception: pass
        print(f"{h:92s} {vals[i]:>18s} {units[i]}")
src = open(os.path.join(root, "fastq_utils_b200/csrc/fq_lanes.cuh")).read().splitlines()
marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r"/\* ---- ([A-Z]\d?)[ :]", l)] if m]
def phase(line):
    cur = "prologue"
    for ln, name in marks:
        if line >= ln: cur = name
    return cur
out = subprocess.run([sys.executable, os.path.join(root, "tools/ncu_lines.py"), rep, kern, "100000"], capture_output=True, text=True).stdout
agg = {}
for ln in out.splitlines():
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% stall\s+lanes\s+([\d.]+)\s+(\S+):\s*(\d+)", ln)
    if not m: print(ln) if ln.startswith(("disasm", "total")) else None; continue
    pct, st, f, l = float(m.group(1)), float(m.group(2)), m.group(4), int(m.group(5))
    if f == "fq_lanes.cuh":
        key = phase(l) if l >= marks[0][0] - 60 and l > 110 else f"helpers:{l}"
    else: key = f"{f}"
    a = agg.setdefault(key, [0, 0]); a[0] += pct; a[1] += st
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]: print(f"{v[0]:6.1f}% inst {v[1]:6.1f}% stall  {k}")

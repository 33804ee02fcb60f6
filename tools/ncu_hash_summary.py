"""Summarise `ncu --set full` reports of the hash kernels: per kernel launch the duration, DRAM bytes / sectors, L2 sectors (all, atomics),
global atomics issued, occupancy and issue rate, and the derived random-access figures (ops/s, DRAM sectors per op).

    python tools/ncu_hash_summary.py <report.ncu-rep> [items-per-launch]"""
import csv, subprocess, sys
rep = sys.argv[1]
items = float(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__sectors_read.sum", "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print(f"== {r[ki]}")
    v = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            v[w] = r[i]
            print(f"  {w:86s} {r[i]:>18s} {units[i]}")
    try:
        def num(k):
            return float(v[k].replace(",", ""))
        i = hdr.index("gpu__time_duration.sum")
        t = num("gpu__time_duration.sum") * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(units[i], 1e-6)
        if items:
            print(f"  -> {items / t / 1e9:.2f} G ops/s under ncu (one launch, cold caches); DRAM sectors per op {(num('dram__sectors_read.sum') + num('dram__sectors_write.sum')) / items:.2f}; "
                  f"L2 sectors per op {num('lts__t_sectors.sum') / items:.2f}; global atomic sectors per op {num('l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum') / items:.2f}")
    except Exception as ex:
        print("  (derived figures unavailable:", ex, ")")

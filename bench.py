#!/usr/bin/env python
"""bench.py — fastq_info hot path throughput on B200 (BASELINE.json metric: FASTQ GB/s and reads/s validated).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--workload illumina_se|longread]

One step = one complete fastq_info job (default mode: parse + validate + read-name uniqueness) over the synthetic
Illumina single-end stream of BASELINE.json configs[2] (100 M × 150 bp, 35.9 GB), input already resident in HBM.
`value` is GB/s of decompressed FASTQ over the whole job; `e2e` is the same job fed from pinned HOST memory through
fqg_feed (H2D copy inside the timed region) with the report read back.  The reference arm times the unmodified
reference binary (oracle/_ref/fastq_info, 1 thread — it has no threading) on a bounded sample of the same bytes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            v = d.get("hbm_gbs") or d.get("hbm_gb_s")
            if v:
                return float(v), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "fastq_info")
    return p if os.path.exists(p) else None


def time_reference(sample_bytes, argv_tail, repeats=1):
    """Wall time of the reference's own CPU implementation on `sample_bytes` (plain text, so inflate is excluded on both sides)."""
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(shm, f"fqg_bench_{os.getpid()}.fastq")
    with open(path, "wb") as fh:
        fh.write(sample_bytes)
    try:
        ref = reference_binary()
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            if ref:
                p = subprocess.run([ref] + argv_tail + [path], capture_output=True)
                rc, kind = p.returncode, "reference"
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from _util import oracle_run
                rc, kind = oracle_run(argv_tail + ["a.fq"], sample_bytes, None)[0], "port"
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, rc, kind
    finally:
        os.unlink(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=100_000_000, help="records per GPU (weak scaling)")
    ap.add_argument("--workload", default="illumina_se")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")

    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes() if a.impl != "reference" or True else 359
    config = {"workload": f"illumina_se_{a.reads // 1_000_000}M_150bp_uniqueness_on", "reads_per_gpu": a.reads, "record_bytes": rb,
              "mode": "fastq_info default (index + validate)", "l2": "inputs (35.9 GB) far larger than L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm: the reference's CPU implementation
    if a.impl == "reference":
        if rank != 0:
            return
        torch.cuda.set_device(0)
        n = a.cpu_sample_reads
        t = torch.empty(n * rb + 64, dtype=torch.uint8, device="cuda")
        fq.synth_illumina(t, 0, n, seed=42, mate=1, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        sample = t[:n * rb].cpu().numpy().tobytes()
        times = []
        for i in range(a.warmup + a.steps):
            dt, rc, kind = time_reference(sample, [])
            assert rc == 0, rc
            if i >= a.warmup:
                times.append(dt)
        per = sum(times) / len(times)
        gbs = n * rb / per / 1e9
        print(json.dumps({"impl": "reference", "metric": "fastq_info_validated_GBps", "value": gbs, "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic", "config": config, "reads_per_s": n / per,
                          "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind, "sample": f"first {n} records ({n * rb / 1e6:.0f} MB plain text) of the workload, wall clock, host nproc={os.cpu_count()}"},
                          "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    n = a.reads
    nb = n * rb
    config["l2"] = f"inputs ({nb / 1e9:.1f} GB per GPU) {'far larger than' if nb > (1 << 30) else 'NOT larger than'} L2; no flush between steps"
    data = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
    first = rank * n
    piece = 8_000_000
    for s in range(0, n, piece):
        k = min(piece, n - s)
        fq.synth_illumina(data[s * rb:], first + s, k, seed=42, mate=1, stream=stream)
    data[nb:].zero_()
    torch.cuda.synchronize()

    if world > 1:
        from fastq_utils_b200 import dist as fqdist
        runner = fqdist.ShardedFastqInfo(fq.MODE_INDEX, local, n_hint=n)
        ctx = runner.ctx

        def step():
            res = runner.run_device(data.data_ptr(), nb, name="synthetic.fastq")
            assert res["event_key"] == (1 << 64) - 1 and res["n_index_entries"] == n * world, (res["event_key"], res["n_index_entries"])
            return res
    else:
        ctx = fq.FastqInfo(fq.MODE_INDEX, device=local, index_capacity_hint=n)

        def step():
            ctx.reset()
            ctx.feed_device(0, data.data_ptr(), nb, last=True)
            rep = ctx.finish()
            assert rep.error.code == 0 and rep.n_index_entries == n, (rep.error.code, rep.n_index_entries)
            return rep

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    ctx.kernel_stats(reset=True)
    if world > 1:
        runner.shard.kernel_stats(reset=True)  # (the index shard's kernels were counted from the first warm-up step before)
    l0 = ctx.launch_count() + (runner.shard.launch_count() if world > 1 else 0)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()  # (torch's stream is idle here: the event marks the device's clock at the start of the timed region)
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(a.steps):
        rep = step()
        dev_ms += ctx.device_ms()
    barrier()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    span_ms = ev0.elapsed_time(ev1)  # device clock over the whole timed region, every stream's work and every host gap included
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = ctx.launch_count() - l0
    ks = ctx.kernel_stats()
    if world > 1 and os.environ.get("FQG_DEBUG_ROUTE"):
        print(f"[rank {rank}] kernel ms per step:", {k: round(v["ms"] / a.steps, 2) for k, v in ks.items() if v["launches"]},
              "index", round(runner.shard.kernel_stats()["index"]["ms"] / a.steps, 2), "wall", round(wall / a.steps * 1e3, 2), file=sys.stderr)
        print(f"[rank {rank}] host ms per step inside the routing rounds:", {k: round(v / (a.steps + a.warmup), 2) for k, v in runner.host_ms.items()}, "rounds", runner.rounds_done, "p2p", runner._p2p_ok, file=sys.stderr)
    if world > 1:  # the index shard lives in its own context
        launches += runner.shard.launch_count()
        ks["index"] = runner.shard.kernel_stats()["index"]
        dev_ms = span_ms  # the sharded step runs on several streams of two contexts: the span between the two events is its device time
    # the step is host-driven (several synchronising read-backs); device-event time and wall time are both reported, the larger one counts
    per_step = max(dev_ms / 1e3, wall) / a.steps
    if world > 1:
        tt = torch.tensor([per_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per_step = float(tt.item())
    total_bytes = nb * world
    value = total_bytes / per_step / 1e9

    # dominant kernel and its roofline
    peak, peak_kind = measured_peak()
    dom = max(["scan", "records", "tile", "lanes"], key=lambda k: ks[k]["ms"])
    kd = ks[dom]
    ach = kd["bytes"] / (kd["ms"] / 1e3) / 1e9 if kd["ms"] > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of the same kernel on a
    # 2.118 GB chunk (profiles/r1_lanes_perline_*_ncu_summary.txt), scaled to this run's bytes per launch
    traffic = None
    try:
        import glob
        import re
        if dom == "lanes":
            f = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_lanes_perline_v*_ncu_summary.txt")))[-1]
            txt = open(f, errors="ignore").read()
            rd = float(re.search(r"dram__bytes_read\.sum\s+([\d.]+)\s+Gbyte", txt).group(1))
            wr = float(re.search(r"dram__bytes_write\.sum\s+([\d.]+)\s+Mbyte", txt).group(1)) / 1e3
            traffic = (rd + wr) * 1e9 / 2.1181e9 * (kd["bytes"] / max(1, kd["launches"]))
    except Exception:
        traffic = None
    roof = {"bound": "hbm", "kernel": {"scan": "fq_scan_kernel", "records": "fq_records_kernel", "tile": "fq_tile_kernel", "lanes": "fq_lanes_kernel"}[dom], "achieved": ach, "peak": peak, "peak_kind": peak_kind,
            "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write of the committed capture, scaled by launch size)", "launches": kd["launches"], "avg_launch_ms": kd["ms"] / max(1, kd["launches"]),
            "algorithmic_bytes_per_launch": kd["bytes"] / max(1, kd["launches"]),
            "all_kernels_ms_per_step": {k: v["ms"] / a.steps for k, v in ks.items()},
            "parse_validate_GBps": nb * a.steps / (sum(ks[k]["ms"] for k in ("scan", "records", "tile", "lanes")) / 1e3) / 1e9 if sum(ks[k]["ms"] for k in ("scan", "records", "tile", "lanes")) > 0 else None,
            "index_Mops_per_s": ks["index"]["items"] / (ks["index"]["ms"] / 1e3) / 1e6 if ks["index"]["ms"] > 0 else None}

    out = {"metric": "fastq_info_validated_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": config, "reads_per_s": n * world / per_step, "device_ms_per_step": dev_ms / a.steps, "wall_ms_per_step": wall / a.steps * 1e3,
           "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary()}
    if world > 1:
        how = ("chunk by chunk into the owners' peer memory over NVLink (CUDA IPC), beside the next chunk's pass" if getattr(runner, "_p2p_ok", False)
               else "chunk by chunk with all-to-all exchanges" if runner.pipeline else "one all-to-all over NCCL")
        out["config"]["parallelism"] = f"byte-range shards x{world}, names routed by hash {how}, stats all-reduced"
        out["config"]["routing_rounds"] = runner.rounds_done
        out["e2e"] = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "measured at N=1 only"}

    if rank == 0 and world == 1:
        # ---------------- the dominant kernel without the index kernel beside it: the same bytes through the -r loop (no name index)
        try:
            solo = fq.FastqInfo(fq.MODE_SINGLE, device=local)
            for i in range(3):
                if i == 1:
                    solo.kernel_stats(reset=True)
                solo.reset()
                solo.feed_device(0, data.data_ptr(), nb, last=True)
                r1 = solo.finish()
                assert r1.error.code == 0
            ks1 = solo.kernel_stats()[dom]
            solo.close()
            if ks1["ms"] > 0:
                a1 = ks1["bytes"] / (ks1["ms"] / 1e3) / 1e9
                out["roofline"]["alone"] = {"achieved": a1, "frac": a1 / peak, "avg_launch_ms": ks1["ms"] / max(1, ks1["launches"]),
                                            "note": "same kernel, same bytes, -r loop: no index kernel running beside it"}
        except Exception as ex:
            out["roofline"]["alone"] = {"error": str(ex)[:200]}
        # ---------------- e2e: the same job through fqg_feed from pinned host memory (H2D inside the timed region)
        if not a.no_e2e:
            try:
                host = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
                host.copy_(data[:nb])
                torch.cuda.synchronize()
                del data
                torch.cuda.empty_cache()
                chunk = 1 << 30
                def e2e_step():
                    ctx.reset()
                    for off in range(0, nb, chunk):
                        k = min(chunk, nb - off)
                        ctx.feed(0, (host.data_ptr() + off, k), last=(off + k == nb))
                    r = ctx.finish()
                    assert r.error.code == 0 and r.n_index_entries == n
                e2e_step()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                reps = max(1, min(a.steps, 2))
                for _ in range(reps):
                    e2e_step()
                torch.cuda.synchronize()
                e2e_t = (time.perf_counter() - t1) / reps
                out["e2e"] = {"value": nb / e2e_t / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": 4096,
                              "ms_per_step": e2e_t * 1e3, "reads_per_s": n / e2e_t, "note": "fqg_feed from pinned host memory in 1 GiB pieces + fqg_finish report"}
                sample_n = a.cpu_sample_reads
                sample = host[:sample_n * rb].numpy().tobytes()
            except Exception as ex:  # e.g. not enough pinned memory on this host
                out["e2e"] = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}
                sample = None
        else:
            sample = None
        # ---------------- CPU baseline: the unmodified reference binary on a bounded sample of the same bytes
        if sample is None:
            sample_n = a.cpu_sample_reads
            t = torch.empty(sample_n * rb + 64, dtype=torch.uint8, device="cuda")
            fq.synth_illumina(t, 0, sample_n, seed=42, mate=1, stream=stream)
            torch.cuda.synchronize()
            sample = t[:sample_n * rb].cpu().numpy().tobytes()
        dt, rc, kind = time_reference(sample, [])
        out["cpu_baseline"] = {"value": len(sample) / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": kind, "reads_per_s": sample_n / dt, "exit_status": rc,
                               "sample": f"first {sample_n} records ({len(sample) / 1e6:.0f} MB plain text) of the workload; reference is single-threaded; host nproc={os.cpu_count()}"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

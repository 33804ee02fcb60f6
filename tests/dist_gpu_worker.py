"""Run by torch.distributed.run on N GPUs (tests/test_gpu_dist.py): sharded fastq_info over NCCL vs the CPU oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fastq_utils_b200 as fq  # noqa: E402
from fastq_utils_b200 import dist as fqdist  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r, W = dist.get_rank(), dist.get_world_size()
    rb = fq.illumina_record_bytes()
    n_total = 120_000
    # the whole stream is generated everywhere (cheap) so that ranges can be cut at arbitrary byte positions
    whole = torch.zeros(n_total * rb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_illumina(whole, 0, n_total, seed=42, mate=1, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    results = []
    for variant in ("clean", "dup", "bad"):
        w = whole.clone()
        if variant == "dup":
            w[110_000 * rb:110_001 * rb] = w[3 * rb:4 * rb].clone()
        if variant == "bad":
            w[110_000 * rb:110_001 * rb] = w[3 * rb:4 * rb].clone()
            w[70_000 * rb + 80] = ord("*")
        nb = n_total * rb
        cuts = [0] + [int(nb * (i + 1) / W) + 37 * (i + 1) for i in range(W - 1)] + [nb]
        lo, hi = cuts[r], cuts[r + 1]
        mine = torch.zeros(hi - lo + 64, dtype=torch.uint8, device="cuda")
        mine[:hi - lo] = w[lo:hi]
        torch.cuda.synchronize()
        for mode, argv in ((fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])):
            run = fqdist.ShardedFastqInfo(mode, device=local, n_hint=n_total // W)
            res = run.run_device(mine.data_ptr(), hi - lo, name="a.fq")
            if r == 0:
                from _util import oracle_run
                want = oracle_run(argv, bytes(w[:nb].cpu().numpy()), None)
                results.append({"variant": variant, "argv": argv, "ok": tuple(res["transcript"]) == want, "got": res["transcript"][2][-200:], "want": want[2][-200:]})
            dist.barrier()
    # default two-file mode: index loop over file 1, mate loop over file 2 (mates permuted inside windows of 1024)
    n2 = 81_920
    f1 = torch.zeros(n2 * rb + 64, dtype=torch.uint8, device="cuda")
    f2 = torch.zeros(n2 * rb + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fq.synth_illumina(f1, 0, n2, seed=43, mate=1, stream=st)
    fq.synth_illumina(f2, 0, n2, seed=43, mate=2, perm_window=1024, stream=st)
    torch.cuda.synchronize()
    for variant in ("clean", "missing_mate", "extra_in_1"):
        a, b = f1[:n2 * rb].clone(), f2[:n2 * rb].clone()
        if variant == "missing_mate":
            b = torch.cat([b[:777 * rb], b[778 * rb:]])
        if variant == "extra_in_1":
            b = b[:(n2 - 2048) * rb]
        parts = []
        for t in (a, b):
            nb = t.numel()
            cuts = [0] + [int(nb * (i + 1) / W) + 53 * (i + 1) for i in range(W - 1)] + [nb]
            mine = torch.zeros(cuts[r + 1] - cuts[r] + 64, dtype=torch.uint8, device="cuda")
            mine[:cuts[r + 1] - cuts[r]] = t[cuts[r]:cuts[r + 1]]
            parts.append((mine, cuts[r + 1] - cuts[r]))
        torch.cuda.synchronize()
        run = fqdist.ShardedFastqInfo(fq.MODE_INDEX_PAIR, device=local, n_hint=n2 // W)
        res = run.run_device(parts[0][0].data_ptr(), parts[0][1], name="a.fq", ptr2=parts[1][0].data_ptr(), nbytes2=parts[1][1], name2="b.fq")
        if r == 0:
            from _util import oracle_run
            want = oracle_run(["a.fq", "b.fq"], bytes(a.cpu().numpy()), bytes(b.cpu().numpy()))
            results.append({"variant": "pair_" + variant, "argv": ["a.fq", "b.fq"], "ok": tuple(res["transcript"]) == want, "got": res["transcript"][2][-200:], "want": want[2][-200:]})
        dist.barrier()
    # interleaved and sorted-pair files are not sharded: the ranges are gathered on rank 0 (dist.py, _run_gathered)
    n3 = 20_480
    g1 = torch.zeros(n3 * rb + 64, dtype=torch.uint8, device="cuda")
    g2 = torch.zeros(n3 * rb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_illumina(g1, 0, n3, seed=44, mate=1, stream=st)
    fq.synth_illumina(g2, 0, n3, seed=44, mate=2, perm_window=0, stream=st)
    torch.cuda.synchronize()
    inter = torch.stack([g1[:n3 * rb].view(n3, rb), g2[:n3 * rb].view(n3, rb)], 1).reshape(-1).contiguous()
    jobs = [("interleaved", fq.MODE_INTERLEAVED, ["a.fq", "pe"], [inter]), ("sorted_pair", fq.MODE_SORTED_PAIR, ["-r", "-s", "a.fq", "b.fq"], [g1[:n3 * rb], g2[:n3 * rb]]),
            ("sorted_pair_permuted", fq.MODE_SORTED_PAIR, ["-r", "-s", "a.fq", "b.fq"], [f1[:n3 * rb], f2[:n3 * rb]])]
    for variant, mode, argv, files in jobs:
        parts = []
        for t in files:
            nb = t.numel()
            cuts = [0] + [int(nb * (i + 1) / W) + 41 * (i + 1) for i in range(W - 1)] + [nb]
            mine = torch.zeros(cuts[r + 1] - cuts[r] + 64, dtype=torch.uint8, device="cuda")
            mine[:cuts[r + 1] - cuts[r]] = t[cuts[r]:cuts[r + 1]]
            parts.append((mine, cuts[r + 1] - cuts[r]))
        torch.cuda.synchronize()
        run = fqdist.ShardedFastqInfo(mode, device=local)
        kw = {"ptr2": parts[1][0].data_ptr(), "nbytes2": parts[1][1], "name2": "b.fq"} if len(parts) > 1 else {}
        res = run.run_device(parts[0][0].data_ptr(), parts[0][1], name="a.fq", **kw)
        if r == 0:
            from _util import oracle_run
            want = oracle_run(argv, bytes(files[0].cpu().numpy()), bytes(files[1].cpu().numpy()) if len(files) > 1 else None)
            results.append({"variant": variant, "argv": argv, "ok": tuple(res["transcript"]) == want, "got": res["transcript"][2][-200:], "want": want[2][-200:]})
        dist.barrier()
    if r == 0:
        json.dump(results, open(sys.argv[1], "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

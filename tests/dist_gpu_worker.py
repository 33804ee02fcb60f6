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
    if r == 0:
        json.dump(results, open(sys.argv[1], "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

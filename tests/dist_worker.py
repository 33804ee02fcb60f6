"""Worker for tests/test_dist_gloo.py: run by torch.distributed.run with N ranks on CPU (gloo).  Uses the sequential
stand-in device (tests/sim) behind the same C ABI and the same fastq_utils_b200.dist orchestration as the GPU path."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fastq_utils_b200 as fq  # noqa: E402
from sim_lib import use_sim_library  # noqa: E402
from fastq_utils_b200 import dist as fqdist  # noqa: E402


def main():
    use_sim_library()
    cases = json.load(open(sys.argv[1]))
    dist.init_process_group("gloo")
    r, W = dist.get_rank(), dist.get_world_size()
    out, rounds, peer, gathered, runners = [], [], [], [], {}
    order = list(range(len(cases)))
    if os.environ.get("FQG_TEST_REUSE_RUNNER"):  # small jobs first, so that the reused runner's arena has to grow
        order.sort(key=lambda i: len(cases[i]["hex"]))
    grown = 0
    for ci in order:
        c = cases[ci]
        data = bytes.fromhex(c["hex"])
        cuts = [int(len(data) * x) for x in c["cuts"]][:W - 1]
        cuts = [0] + sorted(cuts) + [len(data)]
        while len(cuts) < W + 1:
            cuts.insert(-1, cuts[-1])
        lo, hi = cuts[r], cuts[r + 1]
        mine = bytearray(data[lo:hi]) + bytearray(64)
        buf = (ctypes.c_uint8 * len(mine)).from_buffer(mine)
        mode = {"single": fq.MODE_SINGLE, "index": fq.MODE_INDEX, "pair": fq.MODE_INDEX_PAIR, "interleaved": fq.MODE_INTERLEAVED, "sorted": fq.MODE_SORTED_PAIR}[c["mode"]]
        kw = {}
        if mode in (fq.MODE_INDEX_PAIR, fq.MODE_SORTED_PAIR):
            data2 = bytes.fromhex(c["hex2"])
            cuts2 = [0] + sorted(int(len(data2) * x) for x in c["cuts2"])[:W - 1] + [len(data2)]
            while len(cuts2) < W + 1:
                cuts2.insert(-1, cuts2[-1])
            mine2 = bytearray(data2[cuts2[r]:cuts2[r + 1]]) + bytearray(64)
            buf2 = (ctypes.c_uint8 * len(mine2)).from_buffer(mine2)
            kw = {"ptr2": ctypes.addressof(buf2), "nbytes2": cuts2[r + 1] - cuts2[r], "name2": "b.fq"}
        try:
            if os.environ.get("FQG_TEST_REUSE_RUNNER"):  # one runner per mode for all cases: arenas are reused and regrown between jobs
                if mode not in runners:
                    runners[mode] = fqdist.ShardedFastqInfo(mode, device=0, tensor_device=torch.device("cpu"))
                run = runners[mode]
            else:
                run = fqdist.ShardedFastqInfo(mode, device=0, tensor_device=torch.device("cpu"))
            res = run.run_device(ctypes.addressof(buf), hi - lo, name="a.fq", **kw)
            tr = res.get("transcript")
            if run._arena is not None:
                grown += int(run._arena[1] != getattr(run, "_seen_arena", run._arena[1]))
                run._seen_arena = run._arena[1]
            rounds.append(run.rounds_done if not os.environ.get("FQG_TEST_SLOT_CAP") else run.exact_reruns)
            peer.append(bool(run._p2p_ok))
            gathered.append(run.gathered_jobs)
        except (NotImplementedError, RuntimeError) as ex:
            tr = ["EXC", type(ex).__name__, str(ex)]
            rounds.append(-1)
            peer.append(False)
            gathered.append(-1)
        if r == 0:
            out.append(list(tr))
        dist.barrier()
    if r == 0:
        back = {ci: k for k, ci in enumerate(order)}  # results in the order of the case list
        out, rounds, peer, gathered = [[x[back[i]] for i in range(len(cases))] for x in (out, rounds, peer, gathered)]
        json.dump({"transcripts": out, "rounds": rounds, "peer": peer, "gathered": gathered, "arena_regrown": grown}, open(sys.argv[2], "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

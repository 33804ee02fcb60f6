"""Sharded fastq_info on every visible GPU over NCCL (skipped with fewer than 2 GPUs): transcripts equal the oracle's."""
import json
import os
import subprocess
import sys

import pytest

from _util import ROOT

pytestmark = pytest.mark.gpu


def test_nccl_sharded_matches_oracle(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = tmp_path / "out.json"
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                           "--master-port", "29701", os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(out)], timeout=900)
    res = json.load(open(out))
    assert len(res) == 9
    for x in res:
        assert x["ok"], x


@pytest.mark.parametrize("variant", ["clean", "dup", "bad", "clean_exchange", "overflow"])
def test_pipelined_routing_one_gpu(variant, monkeypatch):
    """The chunk-by-chunk routing of the sharded index (chunk hook beside the clean-data pass, fixed-capacity regions, owner-side
    insert on the side stream) with a world of one: same kernels and stream choreography as N ranks, the exchange is a no-op."""
    import torch
    import fastq_utils_b200 as fq
    from fastq_utils_b200 import dist as fqdist
    from _util import oracle_run
    monkeypatch.setenv("FQG_MAX_CHUNK_BYTES", str(24 << 20))  # ~9 chunks: the hook fires beside every pass after the first
    if variant == "clean_exchange":  # the rounds as exchanges of packed send buffers instead of stores into the owner's arena
        monkeypatch.setenv("FQG_P2P", "0")
        variant = "clean"
    overflow = variant == "overflow"
    if overflow:  # regions of 1000 tuples for 65 000 names a round: the owner reports the overflow, the exact path redoes the job
        monkeypatch.setenv("FQG_TEST_SLOT_CAP", "1000")
        variant = "clean"
    rb = fq.illumina_record_bytes()
    n = 600_000
    t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_illumina(t, 0, n, seed=42, mate=1, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if variant in ("dup", "bad"):
        t[550_000 * rb:550_001 * rb] = t[3 * rb:4 * rb].clone()
    if variant == "bad":
        t[300_000 * rb + 80] = ord("*")
    torch.cuda.synchronize()
    run = fqdist.ShardedFastqInfo(fq.MODE_INDEX, device=0, n_hint=n)
    res = run.run_device(t.data_ptr(), n * rb, name="a.fq")
    want = oracle_run(["a.fq"], bytes(t[:n * rb].cpu().numpy()), None)
    assert tuple(res["transcript"]) == want
    assert run.exact_reruns == (0 if variant == "clean" and not overflow else 1)
    if variant == "clean" and not overflow:
        assert run.rounds_done >= 5 and run._p2p_ok == (os.environ.get("FQG_P2P", "1") != "0")
        assert res["n_index_entries"] == n and run.ctx.path_counts()["lanes"] >= 8
        # a second job on the same objects (bench loop): the table is cleared, the rounds start over
        res2 = run.run_device(t.data_ptr(), n * rb, name="a.fq")
        assert tuple(res2["transcript"]) == want

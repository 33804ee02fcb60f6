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
    assert len(res) == 12
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
    if overflow:  # stretches of 100 slots per CTA: a CTA cannot take a second tile, tiles stay unvalidated, the pass rejects the chunk
        monkeypatch.setenv("FQG_TEST_SLOT_CAP", "100")  # and the exact path redoes the job
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


@pytest.mark.parametrize("variant", ["clean", "missing_mate", "extra_in_1", "dup_in_1", "weak_hash", "pack_mode"])
def test_pipelined_pair_routing_one_gpu(variant, monkeypatch):
    """Two files through the sharded path with a world of one: the clean-data pass writes every name (with its bytes) into the
    owner's region, the owner inserts file 1's and lets file 2's claim them; anything but a clean job is redone by the exact path."""
    import torch
    import fastq_utils_b200 as fq
    from fastq_utils_b200 import dist as fqdist
    from _util import oracle_run
    monkeypatch.setenv("FQG_MAX_CHUNK_BYTES", str(24 << 20))
    if variant == "weak_hash":  # 12-bit hashes: the owner meets different names with equal hashes all the time and must walk on
        monkeypatch.setenv("FQG_TEST_WEAK_HASH", "1")
    if variant == "pack_mode":  # the names packed by a kernel beside the pass instead of written by the pass
        monkeypatch.setenv("FQG_ROUTE_IN_PASS", "0")
    rb = fq.illumina_record_bytes()
    n = 300 * 1024
    st = torch.cuda.current_stream().cuda_stream
    f1 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
    f2 = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_illumina(f1, 0, n, seed=43, mate=1, stream=st)
    fq.synth_illumina(f2, 0, n, seed=43, mate=2, perm_window=1024, stream=st)
    torch.cuda.synchronize()
    n1 = n2 = n
    if variant == "missing_mate":
        f2[777 * rb:(n - 1) * rb] = f2[778 * rb:n * rb].clone()
        n2 = n - 1
    if variant == "extra_in_1":
        n2 = n - 2048
    if variant == "dup_in_1":
        f1[250_000 * rb:250_001 * rb] = f1[3 * rb:4 * rb].clone()
    torch.cuda.synchronize()
    run = fqdist.ShardedFastqInfo(fq.MODE_INDEX_PAIR, device=0, n_hint=n)
    res = run.run_device(f1.data_ptr(), n1 * rb, name="a.fq", ptr2=f2.data_ptr(), nbytes2=n2 * rb, name2="b.fq")
    want = oracle_run(["a.fq", "b.fq"], bytes(f1[:n1 * rb].cpu().numpy()), bytes(f2[:n2 * rb].cpu().numpy()))
    assert tuple(res["transcript"]) == want, (res["transcript"][2][-300:], want[2][-300:])
    clean = variant in ("clean", "weak_hash", "pack_mode")
    assert run.exact_reruns == (0 if clean else 1)
    if clean:
        assert run.rounds_done >= 8 and run._plan[0]["mode"] == ("pack" if variant == "pack_mode" else "pass")
        res2 = run.run_device(f1.data_ptr(), n1 * rb, name="a.fq", ptr2=f2.data_ptr(), nbytes2=n2 * rb, name2="b.fq")
        assert tuple(res2["transcript"]) == want

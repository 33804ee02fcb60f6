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

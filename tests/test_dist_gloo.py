"""Multi-rank orchestration (fastq_utils_b200/dist.py) on CPU: world_size 2 and 3 under gloo, with the stand-in device.
Byte ranges are cut at arbitrary positions (inside headers, sequences, qualities); the rank-0 transcript must equal the
CPU oracle's for the whole file."""
import json
import os
import random
import subprocess
import sys

import pytest

from _util import GOLDEN, ROOT, oracle_run, read_stream


def _cases():
    rng = random.Random(5)
    files = ["c18_10000_1.fastq.gz", "test_1.fastq.gz", "test_e9.fastq.gz", "test_e3.fastq.gz", "test_21_1.fastq.gz", "test_e5.fastq.gz",
             "casava.1.8_readname_trunc_1.err2.fastq.gz", "edge_dup3.fastq", "edge_crlf.fastq", "edge_lens.fastq", "edge_tail2.fastq", "edge_no_trailing_nl.fastq"]
    cases = []
    for f in files:
        data = read_stream(os.path.join(GOLDEN, "inputs", f))
        for mode in ("index", "single"):
            cases.append({"file": f, "mode": mode, "hex": data.hex(), "cuts": sorted([rng.random(), rng.random()])})
    # synthetic: a duplicate far apart, split anywhere
    recs = [f"@M0:1:FC:1:11:{i}:{i * 7} 1:N:0:AC\n{'ACGTN' * (3 + i % 5)}\n+\n{'F' * (5 * (3 + i % 5))}\n" for i in range(400)]
    recs[390] = recs[17]
    data = "".join(recs).encode()
    for cuts in ([0.1, 0.5], [0.5, 0.97], [0.045, 0.046]):
        cases.append({"file": "synthetic_dup", "mode": "index", "hex": data.hex(), "cuts": cuts})
        cases.append({"file": "synthetic_dup", "mode": "single", "hex": data.hex(), "cuts": cuts})
    return cases


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_transcripts_match_oracle(tmp_path, world):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim")], stdout=subprocess.DEVNULL)
    cases = _cases()
    cin, cout = tmp_path / "cases.json", tmp_path / "out.json"
    json.dump(cases, open(cin, "w"))
    port = 29600 + world
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), str(cin), str(cout)], env=env, timeout=600)
    got = json.load(open(cout))
    assert len(got) == len(cases)
    for c, g in zip(cases, got):
        argv = (["-r"] if c["mode"] == "single" else []) + ["a.fq"]
        want = oracle_run(argv, bytes.fromhex(c["hex"]), None)
        assert tuple(g) == want, (c["file"], c["mode"], c["cuts"], g)

"""Progress lines (`\\b` x 15 + count, every 100 000 reads; src/fastq.h:82, src/fastq_info.c:139 counts lines/2 in the sorted-pair loop and
lines/4 of an 8-line step in the interleaved one) on files of 230 000 reads, in all five modes: the unmodified reference binary, the
oracle's restatement and the library must print the same bytes.  CPU: stand-in device; -m gpu: the CUDA kernels."""
import os
import tempfile

import pytest

from _util import REF_BIN, fqg_run, oracle_run, ref_run

N = 230_000


def _files():
    r1 = "".join(f"@R{i}/1\n{'ACGTN' * 4}\n+\n{'IIIIF' * 4}\n" for i in range(N)).encode()
    r2 = "".join(f"@R{i}/2\n{'TTGCA' * 4}\n+\n{'FFFFI' * 4}\n" for i in range(N)).encode()
    inter = "".join(f"@R{i}/1\n{'ACGTN' * 4}\n+\n{'IIIIF' * 4}\n@R{i}/2\n{'TTGCA' * 4}\n+\n{'FFFFI' * 4}\n" for i in range(N // 2)).encode()
    return r1, r2, inter


CASES = [("index", ["a.fq"], 0, None), ("single", ["-r", "a.fq"], 0, None), ("pair", ["a.fq", "b.fq"], 0, 1), ("sorted", ["-r", "-s", "a.fq", "b.fq"], 0, 1),
         ("interleaved", ["a.fq", "pe"], 2, None)]


@pytest.mark.parametrize("name,argv,i1,i2", CASES)
def test_progress_lines_reference_oracle_standin(name, argv, i1, i2):
    fs = _files()
    d1, d2 = fs[i1], (fs[i2] if i2 is not None else None)
    want = oracle_run(argv, d1, d2)
    assert want[0] == 0 and want[2].count("\b" * 15) == {"index": 2, "single": 2, "pair": 4, "sorted": 4, "interleaved": 2}[name], want[2][:300]
    if os.path.exists(REF_BIN):  # the reference itself, on plain-text files
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "a.fq"), "wb").write(d1)
            if d2 is not None:
                open(os.path.join(d, "b.fq"), "wb").write(d2)
            assert ref_run(argv, cwd=d) == want
    assert fqg_run(argv, d1, d2, chunk=1 << 22, kind="sim") == want


@pytest.mark.gpu
@pytest.mark.parametrize("name,argv,i1,i2", CASES)
def test_gpu_progress_lines(name, argv, i1, i2):
    fs = _files()
    d1, d2 = fs[i1], (fs[i2] if i2 is not None else None)
    assert fqg_run(argv, d1, d2, chunk=0, kind="gpu") == oracle_run(argv, d1, d2)

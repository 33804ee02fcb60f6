"""The command lines of the reader- and writer-style tools (fastq_utils_b200/fastq_tools_gpu and its links <tool>_gpu) against the
committed transcripts of the reference's binaries: same exit status, stdout, stderr and — gunzipped — the same output files."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

from _util import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "fastq_utils_b200")


def _same(text, c, key):
    if key in c:
        return text.decode("latin-1") == c[key]
    return (len(text), hashlib.sha256(text).hexdigest()) == (c[key + "_len"], c[key + "_sha256"])


READER = json.load(open(os.path.join(GOLDEN, "reader_transcripts.json")))
WRITER = json.load(open(os.path.join(GOLDEN, "writer_transcripts.json")))
TRIM = json.load(open(os.path.join(GOLDEN, "trim_transcripts.json")))
PAIR = json.load(open(os.path.join(GOLDEN, "filterpair_transcripts.json")))


@pytest.mark.parametrize("idx", range(0, len(READER), 67))
def test_reader_cli(idx):
    c = READER[idx]
    p = subprocess.run([os.path.join(BIN, c["tool"] + "_gpu")] + c["argv"], cwd=GOLDEN, capture_output=True)
    assert (p.returncode, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")) == (c["rc"], c["stdout"], c["stderr"]), (c["tool"], c["argv"])


@pytest.mark.parametrize("idx", range(0, len(WRITER), 131))
def test_writer_cli(idx):
    c = WRITER[idx]
    p = subprocess.run([os.path.join(BIN, c["tool"] + "_gpu")] + c["argv"], cwd=GOLDEN, capture_output=True)
    assert (p.returncode, p.stderr.decode("latin-1")) == (c["rc"], c["stderr"]), (c["tool"], c["argv"])
    assert _same(p.stdout, c, "stdout"), (c["tool"], c["argv"])


@pytest.mark.parametrize("idx", [i for i, c in enumerate(TRIM) if "trim_inputs/poly_100k.fq" not in " ".join(c["argv"])][::45])
def test_trim_cli(idx, tmp_path):
    c = TRIM[idx]
    out = str(tmp_path / "OUT")
    argv = [out if w == "OUT" else ("--outfile=" + out if w == "--outfile=OUT" else w) for w in c["argv"]]
    p = subprocess.run([os.path.join(BIN, "fastq_trim_poly_at_gpu")] + argv, cwd=GOLDEN, capture_output=True)
    assert (p.returncode, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")) == (c["rc"], c["stdout"], c["stderr"]), c["argv"]
    assert os.path.exists(out) == c["created"], c["argv"]
    if c["rc"] == 0 and c["created"]:
        assert _same(gzip.open(out, "rb").read(), c, "outfile"), c["argv"]


@pytest.mark.parametrize("idx", [i for i, c in enumerate(PAIR) if "pair_inputs/big_" not in " ".join(c["argv"])][::31])
def test_filterpair_cli(idx, tmp_path):
    c = PAIR[idx]
    names = {"P1.gz": str(tmp_path / "P1.gz"), "P2.gz": str(tmp_path / "P2.gz"), "UP.gz": str(tmp_path / "UP.gz")}
    argv = [names.get(w, w) for w in c["argv"]]
    p = subprocess.run([os.path.join(BIN, "fastq_filterpair_gpu")] + argv, cwd=GOLDEN, capture_output=True)
    assert (p.returncode, p.stdout.decode("latin-1")) == (c["rc"], c["stdout"]), c["argv"]
    assert _same(p.stderr, c, "stderr"), (c["argv"], p.stderr[-300:])
    assert all(os.path.exists(v) for v in names.values()) == c["created"], c["argv"]
    if b"\nPaired: " in p.stderr and ("out0" in c or "out0_len" in c):
        for k, w in enumerate(("P1.gz", "P2.gz", "UP.gz")):
            assert _same(gzip.open(names[w], "rb").read(), c, f"out{k}"), (c["argv"], k)

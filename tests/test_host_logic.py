"""Host-side logic of libfastq_gpu (chunking, bridging across feeds, over-long line re-segmentation, event ordering,
report, text rendering, option handling) on the sequential stand-in device of tests/sim/ — no GPU needed.  The same
engine / renderer / ABI sources are linked into the product with the CUDA device instead; the -m gpu tests repeat
these comparisons on the real kernels."""
import random

import pytest

from _util import fqg_run, fqg_run_files, golden_transcripts, oracle_run
from test_oracle_fuzz import NAMES, make_file, mutate, render

CASES = golden_transcripts()


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_sim_matches_reference_transcript(idx):
    c = CASES[idx]
    chunk = [0, 64, 1000, 7][idx % 4] if idx % 3 == 0 else 0
    if chunk == 7 and sum(len(a) for a in c["argv"]) and any("seq_" in a or "c18" in a for a in c["argv"]):
        chunk = 4096
    got = fqg_run_files(c["argv"], chunk=chunk, kind="sim")
    assert got == (c["rc"], c["stdout"], c["stderr"]), (c["argv"], chunk)


@pytest.mark.parametrize("seed", range(300))
def test_sim_fuzz_against_oracle(seed):
    rng = random.Random(10_000 + seed)
    style = rng.randrange(len(NAMES))
    n = rng.choice([0, 1, 2, 3, 5, 8, 13, 40])
    mode = rng.choice(["single", "single_r", "pe", "pair", "pair_rs", "pair_s", "pair_r"])
    r1 = make_file(rng, n, style, 1)
    two = mode.startswith("pair")
    r2 = make_file(rng, n, style, 2) if two else None
    if mode == "pe":
        inter = []
        for a, b in zip(r1, make_file(rng, n, style, 2)):
            inter += [a, b]
        r1 = inter
    if two and rng.random() < 0.5:
        rng.shuffle(r2)
    for _ in range(rng.choice([0, 0, 1, 1, 2])):
        mutate(rng, r1 if (not two or rng.random() < 0.5) else r2)
    nl = "crlf" if rng.random() < 0.08 else "lf"
    d1 = render(rng, r1, nl)
    d2 = render(rng, r2, nl) if two else None
    argv = {"single": [], "single_r": ["-r"], "pe": [], "pair": [], "pair_rs": ["-r", "-s"], "pair_s": ["-s"], "pair_r": ["-r"]}[mode]
    if rng.random() < 0.2: argv = ["-q"] + argv
    if rng.random() < 0.2: argv = ["-e"] + argv
    argv = argv + ["a.fq"] + (["b.fq"] if two else []) + (["pe"] if mode == "pe" else [])
    chunk = rng.choice([0, 0, 1, 5, 33, 200])
    assert fqg_run(argv, d1, d2, chunk=chunk, kind="sim") == oracle_run(argv, d1, d2), (argv, chunk, d1, d2)


def test_fast_record_path_equals_careful_path(tmp_path):
    """fq_record.h: the 16-byte-at-a-time path for clean records vs the byte-wise path (random records, all alignments)."""
    import os
    import subprocess
    from _util import ROOT
    exe = tmp_path / "test_record_paths"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "sim", "test_record_paths.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr

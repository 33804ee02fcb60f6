"""The oracle (oracle/fastq_oracle.c) against the reference's own behaviour.

tests/golden/transcripts.json holds (argv, rc, stdout, stderr) of the UNMODIFIED reference binary for the
fastq_info block of run_tests.sh:252-343, BASELINE config 1 (c18_10000 pair) and ~700 further invocations
(every fixture in every mode, hand-made edge files).  The oracle must reproduce every byte.
"""
import pytest

from _util import golden_transcripts, oracle_lib, oracle_run_files

CASES = golden_transcripts()


def test_corpus_size():
    assert len(CASES) >= 800
    rcs = {c["rc"] for c in CASES}
    assert rcs == {0, 1, 3}


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_matches_reference_transcript(idx):
    c = CASES[idx]
    rc, out, err = oracle_run_files(c["argv"])
    assert rc == c["rc"], (c["argv"], err)
    assert out == c["stdout"], c["argv"]
    assert err == c["stderr"], c["argv"]


def test_run_tests_sh_explicit_statistic():
    # run_tests.sh:302 — the only statistic the reference's suite pins explicitly
    rc, out, err = oracle_run_files(["inputs/test_1.fastq.gz"])
    assert rc == 0
    line = [l for l in err.splitlines() if l.startswith("Read length:")][0]
    assert line.split(":")[1] == " 90 90 90"


def test_baseline_config1_c18_pair_must_fail():
    # BASELINE.json configs[0]; run_tests.sh:347-348 expects failure
    rc, out, err = oracle_run_files(["inputs/c18_10000_1.fastq.gz", "inputs/c18_10000_2.fastq.gz"])
    assert rc == 3
    assert err.endswith("ERROR: Error in file inputs/c18_10000_2.fastq.gz: line 8: unpaired read - 97ZZTR1:325:C1UY6ACXX:1:1101:1549:1941\n")


def test_qual_range_ladder():
    lib = oracle_lib()
    f = lib.oracle_qual_range2enc
    assert f(35, 73) == b"33"
    assert f(38, 74) == b"33"
    assert f(66, 102) == b"64"
    assert f(35, 102) == b"sanger"
    assert f(61, 113) == b"solexa"
    assert f(59, 74) == b"33 *"
    assert f(34, 127) is None
    assert f(64, 126) is None  # max > min + 60 and not sanger

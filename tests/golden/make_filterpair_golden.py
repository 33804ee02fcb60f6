#!/usr/bin/env python
"""Golden transcripts of fastq_filterpair.

Run in the build container (needs oracle/_ref/fastq_filterpair, the UNMODIFIED reference compiled by `make -C oracle ref-tools` from
/root/reference/src/{hash,fastq,fastq_filterpair}.c):

    python tests/golden/make_filterpair_golden.py

Pairs of files under tests/golden/inputs/, a few hand-made pairs (tests/golden/pair_inputs/, written by this script) and one generated
pair of 25 000 records (not committed: the tests regenerate it) go through the tool in both modes (cwd=tests/golden); (argv, rc, stdout,
stderr, whether the three output files were created, and their INFLATED contents when the tool finished them) go to
tests/golden/filterpair_transcripts.json (latin-1; long texts as length + SHA-256)."""
import gzip
import hashlib
import json
import os
import random
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "fastq_filterpair")
OUTS = ["P1.gz", "P2.gz", "UP.gz"]


def big_pair():
    """25 000 pairs: file 2 permuted in windows of 64, some mates missing on either side, some names twice in file 2"""
    rng = random.Random(11)
    n = 25_000
    seq = lambda: "".join(rng.choice("ACGT") for _ in range(24))  # noqa: E731
    r1 = [f"@M{i:06d}/1\n{seq()}\n+\n{'I' * 24}\n" for i in range(n) if i % 97 != 5]
    idx = list(range(n))
    for w in range(0, n, 64):
        blk = idx[w:w + 64]
        rng.shuffle(blk)
        idx[w:w + 64] = blk
    r2 = []
    for i in idx:
        if i % 89 == 7:
            continue
        r2.append(f"@M{i:06d}/2\n{seq()}\n+\n{'H' * 24}\n")
        if i % 1000 == 3:
            r2.append(f"@M{i:06d}/2\n{seq()}\n+\n{'G' * 24}\n")
    return "".join(r1), "".join(r2)


def handmade():
    d = os.path.join(HERE, "pair_inputs")
    os.makedirs(d, exist_ok=True)
    rec = lambda name, s="ACGTACGT", q=None: f"@{name}\n{s}\n+\n{q or 'I' * len(s)}\n"  # noqa: E731
    files = {
        "p_1.fq": "".join(rec(f"r{i}/1") for i in range(8)),
        "p_2_same.fq": "".join(rec(f"r{i}/2", "TTGGCCAA") for i in range(8)),
        "p_2_perm.fq": "".join(rec(f"r{i}/2", "TTGGCCAA") for i in (3, 0, 1, 2, 7, 6, 5, 4)),
        "p_2_some.fq": "".join(rec(f"r{i}/2", "TTGGCCAA") for i in (1, 2, 9, 5, 5, 11)),
        "p_2_none.fq": "".join(rec(f"x{i}/2", "TTGGCCAA") for i in range(3)),
        "p_2_noat.fq": rec("r0/2") + "r1/2\nACGT\n+\nIIII\n" + rec("r2/2"),
        "p_2_noat_first.fq": "r0/2\nACGT\n+\nIIII\n" + rec("r1/2"),
        "p_2_trunc.fq": rec("r0/2") + rec("r4/2") + "@r5/2\nACGT\n",
        "p_2_invalid.fq": rec("r0/2", "ACXT") + rec("r1/2", "ACGT", "II") + rec("r6/2"),  # file 2 is never validated in the default mode
        "p_2_casava.fq": "".join(rec(f"r{i} 2:N:0:ACGT", "TTGGCCAA") for i in (0, 2, 4)),
        "p_1_casava.fq": "".join(rec(f"r{i} 1:N:0:ACGT") for i in range(5)),
        "p_1_dup.fq": rec("r0/1") + rec("r1/1") + rec("r0/1"),
        "p_1_bad.fq": rec("r0/1") + rec("r1/1", "ACXT") + rec("r2/1"),
        "p_2_nolf.fq": (rec("r1/2") + rec("r0/2"))[:-1],
        "p_2_nul.fq": rec("r1/2") + "@r2/2\nAC\x00GT\n+\nIIII\n" + rec("r0/2"),
        "p_empty.fq": "",
    }
    for name, text in files.items():
        with open(os.path.join(d, name), "wb") as fh:
            fh.write(text.encode("latin-1"))
    a, b = big_pair()
    for name, text in (("big_1.fq", a), ("big_2.fq", b)):
        with open(os.path.join(d, name), "wb") as fh:
            fh.write(text.encode("latin-1"))
    return sorted(files)


def compact(text, key, rec):
    if len(text) <= 4096:
        rec[key] = text.decode("latin-1")
    else:
        rec[key + "_len"], rec[key + "_sha256"] = len(text), hashlib.sha256(text).hexdigest()


def main():
    handmade()
    I = lambda f: "inputs/" + f  # noqa: E731
    P = lambda f: "pair_inputs/" + f  # noqa: E731
    pairs = [(I(a), I(b)) for a, b in [
        ("a_1.fastq.gz", "a_2.fastq.gz"), ("c18_10000_1.fastq.gz", "c18_10000_2.fastq.gz"), ("casava.1.8_1.fastq.gz", "casava.1.8_2.fastq.gz"),
        ("casava.1.8_readname_trunc_1.fastq.gz", "casava.1.8_readname_trunc_2.fastq.gz"), ("barcode_test_1.fastq.gz", "barcode_test_2.fastq.gz"),
        ("barcode_test2_1.fastq.gz", "barcode_test2_2.fastq.gz"), ("solexa_1.fastq.gz", "solexa_2.fastq.gz"), ("test_21_1.fastq.gz", "test_21_2.fastq.gz"),
        ("test_22_1.fastq.gz", "test_22_2.fastq.gz"), ("test_30_1.fastq.gz", "test_30_2.fastq.gz"), ("test_e19_1.fastq.gz", "test_e19_2.fastq.gz"),
        ("test_solid_1.fastq.gz", "test_solid_2.fastq.gz"), ("test_solid2_1.fastq.gz", "test_solid2_2.fastq.gz"),
        ("pbmc8k_S1_L007_R1_001.fastq.gz", "pbmc8k_S1_L007_R2_001.fastq.gz"), ("10xv1a_R1.fastq.gz", "10xv1a_R2.fastq.gz"), ("test_1.fastq.gz", "test_2.fastq.gz"),
        ("a_2.fastq.gz", "a_1.fastq.gz"), ("casava.1.8_2.fastq.gz", "casava.1.8_1.fastq.gz"), ("test_1.fastq.gz", "a_2.fastq.gz"), ("pe_bug14.fastq.gz", "pe_bug14.fastq.gz")]]
    e2 = [f for f in sorted(os.listdir(os.path.join(HERE, "inputs"))) if f.startswith("edge_pair_2")]
    pairs += [(I("edge_pair_1.fastq"), I(f)) for f in e2] + [(I(f), I("edge_pair_1.fastq")) for f in e2[:4]]
    singles = [f for i, f in enumerate(sorted(os.listdir(os.path.join(HERE, "inputs")))) if i % 3 == 0]
    pairs += [(I(f), I(f)) for f in singles]
    hm = ["p_2_same.fq", "p_2_perm.fq", "p_2_some.fq", "p_2_none.fq", "p_2_noat.fq", "p_2_noat_first.fq", "p_2_trunc.fq", "p_2_invalid.fq", "p_2_casava.fq", "p_2_nolf.fq",
          "p_2_nul.fq", "p_empty.fq"]
    pairs += [(P("p_1.fq"), P(f)) for f in hm] + [(P("p_1_casava.fq"), P("p_2_casava.fq")), (P("p_1_dup.fq"), P("p_2_same.fq")), (P("p_1_bad.fq"), P("p_2_same.fq")),
                                                    (P("p_empty.fq"), P("p_2_same.fq")), (P("p_empty.fq"), P("p_empty.fq")), (P("p_2_same.fq"), P("p_1.fq")),
                                                    (P("big_1.fq"), P("big_2.fq")), (P("big_2.fq"), P("big_1.fq"))]
    cases = []
    for a, b in pairs:
        cases.append([a, b] + OUTS)
        cases.append([a, b] + OUTS + ["sorted"])
    a, b = P("p_1.fq"), P("p_2_perm.fq")
    cases += [[], [a], [a, b, "P1.gz", "P2.gz"], [a, b] + OUTS + ["sorted", "x"], [a, b] + OUTS + ["unsorted"], ["inputs/nope.fq", b] + OUTS, [a, "inputs/nope.fq"] + OUTS,
              ["inputs/nope.fq", "inputs/nope2.fq"] + OUTS + ["sorted"]]
    out = []
    for argv in cases:
        for o in OUTS:
            if os.path.exists(os.path.join(HERE, o)):
                os.unlink(os.path.join(HERE, o))
        pr = subprocess.run([REF] + argv, cwd=HERE, capture_output=True)
        rec = {"argv": argv, "rc": pr.returncode, "stdout": pr.stdout.decode("latin-1"), "created": all(os.path.exists(os.path.join(HERE, o)) for o in OUTS)}
        compact(pr.stderr, "stderr", rec)
        if rec["created"]:
            try:
                datas = [gzip.open(os.path.join(HERE, o), "rb").read() for o in OUTS]
                for k, dta in enumerate(datas):
                    compact(dta, f"out{k}", rec)
            except (EOFError, OSError):
                pass  # the tool exited without closing its files
        out.append(rec)
    for o in OUTS:
        if os.path.exists(os.path.join(HERE, o)):
            os.unlink(os.path.join(HERE, o))
    with open(os.path.join(HERE, "filterpair_transcripts.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(len(out), "transcripts")


if __name__ == "__main__":
    main()

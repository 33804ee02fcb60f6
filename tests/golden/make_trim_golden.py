#!/usr/bin/env python
"""Golden transcripts of fastq_trim_poly_at.

Run in the build container (needs oracle/_ref/fastq_trim_poly_at, the UNMODIFIED reference compiled by `make -C oracle ref-tools` from
/root/reference/src/{hash,fastq,fastq_trim_poly_at}.c):

    python tests/golden/make_trim_golden.py

Every file under tests/golden/inputs/ and a set of hand-made poly-A / poly-T files (tests/golden/trim_inputs/, written by this script)
go through the tool with a few option sets (cwd=tests/golden); (argv, rc, stderr, stdout, and the INFLATED contents of the output file —
or their length and SHA-256 when longer than 4 KiB; null when the tool did not create or finish one) go to
tests/golden/trim_transcripts.json (latin-1)."""
import gzip
import hashlib
import json
import os
import random
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "fastq_trim_poly_at")


def poly_100k():
    rng = random.Random(6)
    body = lambda n: "".join(rng.choice("ACGT") for _ in range(n))  # noqa: E731
    return "".join(f"@r{i} x\n{(body(20) + 'A' * 12) if i % 3 == 0 else body(32)}\n+\n{'I' * 32}\n" for i in range(100_050))


def handmade():
    """records with poly-A tails, poly-T heads, both, N runs, lower case, CRLF, a last line without LF, short quality lines"""
    rng = random.Random(5)
    d = os.path.join(HERE, "trim_inputs")
    os.makedirs(d, exist_ok=True)
    files = {}

    def rec(i, seq, qual=None, eol="\n"):
        q = qual if qual is not None else "".join(rng.choice("FGHIJ#5") for _ in seq)
        return f"@r{i} x{eol}{seq}{eol}+{eol}{q}{eol}"
    body = lambda n: "".join(rng.choice("ACGT") for _ in range(n))  # noqa: E731
    recs = []
    for i in range(60):
        k = rng.choice([0, 3, 9, 10, 11, 25])
        kind = i % 6
        core = body(rng.choice([1, 5, 12, 40]))
        if kind == 0:
            s = core + "G" + "".join(rng.choice("AAAANan") for _ in range(k))
        elif kind == 1:
            s = "".join(rng.choice("TTTTNtn") for _ in range(k)) + "C" + core
        elif kind == 2:
            s = "T" * k + core + "A" * k
        elif kind == 3:
            s = rng.choice("ANan") * (k + 1)
        elif kind == 4:
            s = rng.choice("TNtn") * (k + 1)
        else:
            s = core
        recs.append(rec(i, s))
    files["poly_mixed.fq"] = "".join(recs)
    files["poly_crlf.fq"] = "".join(rec(i, body(8) + "A" * 14, eol="\r\n") for i in range(5))
    files["poly_nolf_end.fq"] = (rec(0, body(20)) + rec(1, body(6) + "A" * 15))[:-1]
    files["poly_nolf_allT.fq"] = rec(0, body(20)) + "@r1 x\n" + "T" * 30 + "\n+\n" + "I" * 30 + "\n@r2 x\n" + "T" * 12 + "\n+\n" + "I" * 12
    # quality lines that are shorter / longer than their sequence line: what the reference prints depends on its line buffers' history
    files["poly_shortqual.fq"] = (rec(0, body(50)) + rec(1, body(10) + "A" * 20, qual="I" * 12) + rec(2, "T" * 15 + body(20), qual="I" * 9)
                                  + rec(3, body(5) + "A" * 12, qual="I" * 40) + rec(4, body(30)) + rec(5, "T" * 11 + body(3), qual="#"))
    files["poly_tail_exact.fq"] = rec(0, body(4) + "A" * 10) + rec(1, body(4) + "A" * 9) + rec(2, "T" * 10 + body(4)) + rec(3, "T" * 9 + body(4))
    files["poly_trunc.fq"] = rec(0, body(4) + "A" * 12) + "@r1 x\n" + body(10) + "\n+\n"
    files["poly_nul.fq"] = rec(0, body(6) + "A" * 12) + "@r1 x\nACGT\x00AAAAAAAAAAAAAAA\n+\nIIII\x00IIIIIIIIIIIIIII\n" + rec(2, "T" * 13 + body(5))
    files["poly_empty.fq"] = ""
    files["poly_100k.fq"] = poly_100k()  # (not committed: tests regenerate it with the same function; it reaches the 100 000-record progress line)
    for name, text in files.items():
        with open(os.path.join(d, name), "wb") as fh:
            fh.write(text.encode("latin-1"))
    return ["trim_inputs/" + n for n in sorted(files)]


def main():
    out = []
    files = ["inputs/" + f for f in sorted(os.listdir(os.path.join(HERE, "inputs")))]
    mine = handmade()
    cases = []
    for i, p in enumerate(files):
        sets = ([], ["--min_poly_at_len", "3"], ["--min_poly_at_len=2", "--min_len", "0"], ["-a", "5", "-d", "30"])
        for o in sets[:4 if i % 5 == 0 else 1]:
            cases.append(o + ["--file", p, "--outfile", "OUT"])
    for p in mine:
        if p.endswith("poly_100k.fq"):
            cases.append(["--file", p, "--outfile", "OUT"])
            continue
        for o in ([], ["--min_poly_at_len", "3"], ["--min_poly_at_len", "1", "--min_len", "0"], ["--min_poly_at_len", "0"], ["--min_len", "-1"], ["--min_len", "25"],
                  ["--min_p", "12", "--min_l=2"]):
            cases.append(o + ["--file", p, "--outfile", "OUT"])
    p = "trim_inputs/poly_mixed.fq"
    cases += [[], ["--help"], ["--help", "--file", p], ["--file", p], ["--outfile", "OUT"], ["--file", "inputs/nope.fq", "--outfile", "OUT"], ["--fi", p, "--out", "OUT"],
              ["-b", p, "-c", "OUT"], ["--file", p, "--outfile", "OUT", "extra", "words"], ["--bogus", "--file", p, "--outfile", "OUT"], ["-x", "--file", p, "--outfile", "OUT"],
              ["--file", p, "--outfile", "OUT", "--min_len"], ["--file=" + p, "--outfile=OUT", "--min_poly_at_len=abc"], ["--", "--file", p, "--outfile", "OUT"],
              ["--file", p, "--file", mine[1], "--outfile", "OUT"], ["--m", "3", "--file", p, "--outfile", "OUT"], ["--min", "3", "--file", p, "--outfile", "OUT"]]
    for argv in cases:
        o = os.path.join(HERE, "OUT")
        if os.path.exists(o):
            os.unlink(o)
        pr = subprocess.run([REF] + argv, cwd=HERE, capture_output=True)
        rec = {"argv": argv, "rc": pr.returncode, "stderr": pr.stderr.decode("latin-1"), "stdout": pr.stdout.decode("latin-1"), "created": os.path.exists(o)}
        if os.path.exists(o) and pr.returncode == 0:
            data = gzip.open(o, "rb").read()
            if len(data) <= 4096:
                rec["outfile"] = data.decode("latin-1")
            else:
                rec["outfile_len"], rec["outfile_sha256"] = len(data), hashlib.sha256(data).hexdigest()
        out.append(rec)
        if os.path.exists(o):
            os.unlink(o)
    with open(os.path.join(HERE, "trim_transcripts.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(len(out), "transcripts")


if __name__ == "__main__":
    main()

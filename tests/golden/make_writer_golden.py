#!/usr/bin/env python
"""Golden transcripts of the record-WRITING reader tools: fastq_truncate and fastq_filter_n.

Run in the build container (needs oracle/_ref/fastq_truncate and oracle/_ref/fastq_filter_n, the UNMODIFIED reference compiled by
`make -C oracle ref-tools` from /root/reference/src/{hash,fastq,fastq_truncate,fastq_filter_n}.c):

    python tests/golden/make_writer_golden.py

Every file under tests/golden/inputs/ goes through both tools with a few parameter sets (cwd=tests/golden); (tool, argv, rc, stderr and
stdout — or its length and SHA-256 when it is longer than 4 KiB) go to tests/golden/writer_transcripts.json (latin-1)."""
import hashlib
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref")


def main():
    out = []
    files = sorted(os.listdir(os.path.join(HERE, "inputs")))
    cases = []
    for i, f in enumerate(files):
        p = "inputs/" + f
        for k in (["0"], ["1"], ["3"], ["100000"], ["-1"], ["abc"])[:6 if i % 4 == 0 else 3]:
            cases.append(("fastq_truncate", [p] + k))
        for o in ([], ["-n", "0"], ["-n", "10"], ["-n", "50"], ["-n", "200"])[:5 if i % 4 == 0 else 2]:
            cases.append(("fastq_filter_n", o + [p]))
    p = "inputs/" + files[0]
    cases += [("fastq_truncate", []), ("fastq_truncate", [p]), ("fastq_truncate", [p, "2", "x"]), ("fastq_truncate", ["inputs/nope.fq", "2"]),
              ("fastq_filter_n", []), ("fastq_filter_n", ["-n10", p]), ("fastq_filter_n", [p, "-n", "30"]), ("fastq_filter_n", ["-x", p]), ("fastq_filter_n", ["-n"]),
              ("fastq_filter_n", ["inputs/nope.fq"]), ("fastq_filter_n", ["-n", "5", "inputs/nope.fq"]), ("fastq_filter_n", [p, "extra"]), ("fastq_filter_n", [p, "a", "b"]),
              ("fastq_filter_n", ["--", p])]
    for tool, argv in cases:
        pr = subprocess.run([os.path.join(REF, tool)] + argv, cwd=HERE, capture_output=True)
        so = pr.stdout
        rec = {"tool": tool, "argv": argv, "rc": pr.returncode, "stderr": pr.stderr.decode("latin-1")}
        if len(so) <= 4096:
            rec["stdout"] = so.decode("latin-1")
        else:
            rec["stdout_len"], rec["stdout_sha256"] = len(so), hashlib.sha256(so).hexdigest()
        out.append(rec)
    with open(os.path.join(HERE, "writer_transcripts.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(len(out), "transcripts")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden transcripts of the reader-style tools for FQG_MODE_READER.

Run in the build container (needs oracle/_ref/fastq_num_reads and oracle/_ref/fastq_not_empty, the UNMODIFIED reference
compiled by `make -C oracle ref-tools` from /root/reference/src/{hash,fastq,fastq_num_reads,fastq_not_empty}.c):

    python tests/golden/make_reader_golden.py

Every file under tests/golden/inputs/ (the reference's own fixtures plus the hand-made edge cases of make_golden.py) is given
to both tools with cwd=tests/golden; (tool, argv, rc, stdout, stderr) go to tests/golden/reader_transcripts.json (latin-1)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref")


def main():
    out = []
    files = sorted(os.listdir(os.path.join(HERE, "inputs")))
    for tool in ("fastq_num_reads", "fastq_not_empty"):
        cases = [["inputs/" + f] for f in files] + [[], ["inputs/does_not_exist.fastq"], ["inputs/test_1.fastq.gz", "extra"]]
        for argv in cases:
            p = subprocess.run([os.path.join(REF, tool)] + argv, cwd=HERE, capture_output=True)
            out.append({"tool": tool, "argv": argv, "rc": p.returncode, "stdout": p.stdout.decode("latin-1"), "stderr": p.stderr.decode("latin-1")})
    with open(os.path.join(HERE, "reader_transcripts.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(len(out), "transcripts")


if __name__ == "__main__":
    main()

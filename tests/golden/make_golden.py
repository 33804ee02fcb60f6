#!/usr/bin/env python
"""Generate the golden vectors for the fastq_info hot path.

Run in the build container (needs oracle/_ref/fastq_info, i.e. the UNMODIFIED reference compiled by
oracle/Makefile from /root/reference/src/{hash,fastq,fastq_info}.c):

    python tests/golden/make_golden.py

It (1) writes the hand-made edge-case inputs under tests/golden/inputs/edge_* (the reference's own
fixtures tests/*.fastq.gz were copied there verbatim), (2) runs the reference on every invocation of
run_tests.sh:252-343 plus the extra modes/edge files listed below, with cwd=tests/golden, and
(3) stores (argv, rc, stdout, stderr) in tests/golden/transcripts.json (latin-1 text).
"""
import gzip
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "fastq_info")
INP = os.path.join(HERE, "inputs")


def rec(name, seq, qual=None, plus="+", nl="\n"):
    if qual is None:
        qual = "I" * len(seq)
    return f"@{name}{nl}{seq}{nl}{plus}{nl}{qual}{nl}"


def write(fname, data, gz=False):
    if isinstance(data, str):
        data = data.encode("latin-1")
    path = os.path.join(INP, fname)
    if gz:
        with open(path, "wb") as fh:
            fh.write(gzip.compress(data, 6, mtime=0))
    else:
        with open(path, "wb") as fh:
            fh.write(data)
    return "inputs/" + fname


def edge_files():
    """name -> relative path. Every file is tiny except the over-long-line ones (gz, highly compressible)."""
    f = {}
    S = "ACGTACGTAC"
    ok3 = "".join(rec(f"r{i}/1", S) for i in range(3))
    f["no_trailing_nl"] = write("edge_no_trailing_nl.fastq", ok3[:-1])
    f["trailing_blank"] = write("edge_trailing_blank.fastq", ok3 + "\n")
    f["tail2"] = write("edge_tail2.fastq", ok3 + "@r9/1\nACGT\n")
    f["tail3"] = write("edge_tail3.fastq", ok3 + "@r9/1\nACGT\n+\n")
    f["tail1_nonl"] = write("edge_tail1_nonl.fastq", ok3 + "@r9/1")
    f["blank_mid"] = write("edge_blank_mid.fastq", rec("r0/1", S) + "\n" + rec("r1/1", S))
    # header of exactly 999 bytes incl. newline (fits gzgets(…,1000)), and 1000 bytes (split)
    f["hdr999"] = write("edge_hdr999.fastq", rec("h" * 997, S) + rec("r1/1", S))
    f["hdr1000"] = write("edge_hdr1000.fastq", rec("h" * 998, S) + rec("r1/1", S))
    # split header whose remainder is a legal sequence: file is VALID for the reference after re-phasing
    f["hdr_split_valid"] = write("edge_hdr_split_valid.fastq", "@" + "h" * 998 + "ACG\n+\nIII\n" + "@" + "k" * 998 + "TTG\n+\nIII\n")
    f["plus1000"] = write("edge_plus1000.fastq", "@r0\nACGT\n+" + "p" * 999 + "\nIIII\n")
    # NUL handling
    f["nul_hdr_first"] = write("edge_nul_hdr_first.fastq", rec("r0/1", S) + "\0" + rec("r1/1", S)[1:] + rec("r2/1", S))
    f["nul_in_hdr"] = write("edge_nul_in_hdr.fastq", rec("r0/1", S) + "@r1\0junk/1\n" + S + "\n+\n" + "I" * 10 + "\n")
    f["nul_in_seq"] = write("edge_nul_in_seq.fastq", rec("r0/1", S) + "@r1/1\nACGT\0ACGTA\n+\nIIIIIIIIII\n")
    f["nul_in_seq_ok"] = write("edge_nul_in_seq_ok.fastq", rec("r0/1", S) + "@r1/1\nACGT\0ACGTA\n+\nIIII\n")
    f["nul_seq_first"] = write("edge_nul_seq_first.fastq", rec("r0/1", S) + "@r1/1\n\0CGT\n+\nIIII\n")
    f["nul_plus_first"] = write("edge_nul_plus_first.fastq", rec("r0/1", S) + "@r1/1\nACGT\n\0\nIIII\n")
    f["nul_qual_first"] = write("edge_nul_qual_first.fastq", rec("r0/1", S) + "@r1/1\nACGT\n+\n\0III\n" + rec("r2/1", S))
    f["nul_in_qual"] = write("edge_nul_in_qual.fastq", rec("r0/1", S) + "@r1/1\nACGT\n+\nII\0I\n")
    f["nul_in_qual_ok"] = write("edge_nul_in_qual_ok.fastq", rec("r0/1", S) + "@r1/1\nAC\n+\nII\0I\n" + rec("r2/1", S))
    f["nul_in_plus"] = write("edge_nul_in_plus.fastq", rec("r0/1", S) + "@r1/1\nACGT\n+\0zzz\nIIII\n")
    # CR handling
    f["cr_mid_seq"] = write("edge_cr_mid_seq.fastq", rec("r0/1", S) + "@r1/1\nACGT\rACGTA\n+\nIIII\n")
    f["cr_mid_seq_bad"] = write("edge_cr_mid_seq_bad.fastq", rec("r0/1", S) + "@r1/1\nACGT\rACGTA\n+\nIIIIIIIIII\n")
    f["crlf"] = write("edge_crlf.fastq", "".join(rec(f"r{i}/1", S, nl="\r\n") for i in range(3)))
    f["crlf_int"] = write("edge_crlf_int.fastq", "".join(rec(f"{i + 10}", S, nl="\r\n") for i in range(3)))
    f["cr_mid_qual"] = write("edge_cr_mid_qual.fastq", rec("r0/1", S) + "@r1/1\nACGT\n+\nII\rI\n")
    # colour-space flips
    f["cs_flip_TN"] = write("edge_cs_flip_TN.fastq", rec("r0/1", "TNTNTTNN", "IIIIIII") + rec("r1/1", "ACGT", "III"))
    f["cs_flip_allN"] = write("edge_cs_flip_allN.fastq", rec("r0/1", "NNNN", "IIII") + rec("r1/1", "ACGT", "IIII") + rec("r2/1", "ACGT", "II"))
    f["cs_backslash"] = write("edge_cs_backslash.fastq", rec("r0/1", "T01\\23", "IIIII"))
    f["cs_G"] = write("edge_cs_G.fastq", rec("r0/1", "G0123", "IIII") + rec("r1/1", "G01.3", "IIIII"))
    # alphabet
    f["highbit_base"] = write("edge_highbit_base.fastq", rec("r0/1", S) + b"@r1/1\nAC\xe9T\n+\nIIII\n".decode("latin-1"))
    f["lower_dot_digits"] = write("edge_lower_dot_digits.fastq", rec("r0/1", "acgtnN.0123ACGT"))
    f["ut_mix_tu"] = write("edge_ut_mix_tu.fastq", rec("r0/1", S) + rec("r1/1", "ACGTTTUAC"))
    f["ut_mix_ut"] = write("edge_ut_mix_ut.fastq", rec("r0/1", "ACGUUUA") + rec("r1/1", "ACGUUtAC"))
    f["ut_then_bad"] = write("edge_ut_then_bad.fastq", rec("r0/1", "ATUX"))
    f["bad_then_ut"] = write("edge_bad_then_ut.fastq", rec("r0/1", "AXTU"))
    f["u_only"] = write("edge_u_only.fastq", rec("r0/1", "ACGUUACGUU") + rec("r1/1", "ACGTTACGTT"))
    # quality
    f["qual_highbit"] = write("edge_qual_highbit.fastq", rec("r0/1", "ACGT", b"II\xe2I".decode("latin-1")))
    f["qual_all_highbit"] = write("edge_qual_all_highbit.fastq", rec("r0/1", "ACGT", b"\xe2\xe3\xe4\xe5".decode("latin-1")))
    f["qual_126_127"] = write("edge_qual_126_127.fastq", rec("r0/1", "ACGT", "~~\x7f~"))
    f["qual_low"] = write("edge_qual_low.fastq", rec("r0/1", "ACGT", "!\x0e\x01#"))
    f["qual_space_tab"] = write("edge_qual_space_tab.fastq", rec("r0/1", "ACGT", "I \tI"))
    f["qual_short"] = write("edge_qual_short.fastq", rec("r0/1", S) + rec("r1/1", S, "III"))
    f["qual_long"] = write("edge_qual_long.fastq", rec("r0/1", S) + rec("r1/1", S, "I" * 12))
    # header 2
    f["plus_name_ok"] = write("edge_plus_name_ok.fastq", rec("r0/1", S, plus="+r0/1") + rec("r1/1", S, plus="+r1/1"))
    f["plus_name_bad"] = write("edge_plus_name_bad.fastq", rec("r0/1", S, plus="+r0/1") + rec("r1/1", S, plus="+r9/1"))
    f["plus_name_prefix"] = write("edge_plus_name_prefix.fastq", rec("r0/1", S, plus="+r0/1") + rec("r10/1", S, plus="+r1/1"))
    f["plus_name_casava"] = write("edge_plus_name_casava.fastq", rec("M1:1:2 1:N:0:A", S, plus="+M1:1:2 1:N:0:A") + rec("M1:1:3 1:N:0:A", S, plus="+M1:1:3 2:Y:0:C"))
    f["plus_name_casava_bad"] = write("edge_plus_name_casava_bad.fastq", rec("M1:1:2 1:N:0:A", S, plus="+M1:1:2 1:N:0:A") + rec("M1:1:3 1:N:0:A", S, plus="+M1:1:4 1:N:0:A"))
    f["plus_cr"] = write("edge_plus_cr.fastq", rec("r0/1", S, plus="+\r") + rec("r1/1", S, plus="+\rxx"))
    f["plus_missing"] = write("edge_plus_missing.fastq", rec("r0/1", S) + rec("r1/1", S, plus="-"))
    f["plus_space"] = write("edge_plus_space.fastq", rec("r0/1", S) + rec("r1/1", S, plus="+ "))
    # headers / names
    f["at_only"] = write("edge_at_only.fastq", rec("r0/1", S) + rec("", S))
    f["at_cr"] = write("edge_at_cr.fastq", rec("r0/1", S) + "@\r\n" + S + "\n+\n" + "I" * 10 + "\n")
    f["no_at"] = write("edge_no_at.fastq", rec("r0/1", S) + "r1/1\n" + S + "\n+\n" + "I" * 10 + "\n")
    f["no_at_first"] = write("edge_no_at_first.fastq", "r1/1\n" + S + "\n+\n" + "I" * 10 + "\n")
    f["dup3"] = write("edge_dup3.fastq", "".join(rec(n, S) for n in ["a/1", "b/1", "c/1", "b/1", "a/1", "b/1"]))
    f["dup_after_bad"] = write("edge_dup_after_bad.fastq", rec("a/1", S) + rec("b/1", "ACXT") + rec("a/1", S))
    f["bad_after_dup"] = write("edge_bad_after_dup.fastq", rec("a/1", S) + rec("a/1", "ACXT") + rec("b/1", "ACXT"))
    f["dup_is_bad"] = write("edge_dup_is_bad.fastq", rec("a/1", S) + rec("b/1", S) + rec("a/1", "ACXT"))
    f["default_pe_names"] = write("edge_default_pe_names.fastq", rec("ab:1", S) + rec("ab:2", S))
    f["casava_slash"] = write("edge_casava_slash.fastq", rec("S4:4:1:9:1/1 1:Y:0:0", S) + rec("S4:4:1:9:1/2 2:Y:0:0", S))
    f["casava_then_plain"] = write("edge_casava_then_plain.fastq", rec("M1:1:2 1:N:0:A", S) + rec("plainname", S) + rec("plainname", S))
    f["casava_no_space_dup"] = write("edge_casava_nospace.fastq", rec("M1:1:2 1:N:0:A", S) + rec("M1:1:2", S))
    f["int_names"] = write("edge_int_names.fastq", "".join(rec(str(i), S) for i in range(1, 5)))
    f["int_then_dup"] = write("edge_int_then_dup.fastq", rec("1", S) + rec("2", S) + rec("1", S))
    f["nosuffix"] = write("edge_nosuffix.fastq", rec("read_one extra", S) + rec("read_two extra", S))
    f["suffix_variants"] = write("edge_suffix_variants.fastq", rec("x#0", S) + rec("x#1", S))
    f["tab_suffix"] = write("edge_tab_suffix.fastq", rec("x\tA", S) + rec("x\tB", S))
    f["short_name_pe"] = write("edge_short_name_pe.fastq", rec("a/1", S) + rec("b", S))
    # interleaved / paired
    il = "".join(rec(f"p{i}/1", S) + rec(f"p{i}/2", S) for i in range(3))
    f["il_ok"] = write("edge_il_ok.fastq", il)
    f["il_odd"] = write("edge_il_odd.fastq", il + rec("p9/1", S))
    f["il_mismatch"] = write("edge_il_mismatch.fastq", rec("p0/1", S) + rec("p0/2", S) + rec("p1/1", S) + rec("p2/2", S))
    f["il_bad_m2"] = write("edge_il_bad_m2.fastq", rec("p0/1", S) + rec("p0/2", "ACXT"))
    f["il_bad_m1_and_mismatch"] = write("edge_il_bad_m1_mm.fastq", rec("p0/1", "ACXT") + rec("p1/2", S))
    f["il_m2_noat"] = write("edge_il_m2_noat.fastq", rec("p0/1", S) + "p0/2\n" + S + "\n+\n" + "I" * 10 + "\n")
    f["il_trunc_m2"] = write("edge_il_trunc_m2.fastq", rec("p0/1", S) + "@p0/2\n" + S + "\n")
    p1 = "".join(rec(f"q{i}/1", S) for i in range(4))
    p2 = "".join(rec(f"q{i}/2", S) for i in range(4))
    f["pair_1"] = write("edge_pair_1.fastq", p1)
    f["pair_2"] = write("edge_pair_2.fastq", p2)
    f["pair_2_perm"] = write("edge_pair_2_perm.fastq", "".join(rec(f"q{i}/2", S) for i in [2, 0, 3, 1]))
    f["pair_2_short"] = write("edge_pair_2_short.fastq", "".join(rec(f"q{i}/2", S) for i in range(3)))
    f["pair_2_short2"] = write("edge_pair_2_short2.fastq", "".join(rec(f"q{i}/2", S) for i in range(2)))
    f["pair_2_long"] = write("edge_pair_2_long.fastq", p2 + rec("q7/2", S))
    f["pair_2_rep"] = write("edge_pair_2_rep.fastq", "".join(rec(f"q{i}/2", S) for i in [0, 1, 1, 2]))
    f["pair_2_bad"] = write("edge_pair_2_bad.fastq", "".join(rec(f"q{i}/2", S if i != 2 else "ACXT") for i in range(4)))
    f["pair_2_trunc"] = write("edge_pair_2_trunc.fastq", p2[:-14])
    f["pair_2_unp_bad"] = write("edge_pair_2_unp_bad.fastq", rec("q0/2", S) + rec("zz/2", "ACXT"))
    f["pair_2_casava"] = write("edge_pair_2_casava.fastq", "".join(rec(f"q{i}/1 2:N:0:A", S) for i in range(4)))
    f["pair_2_qual"] = write("edge_pair_2_qual.fastq", "".join(rec(f"q{i}/2", S, "#" * 10) for i in range(4)))
    f["pair_2_len"] = write("edge_pair_2_len.fastq", "".join(rec(f"q{i}/2", S * (i + 1)) for i in range(4)))
    # over-long sequence line (>= 2 500 000 incl. newline): split by gzgets
    big = "A" * 2499999
    f["seq_max_ok"] = write("edge_seq_max_ok.fastq.gz", "@r0/1\n" + "A" * 2499998 + "\n+\n" + "I" * 2499998 + "\n", gz=True)
    f["seq_too_long"] = write("edge_seq_too_long.fastq.gz", "@r0/1\n" + big + "\n+\n" + "I" * 2499999 + "\n", gz=True)
    # encodings
    f["enc_solexa"] = write("edge_enc_solexa.fastq", rec("r0/1", "ACGT", ";;hh"))
    f["enc_64"] = write("edge_enc_64.fastq", rec("r0/1", "ACGT", "@@hh"))
    f["enc_33star"] = write("edge_enc_33star.fastq", rec("r0/1", "ACGT", "JJJJ"))
    f["enc_range"] = write("edge_enc_range.fastq", rec("r0/1", "ACGT", "<<<}"))
    f["one_read"] = write("edge_one_read.fastq", rec("r0/1", S))
    f["lens"] = write("edge_lens.fastq", "".join(rec(f"r{i}/1", "A" * n) for i, n in enumerate([5, 9, 9, 30, 2, 7, 7, 7])))
    f["garbage"] = write("edge_garbage.fastq", "this is not a fastq file\nat all\n")
    f["only_nl"] = write("edge_only_nl.fastq", "\n\n\n\n\n")
    f["multi_member"] = os.path.join("inputs", "edge_multi_member.fastq.gz")
    with open(os.path.join(INP, "edge_multi_member.fastq.gz"), "wb") as fh:
        fh.write(gzip.compress(rec("r0/1", S).encode(), mtime=0) + gzip.compress(rec("r1/1", S).encode(), mtime=0))
    return f


def corpus_cases(E):
    T = "inputs/"
    c = []
    a = c.append
    # --- run_tests.sh:252-343, in order -------------------------------------------------------
    a([])
    for n in ["e1", "e2", "e3", "e4", "e5", "e6", "e7", "e8", "e9"]:
        a([T + f"test_{n}.fastq.gz"])
    a(["-r", T + "test_e9.fastq.gz"])
    for n in ["e10", "e20", "e21"]:
        a([T + f"test_{n}.fastq.gz"])
    a(["-q", T + "test_e20.fastq.gz"])
    a(["-q", T + "test_e21.fastq.gz"])
    a([T + "test_33.fastq.gz"])
    a(["-q", T + "test_33.fastq.gz"])
    for n in ["e13", "e14", "e15", "e16"]:
        a([T + f"test_{n}.fastq.gz"])
    for n in ["e10", "e13", "e14", "e15", "e16"]:
        a(["-r", T + f"test_{n}.fastq.gz"])
    a([T + "test_e17.fastq.gz"])
    a([T + "test_e19_1.fastq.gz", T + "test_e19_2.fastq.gz"])
    a([T + "test_e19_2.fastq.gz", T + "test_e19_1.fastq.gz"])
    a([T + "test_e19_1.fastq.gz", T + "test_empty.fastq.gz"])
    a([T + "test_empty.fastq.gz", T + "test_e19_1.fastq.gz"])
    a(["-r", "-s", T + "test_e19_1.fastq.gz", T + "test_e19_2.fastq.gz"])
    a(["-r", "-s", T + "test_e19_2.fastq.gz", T + "test_e19_1.fastq.gz"])
    a(["-f", T + "test_dot.fastq.gz"])
    a([T + "test_empty.fastq.gz"])
    a(["-r", T + "test_empty.fastq.gz"])
    a(["-s", "-r", T + "test_empty.fastq.gz", T + "test_1.fastq.gz"])
    a(["-s", "-r", T + "test_1.fastq.gz", T + "test_empty.fastq.gz"])
    a(["-h"])
    a([T + "test_dot.fastq.gz"])
    a(["-e", T + "test_dot.fastq.gz"])
    a([T + "edge_empty.fastq"])
    a(["-e", T + "edge_empty.fastq"])
    a([T + "test_1.fastq.gz"])
    a([T + "test_30_1.fastq.gz", T + "test_30_2.fastq.gz"])
    for n in ["test_2", "test_13", "test_17", "test_pacbio", "test_ont", "test_ont2", "test_pacbio2", "test_21_1"]:
        a([T + n + ".fastq.gz"])
    a([T + "test_21_1.fastq.gz", T + "test_21_2.fastq.gz"])
    a(["-r", "-s", T + "test_21_1.fastq.gz", T + "test_21_2.fastq.gz"])
    a([T + "pe_bug14.fastq.gz", T + "pe_bug14.fastq.gz"])
    for n in ["nanopore_rna1", "nanopore_rna2", "nanopore_rna3", "nanopore_rna4", "nanopore_rna5"]:
        a([T + n + ".fastq.gz"])
    a([T + "casava.1.8i.fastq.gz", "pe"])
    a([T + "test_solid_1.fastq.gz", T + "test_solid_2.fastq.gz"])
    a([T + "test_solid2_1.fastq.gz", T + "test_solid2_2.fastq.gz"])
    a([T + "solexa_1.fastq.gz", T + "solexa_2.fastq.gz"])
    a([T + "casava.1.8_readname_trunc_1.err.fastq.gz", T + "casava.1.8_readname_trunc_2.fastq.gz"])
    a([T + "casava.1.8_readname_trunc_2.fastq.gz", T + "casava.1.8_readname_trunc_1.err.fastq.gz"])
    a([T + "casava.1.8_readname_trunc_1.err2.fastq.gz", T + "casava.1.8_readname_trunc_2.fastq.gz"])
    a([T + "casava.1.8_readname_trunc_1.err.fastq.gz"])
    a(["-s", T + "casava.1.8_readname_trunc_1.fastq.gz", T + "casava.1.8_readname_trunc_2.fastq.gz"])
    a([T + "casava.1.8_readname_trunc_1.fastq.gz", T + "casava.1.8_2.fastq.gz"])
    a(["-r", "-s", T + "casava.1.8_readname_trunc_1.fastq.gz", T + "casava.1.8_2.fastq.gz"])
    a(["--help"])
    # --- BASELINE config 1 and friends ------------------------------------------------------------
    a([T + "c18_10000_1.fastq.gz", T + "c18_10000_2.fastq.gz"])
    a([T + "c18_10000_1.fastq.gz"])
    a([T + "c18_10000_2.fastq.gz"])
    a(["-r", T + "c18_10000_1.fastq.gz"])
    a(["-r", "-s", T + "c18_10000_1.fastq.gz", T + "c18_10000_2.fastq.gz"])
    a([T + "c18_10000_1.fastq.gz", T + "c18_10000_1.fastq.gz"])
    a([T + "c18_10000_1.fastq.gz", "pe"])
    # --- every reference fastq fixture in every single-file mode ------------------------------------
    fixtures = sorted(x for x in os.listdir(INP) if x.endswith(".fastq.gz") and not x.startswith("edge_"))
    for fx in fixtures:
        a([T + fx])
        a(["-r", T + fx])
        a([T + fx, "pe"])
    # pairs of related fixtures in both pair modes
    pairs = [("a_1", "a_2"), ("a_1", "test_2"), ("test_1", "test_e9"), ("test_1", "test_2"), ("test_2", "test_1"),
             ("casava.1.8_1", "casava.1.8_2"), ("test_22_1", "test_22_2"), ("test_solid_2", "test_solid_1"),
             ("solexa_2", "solexa_1"), ("barcode_test_1", "barcode_test_2"), ("tx.I1", "tx.I2"),
             ("10xv1a_R1", "10xv1a_R2"), ("pbmc8k_S1_L007_R1_001", "pbmc8k_S1_L007_R2_001"),
             ("test_21_2", "test_21_1"), ("test_30_2", "test_30_1"), ("test_e19_1", "test_e19_1")]
    for x, y in pairs:
        a([T + x + ".fastq.gz", T + y + ".fastq.gz"])
        a(["-r", "-s", T + x + ".fastq.gz", T + y + ".fastq.gz"])
        a(["-s", T + x + ".fastq.gz", T + y + ".fastq.gz"])
        a(["-r", T + x + ".fastq.gz", T + y + ".fastq.gz"])
    # --- option handling ------------------------------------------------------------------------
    a(["-q", "-e", "-r", T + "test_1.fastq.gz"])
    a([T + "test_1.fastq.gz", "-r"])
    a(["-x", T + "test_1.fastq.gz"])
    a([T + "test_1.fastq.gz", T + "test_2.fastq.gz", T + "test_2.fastq.gz"])
    a(["-e", T + "test_empty.fastq.gz"])
    a(["-e", "-r", T + "test_empty.fastq.gz"])
    a(["-e", T + "test_empty.fastq.gz", T + "test_1.fastq.gz"])
    a(["-e", T + "test_empty.fastq.gz", "pe"])
    a([T + "test_1.fastq.gz", "pex"])
    a([T + "test_1.fastq.gz", "p"])
    # --- hand-made edge files: default, -r and pe on each -------------------------------------------
    skip_pair = {"pair_2", "pair_2_perm", "pair_2_short", "pair_2_short2", "pair_2_long", "pair_2_rep", "pair_2_bad",
                 "pair_2_trunc", "pair_2_unp_bad", "pair_2_casava", "pair_2_qual", "pair_2_len"}
    for k in sorted(E):
        a([E[k]])
        a(["-r", E[k]])
        if k not in skip_pair:
            a([E[k], "pe"])
    for k in sorted(skip_pair):
        a([E["pair_1"], E[k]])
        a(["-r", "-s", E["pair_1"], E[k]])
        a([E[k], E["pair_1"]])
        a(["-r", "-s", E[k], E["pair_1"]])
    for k in ["crlf", "cs_flip_TN", "nul_hdr_first", "hdr1000", "qual_highbit", "tail2", "dup3", "lens", "garbage", "trailing_blank", "no_trailing_nl"]:
        a([E["pair_1"], E[k]])
        a([E[k], E["pair_1"]])
        a(["-r", "-s", E["pair_1"], E[k]])
        a(["-r", "-s", E[k], E["pair_1"]])
    a(["-q", E["qual_highbit"]])
    a(["-q", E["qual_all_highbit"]])
    a(["-q", "-r", E["qual_all_highbit"]])
    a(["-q", E["enc_range"]])
    return c


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/fastq_info missing: run `make -C oracle ref` first")
    open(os.path.join(INP, "edge_empty.fastq"), "wb").close()
    E = edge_files()
    out = []
    for argv in corpus_cases(E):
        p = subprocess.run([REF] + argv, cwd=HERE, capture_output=True)
        out.append({"argv": argv, "rc": p.returncode, "stdout": p.stdout.decode("latin-1"), "stderr": p.stderr.decode("latin-1")})
    with open(os.path.join(HERE, "transcripts.json"), "w") as fh:
        json.dump(out, fh, indent=0, ensure_ascii=True)
    rcs = {}
    for o in out:
        rcs[o["rc"]] = rcs.get(o["rc"], 0) + 1
    print(f"{len(out)} transcripts; exit-status histogram {rcs}")


if __name__ == "__main__":
    main()

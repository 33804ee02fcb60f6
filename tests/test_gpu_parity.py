"""Parity of the CUDA path (through the C ABI of libfastq_gpu.so) with the reference: golden transcripts of the
unmodified reference binary, differential fuzz against the CPU oracle, the CLI, and synthetic Illumina / long-read
streams generated on the device (checked against the oracle at sizes it finishes in seconds, and through
size-independent properties at larger sizes)."""
import os
import random
import subprocess

import pytest

from _util import GOLDEN, ROOT, fqg_run, fqg_run_files, golden_transcripts, oracle_run
from test_oracle_fuzz import NAMES, make_file, mutate, render

pytestmark = pytest.mark.gpu
CASES = golden_transcripts()


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_gpu_matches_reference_transcript(idx):
    c = CASES[idx]
    got = fqg_run_files(c["argv"], chunk=0, kind="gpu")
    assert got == (c["rc"], c["stdout"], c["stderr"]), c["argv"]


@pytest.mark.parametrize("idx", range(0, len(CASES), 3))
def test_gpu_matches_reference_transcript_chunked(idx):
    c = CASES[idx]
    chunk = [64, 1000, 4096][idx % 3]
    got = fqg_run_files(c["argv"], chunk=chunk, kind="gpu")
    assert got == (c["rc"], c["stdout"], c["stderr"]), (c["argv"], chunk)


@pytest.mark.parametrize("seed", range(300))
def test_gpu_fuzz_against_oracle(seed):
    rng = random.Random(20_000 + seed)
    style = rng.randrange(len(NAMES))
    n = rng.choice([0, 1, 2, 3, 5, 8, 13, 40, 300])
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
    chunk = rng.choice([0, 0, 1, 5, 33, 200, 5000])
    assert fqg_run(argv, d1, d2, chunk=chunk, kind="gpu") == oracle_run(argv, d1, d2), (argv, chunk, d1, d2)


CLI = os.path.join(ROOT, "fastq_utils_b200", "fastq_info_gpu")


@pytest.mark.parametrize("idx", [i for i, c in enumerate(CASES) if any(k in " ".join(c["argv"]) for k in ("c18_10000", "test_e1.", "test_e9", "test_21", "solid", "test_empty"))][:40])
def test_cli_matches_reference_transcript(idx):
    c = CASES[idx]
    p = subprocess.run([CLI] + c["argv"], cwd=GOLDEN, capture_output=True)
    assert (p.returncode, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")) == (c["rc"], c["stdout"], c["stderr"]), c["argv"]


# ------------------------------------------------------------------------------------------- synthetic streams
def _illumina(n, mate=1, perm=0, first=0, seed=42):
    import torch
    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes()
    t = torch.zeros(n * rb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_illumina(t, first, n, seed=seed, mate=mate, perm_window=perm, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return t, n * rb


def test_synth_illumina_single_matches_oracle():
    t, nb = _illumina(60_000)
    data = bytes(t[:nb].cpu().numpy())
    for argv in (["a.fq"], ["-r", "a.fq"], ["a.fq", "pe"]):
        assert fqg_run(argv, data, None, chunk=0, kind="gpu") == oracle_run(argv, data, None), argv
    assert fqg_run(["a.fq"], data, None, chunk=1 << 20, kind="gpu") == oracle_run(["a.fq"], data, None)


def test_synth_illumina_pair_matches_oracle():
    t1, nb = _illumina(40_960, mate=1)
    t2, _ = _illumina(40_960, mate=2, perm=1024)
    d1, d2 = bytes(t1[:nb].cpu().numpy()), bytes(t2[:nb].cpu().numpy())
    for argv in (["a.fq", "b.fq"], ["-r", "-s", "a.fq", "b.fq"], ["-s", "a.fq", "b.fq"]):
        got = fqg_run(argv, d1, d2, chunk=0, kind="gpu")
        assert got == oracle_run(argv, d1, d2), argv
    assert fqg_run(["a.fq", "b.fq"], d1, d2, chunk=0, kind="gpu")[0] == 0
    assert "Readnames do not match" in fqg_run(["-r", "-s", "a.fq", "b.fq"], d1, d2, chunk=0, kind="gpu")[2]


def test_synth_negative_twins_match_oracle():
    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes()
    t, nb = _illumina(30_000)
    base = bytearray(t[:nb].cpu().numpy().tobytes())
    # duplicate of record 100 injected at 20000; an invalid base at record 25000 (later: the duplicate wins)
    dup = bytearray(base); dup[20_000 * rb:20_001 * rb] = base[100 * rb:101 * rb]; dup[25_000 * rb + 70] = ord("X")
    # invalid base earlier than the duplicate: it wins
    bad = bytearray(dup); bad[5_000 * rb + 60] = ord("*")
    # quality one byte short / a '+' line replaced / truncated last record
    short = bytearray(base); del short[7_000 * rb + rb - 2]
    plus = bytearray(base); plus[9_000 * rb + 55 + 151] = ord("-")
    trunc = base[:len(base) - 100]
    for data in (dup, bad, short, plus, trunc):
        for argv in (["a.fq"], ["-r", "a.fq"]):
            assert fqg_run(argv, bytes(data), None, chunk=0, kind="gpu") == oracle_run(argv, bytes(data), None), argv
    # mates: one removed from file 2 / one extra name in file 1
    t2, _ = _illumina(30_000, mate=2, perm=1024)
    m2 = bytearray(t2[:nb].cpu().numpy().tobytes())
    removed = m2[:123 * rb] + m2[124 * rb:]
    for f1, f2 in ((bytes(base), bytes(removed)), (bytes(base), bytes(m2[:20_000 * rb])), (bytes(base[:29_000 * rb]), bytes(m2))):
        assert fqg_run(["a.fq", "b.fq"], f1, f2, chunk=0, kind="gpu") == oracle_run(["a.fq", "b.fq"], f1, f2)


def test_large_device_resident_properties():
    """4 M records (1.4 GB) fed straight from device memory in one call: counts, ranges and verdict are known by construction."""
    import torch
    import fastq_utils_b200 as fq
    n = 4_000_000
    t, nb = _illumina(n)
    h = fq.FastqInfo(fq.MODE_INDEX, index_capacity_hint=n)
    h.feed_device(0, t.data_ptr(), nb, last=True)
    rep = h.finish()
    assert rep.error.code == 0
    assert rep.n_index_entries == n and rep.file[0].n_records == n and rep.file[0].num_rds == 2 * n
    assert (rep.file[0].min_rl, rep.file[0].max_rl, rep.median_rl) == (151, 151, 151)
    assert (rep.file[0].min_qual, rep.file[0].max_qual) == (35, 73)
    rc, out, err = h.render(rep, "x.fq")
    assert rc == 0 and err.endswith("Number of reads: 4000000\nQuality encoding range: 35 73\nQuality encoding: 33\nRead length: 150 150 150\nOK\n")
    # same bytes through -r: identical statistics, no index
    h2 = fq.FastqInfo(fq.MODE_SINGLE)
    h2.feed_device(0, t.data_ptr(), nb, last=True)
    r2 = h2.finish()
    assert r2.error.code == 0 and r2.file[0].num_rds == n and r2.median_rl == 151
    # idempotence: reset and run again on the same context
    h.reset()
    h.feed_device(0, t.data_ptr(), nb, last=True)
    rep3 = h.finish()
    assert rep3.error.code == 0 and rep3.n_index_entries == n
    # a duplicate far apart is found at the later record
    rb = fq.illumina_record_bytes()
    t[3_999_000 * rb:3_999_001 * rb] = t[17 * rb:18 * rb].clone()
    h.reset()
    h.feed_device(0, t.data_ptr(), nb, last=True)
    rep4 = h.finish()
    assert rep4.error.code == 13 and rep4.error.record == 3_999_000 and rep4.error.line == 4 * 3_999_001
    h.close(); h2.close()
    del t
    torch.cuda.empty_cache()


def test_longreads_match_oracle():
    import torch
    import fastq_utils_b200 as fq
    g = torch.Generator().manual_seed(7)
    lens = torch.exp(torch.randn(300, generator=g) + 9.0).clamp(1000, 100000).to(torch.int64)
    hdr = fq.lib().fqg_synth_long_header_bytes()
    rec = hdr + 2 * lens + 4
    off = torch.zeros(301, dtype=torch.int64)
    off[1:] = torch.cumsum(rec, 0)
    nb = int(off[-1])
    t = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_longreads(t, off.cuda(), 0, 300, seed=7, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    data = bytes(t[:nb].cpu().numpy())
    for argv in (["-r", "a.fq"], ["a.fq"]):
        assert fqg_run(argv, data, None, chunk=0, kind="gpu") == oracle_run(argv, data, None), argv
    assert fqg_run(["-r", "a.fq"], data, None, chunk=1 << 20, kind="gpu") == oracle_run(["-r", "a.fq"], data, None)


def test_sharded_path_world1_matches_oracle():
    """The multi-GPU orchestration with a single rank: prescan, stream start, name packing, shard insert, merged render."""
    import torch
    import fastq_utils_b200 as fq
    from fastq_utils_b200 import dist as fqdist
    rb = fq.illumina_record_bytes()
    t, nb = _illumina(50_000)
    for mode, argv in ((fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])):
        run = fqdist.ShardedFastqInfo(mode, device=0, n_hint=50_000)
        res = run.run_device(t.data_ptr(), nb, name="a.fq")
        data = bytes(t[:nb].cpu().numpy())
        assert res["transcript"] == oracle_run(argv, data, None)
    t[40_000 * rb:40_001 * rb] = t[5 * rb:6 * rb].clone()
    t[45_000 * rb + 70] = ord("X")
    torch.cuda.synchronize()
    data = bytes(t[:nb].cpu().numpy())
    run = fqdist.ShardedFastqInfo(fq.MODE_INDEX, device=0)
    assert run.run_device(t.data_ptr(), nb, name="a.fq")["transcript"] == oracle_run(["a.fq"], data, None)


# ------------------------------------------------------------------------------------------- the fused single-pass kernel on small inputs
@pytest.mark.parametrize("idx", range(0, len(CASES), 2))
def test_gpu_fused_pass_on_corpus(idx, monkeypatch):
    """FQG_FUSED_MIN_BYTES=1 sends even tiny chunks through the fused scan+validate kernel (with its fallbacks)."""
    monkeypatch.setenv("FQG_FUSED_MIN_BYTES", "1")
    c = CASES[idx]
    chunk = [0, 0, 4096, 100][idx % 4]
    got = fqg_run_files(c["argv"], chunk=chunk, kind="gpu")
    assert got == (c["rc"], c["stdout"], c["stderr"]), (c["argv"], chunk)


@pytest.mark.parametrize("seed", range(200))
def test_gpu_fused_pass_fuzz(seed, monkeypatch):
    monkeypatch.setenv("FQG_FUSED_MIN_BYTES", "1")
    rng = random.Random(30_000 + seed)
    style = rng.randrange(len(NAMES))
    n = rng.choice([1, 2, 5, 13, 40, 300, 1500])
    mode = rng.choice(["single", "single_r", "pe", "pair", "pair_rs"])
    r1 = make_file(rng, n, style, 1, seqlen=(1, 300))
    two = mode.startswith("pair")
    r2 = make_file(rng, n, style, 2, seqlen=(1, 300)) if two else None
    if mode == "pe":
        inter = []
        for a, b in zip(r1, make_file(rng, n, style, 2)):
            inter += [a, b]
        r1 = inter
    for _ in range(rng.choice([0, 0, 1, 2])):
        mutate(rng, r1 if (not two or rng.random() < 0.5) else r2)
    d1 = render(rng, r1, "lf")
    d2 = render(rng, r2, "lf") if two else None
    argv = {"single": [], "single_r": ["-r"], "pe": [], "pair": [], "pair_rs": ["-r", "-s"]}[mode]
    argv = argv + ["a.fq"] + (["b.fq"] if two else []) + (["pe"] if mode == "pe" else [])
    chunk = rng.choice([0, 0, 20000, 70000])
    assert fqg_run(argv, d1, d2, chunk=chunk, kind="gpu") == oracle_run(argv, d1, d2), (argv, chunk)


# ------------------------------------------------------------------------------------------- the clean-data pass (fq_lanes_kernel)
@pytest.mark.parametrize("seed", range(0, 200, 4))
def test_gpu_per_record_fused_kernel_without_lanes(seed, monkeypatch):
    """FQG_NO_LANES=1 keeps the per-record fused kernel (the clean-data pass's fallback) covered on its own."""
    monkeypatch.setenv("FQG_NO_LANES", "1")
    test_gpu_fused_pass_fuzz(seed, monkeypatch)


def _run_ctx(mode, pieces, hint=0):
    """Feed device-resident pieces [(ptr, n)] of one file through a fresh context → (report, render transcript, path counts)."""
    import fastq_utils_b200 as fq
    h = fq.FastqInfo(mode, index_capacity_hint=hint)
    for i, (ptr, n) in enumerate(pieces):
        h.feed_device(0, ptr, n, last=i == len(pieces) - 1)
    rep = h.finish()
    tr = h.render(rep, "a.fq")
    pc = h.path_counts()
    h.close()
    return rep, tr, pc


@pytest.fixture(params=["per-line", "chunk-parallel"])
def lanes_mode(request, monkeypatch):
    """The clean-data pass has two modes (one thread per line for short lines, chunk-parallel for any length): FQG_NO_LINES=1
    forces the second one on short reads too."""
    if request.param == "chunk-parallel":
        monkeypatch.setenv("FQG_NO_LINES", "1")
    return request.param


@pytest.mark.parametrize("cuts", [(), (1 << 20,), (3_000_017, 9_000_001), (5 * 359 * 1000,), (359 * 4096 + 47, 359 * 8192 + 48, 359 * 20000 + 200)])
def test_lanes_pass_clean_illumina(cuts, lanes_mode):
    """Clean synthetic reads in one or several device-resident pieces cut at arbitrary bytes: the clean-data pass must accept
    every piece, and the transcript must be the oracle's."""
    import fastq_utils_b200 as fq
    n = 60_000
    t, nb = _illumina(n)
    data = bytes(t[:nb].cpu().numpy())
    edges = [0] + [c for c in cuts if c < nb] + [nb]
    pieces = [(t.data_ptr() + a, b - a) for a, b in zip(edges[:-1], edges[1:])]
    for mode, argv in ((fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])):
        rep, tr, pc = _run_ctx(mode, pieces, hint=n)
        assert tr == oracle_run(argv, data, None), (argv, cuts)
        assert pc["lanes"] == len(pieces) and pc["lanes_handed_on"] == 0 and pc["two_pass_fallbacks"] == 0, pc


def test_lanes_pass_longreads():
    """Records far longer than a tile: no record-length limit in the clean-data pass."""
    import torch
    import fastq_utils_b200 as fq
    g = torch.Generator().manual_seed(11)
    lens = torch.exp(torch.randn(400, generator=g) + 9.0).clamp(1000, 100000).to(torch.int64)
    hdr = fq.lib().fqg_synth_long_header_bytes()
    off = torch.zeros(401, dtype=torch.int64)
    off[1:] = torch.cumsum(hdr + 2 * lens + 4, 0)
    nb = int(off[-1])
    t = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_longreads(t, off.cuda(), 0, 400, seed=11, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    data = bytes(t[:nb].cpu().numpy())
    for mode, argv in ((fq.MODE_SINGLE, ["-r", "a.fq"]), (fq.MODE_INDEX, ["a.fq"])):
        rep, tr, pc = _run_ctx(mode, [(t.data_ptr(), nb)], hint=400)
        assert tr == oracle_run(argv, data, None), argv
        assert pc["lanes"] == 1 and pc["lanes_handed_on"] == 0, pc
    cut = nb // 3 + 5
    rep, tr, pc = _run_ctx(fq.MODE_SINGLE, [(t.data_ptr(), cut), (t.data_ptr() + cut, nb - cut)])
    assert tr == oracle_run(["-r", "a.fq"], data, None) and pc["lanes"] == 2, pc


def test_lanes_pass_hands_anomalies_on(lanes_mode):
    """Every kind of deviation from the clean shape makes the pass hand the chunk to the per-record kernels; the verdict and
    the statistics stay the oracle's (some of these inputs are still valid FASTQ)."""
    import torch
    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes()
    n = 20_000
    t, nb = _illumina(n)
    base = bytearray(t[:nb].cpu().numpy().tobytes())
    hl = base.index(b"\n") + 1  # header length of record 0 (all synthetic headers of this range have it)
    variants = {}
    v = bytearray(base); v[7000 * rb + hl + 20] = ord("X"); variants["bad base"] = v
    v = bytearray(base); v[7000 * rb + hl + 20] = ord("a"); variants["lower case (valid)"] = v
    v = bytearray(base); v[7000 * rb + hl + 20] = ord("."); variants["dot (valid, careful alphabet)"] = v
    v = bytearray(base); v[7000 * rb + hl + 151 + 2 + 30] = 0x0C; variants["control byte quality (valid)"] = v
    v = bytearray(base); v[7000 * rb + hl + 151 + 2 + 30] = 0xC8; variants["high quality byte"] = v
    v = bytearray(base); v[7000 * rb + hl + 151] = ord("-"); variants["plus line"] = v
    v = bytearray(base); v[7000 * rb] = ord(">"); variants["header marker"] = v
    v = bytearray(base); v[7000 * rb + 5] = 0; variants["NUL in header"] = v
    v = bytearray(base); del v[7000 * rb + rb - 2]; variants["short quality"] = v
    v = bytearray(base); del v[7000 * rb + hl + 10]; variants["short sequence"] = v
    v = bytearray(base); v[7000 * rb + hl:7000 * rb + hl + 150] = b""; v[7000 * rb + hl + 3:7000 * rb + hl + 3 + 150] = b""; variants["empty read"] = v
    v = bytearray(base.replace(b"\n", b"\r\n")); variants["crlf"] = v
    v = bytearray(base[:-1]); variants["no final newline (valid)"] = v
    v = bytearray(base[:-100]); variants["truncated"] = v
    v = bytearray(base); v[7000 * rb + 1:7000 * rb + 1] = b"Q" * 1200; variants["over-long header"] = v
    for name, v in variants.items():
        d = bytes(v)
        tt = torch.frombuffer(bytearray(d + b"\0" * 64), dtype=torch.uint8).cuda()
        for mode, argv in ((fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])):
            rep, tr, pc = _run_ctx(mode, [(tt.data_ptr(), len(d))], hint=n)
            assert tr == oracle_run(argv, d, None), (name, argv)
            if name not in ("no final newline (valid)", "lower case (valid)", "high quality byte"):
                assert pc["lanes_handed_on"] == 1, (name, pc)
            if name == "lower case (valid)":  # the chunk-parallel mode takes acgtn itself, the per-line mode hands them on
                assert pc["lanes"] + pc["lanes_handed_on"] == 1, (name, pc)
        del tt
    # accepted although unusual: lower case bases, a final line without LF, bytes above 0x7F as qualities (the host maps them)
    for name in ("no final newline (valid)", "high quality byte"):
        d = bytes(variants[name])
        tt = torch.frombuffer(bytearray(d + b"\0" * 64), dtype=torch.uint8).cuda()
        rep, tr, pc = _run_ctx(fq.MODE_INDEX, [(tt.data_ptr(), len(d))], hint=n)
        assert pc["lanes"] == 1 and pc["lanes_handed_on"] == 0, (name, pc)


def test_lanes_pass_statistics_roll_back(lanes_mode):
    """A chunk rejected by the record rules (lengths) after the pass itself found nothing must leave no trace in the
    statistics: same report as with the clean-data pass switched off."""
    import torch
    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes()
    t, nb = _illumina(30_000)
    base = bytearray(t[:nb].cpu().numpy().tobytes())
    hl = base.index(b"\n") + 1
    # record 12000: sequence and quality both one base shorter (valid, changes min length); record 25000: quality one short (error)
    v = bytearray(base)
    del v[25_000 * rb + rb - 2]
    del v[12_000 * rb + hl + 151 + 2 + 5]; del v[12_000 * rb + hl + 5]
    d = bytes(v)
    tt = torch.frombuffer(bytearray(d + b"\0" * 64), dtype=torch.uint8).cuda()
    cut = 20_000 * rb + 11
    rep, tr, pc = _run_ctx(fq.MODE_SINGLE, [(tt.data_ptr(), cut), (tt.data_ptr() + cut, len(d) - cut)])
    assert tr == oracle_run(["-r", "a.fq"], d, None)
    assert pc["lanes"] == 1 and pc["lanes_handed_on"] == 1, pc
    ok = d[:24_000 * rb - 2]  # everything before the broken record is valid: statistics include the shorter record 12000
    tt2 = torch.frombuffer(bytearray(ok + b"\0" * 64), dtype=torch.uint8).cuda()
    rep, tr, pc = _run_ctx(fq.MODE_SINGLE, [(tt2.data_ptr(), len(ok))])
    assert tr == oracle_run(["-r", "a.fq"], ok, None) and pc["lanes"] == 1 and rep.file[0].min_rl == 150


@pytest.mark.parametrize("chunk_bytes", [1 << 20, 3 << 20, (5 << 20) + 4096])
def test_lanes_pass_longreads_many_chunks(chunk_bytes, monkeypatch):
    """FQG_MAX_CHUNK_BYTES cuts a device-resident stream into small chunks: chunk boundaries fall inside long records, so every
    starting phase (j0) and the bridge chunks are exercised; the clean-data pass must take every large chunk."""
    import torch
    import fastq_utils_b200 as fq
    monkeypatch.setenv("FQG_MAX_CHUNK_BYTES", str(chunk_bytes))
    g = torch.Generator().manual_seed(5)
    nrec = 1500
    lens = torch.exp(torch.randn(nrec, generator=g) + 8.9).clamp(1000, 100000).to(torch.int64)
    hdr = fq.lib().fqg_synth_long_header_bytes()
    off = torch.zeros(nrec + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(hdr + 2 * lens + 4, 0)
    nb = int(off[-1])
    t = torch.zeros(nb + 64, dtype=torch.uint8, device="cuda")
    fq.synth_longreads(t, off.cuda(), 0, nrec, seed=5, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    data = bytes(t[:nb].cpu().numpy())
    rep, tr, pc = _run_ctx(fq.MODE_SINGLE, [(t.data_ptr(), nb)])
    assert tr == oracle_run(["-r", "a.fq"], data, None)
    assert pc["lanes_handed_on"] == 0 and pc["two_pass_fallbacks"] == 0 and pc["lanes"] >= nb // chunk_bytes, pc


@pytest.mark.parametrize("chunk_bytes", [1 << 20, (2 << 20) + 16])
def test_lanes_pass_illumina_many_chunks(chunk_bytes, lanes_mode, monkeypatch):
    import fastq_utils_b200 as fq
    monkeypatch.setenv("FQG_MAX_CHUNK_BYTES", str(chunk_bytes))
    n = 40_000
    t, nb = _illumina(n)
    data = bytes(t[:nb].cpu().numpy())
    for mode, argv in ((fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])):
        rep, tr, pc = _run_ctx(mode, [(t.data_ptr(), nb)], hint=n)
        assert tr == oracle_run(argv, data, None), argv
        assert pc["lanes_handed_on"] == 0 and pc["two_pass_fallbacks"] == 0 and pc["lanes"] >= nb // chunk_bytes, pc


def test_lanes_pass_winds_down_on_hostile_input(lanes_mode):
    """Inputs on which no tile finds a plus line to vote with (a long run of bytes without LF, old-style '+name' lines): the pass
    must give up quickly (every tile would otherwise wait for the tiles in front) and the per-record kernels decide."""
    import time
    import torch
    import fastq_utils_b200 as fq
    rb = fq.illumina_record_bytes()
    t, nb = _illumina(20_000)
    base = t[:nb].cpu().numpy().tobytes()
    zeros = base + b"\0" * (8 << 20)
    recs = base.split(b"\n")
    named = b"\n".join((b"+" + recs[i - 2][1:] if i % 4 == 2 else r) for i, r in enumerate(recs))
    for name, d in (("zero run", zeros), ("+name lines", named)):
        tt = torch.frombuffer(bytearray(d + b"\0" * 64), dtype=torch.uint8).cuda()
        t0 = time.perf_counter()
        rep, tr, pc = _run_ctx(fq.MODE_SINGLE, [(tt.data_ptr(), len(d))])
        dt = time.perf_counter() - t0
        assert tr == oracle_run(["-r", "a.fq"], d, None), name
        assert pc["lanes_handed_on"] == 1 and dt < 5.0, (name, pc, dt)


@pytest.mark.parametrize("seed", range(int(os.environ.get("FQG_FUZZ_SEEDS", "48"))))
def test_lanes_pass_fuzz_shapes(seed, lanes_mode, monkeypatch):
    """Clean (and now and then slightly broken) files of many shapes — 1-base to 5000-base reads, every read-name style, ragged
    lengths, with and without a final newline — through the clean-data pass on small inputs (FQG_FUSED_MIN_BYTES=1), fed from device
    memory in one to three pieces cut at random bytes, small chunks now and then: transcripts equal the oracle's."""
    import torch
    import fastq_utils_b200 as fq
    monkeypatch.setenv("FQG_FUSED_MIN_BYTES", "1")
    rng = random.Random(40_000 + seed)
    if rng.random() < 0.3:
        monkeypatch.setenv("FQG_MAX_CHUNK_BYTES", str(rng.choice([4096, 65536, 1 << 20])))
    style = rng.randrange(len(NAMES))
    lmax = rng.choice([1, 5, 40, 150, 600, 1000, 1500, 5000])
    lmin = lmax if rng.random() < 0.5 else 1
    n = rng.choice([1, 3, 40, 400, 3000]) if lmax <= 150 else rng.choice([1, 3, 40, 200])
    recs = make_file(rng, n, style, 1, seqlen=(lmin, lmax), qual=(rng.choice([14, 33, 35, 64]), rng.choice([74, 104, 126])))
    if rng.random() < 0.25:
        mutate(rng, recs)
    d = render(rng, recs, "lf")
    mode, argv = rng.choice([(fq.MODE_INDEX, ["a.fq"]), (fq.MODE_SINGLE, ["-r", "a.fq"])])
    tt = torch.frombuffer(bytearray(d + b"\0" * 64), dtype=torch.uint8).cuda()
    cuts = sorted(rng.sample(range(1, max(2, len(d))), k=min(rng.choice([0, 0, 1, 2]), max(0, len(d) - 1)))) if len(d) > 2 else []
    edges = [0] + cuts + [len(d)]
    pieces = [(tt.data_ptr() + a, b - a) for a, b in zip(edges[:-1], edges[1:]) if b > a] or [(tt.data_ptr(), 0)]
    rep, tr, pc = _run_ctx(mode, pieces, hint=n)
    assert tr == oracle_run(argv, d, None), (argv, lmin, lmax, n, cuts, pc)

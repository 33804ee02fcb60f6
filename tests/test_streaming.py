"""The read-name arena and the release of validated chunks: what the library keeps of a record is a copy of its
name, like the reference's new_indexentry (src/fastq.c:590-611); a chunk whose records raised no event is final and its bytes go.
Also the index's behaviour when two DIFFERENT names share a hash (the reference walks its chain, src/hash.c:38-45): forced with
the 12-bit test hash.  CPU: the stand-in device of tests/sim; `-m gpu`: the CUDA kernels."""
import ctypes
import os
import random
import subprocess
import sys

import pytest

from _util import ROOT, fqg_run, oracle_run

sys.path.insert(0, ROOT)


def _rec(i, mate=1, n=30):
    return f"@M01:5:FC:1:{1101 + i % 7}:{1000 + i}:{2000 + 3 * i} {mate}:N:0:ACGT\n{'ACGTN' * (n // 5)}\n+\n{'I' * (n - 1)}#\n"


def _job(kind, mode, pieces, pieces2=None, flags=0, env=None):
    """create / feed piece by piece / finish in a process of its own (library and test hooks are chosen per process) →
    (rendered transcript, memory stats, path counts)"""
    code = (
        "import sys, json, os\n"
        "import fastq_utils_b200 as fq\n"
        f"kind = {kind!r}\n"
        "if kind == 'sim':\n"
        "    from sim_lib import use_sim_library; use_sim_library()\n"
        "spec = json.load(open(sys.argv[1]))\n"
        "h = fq.FastqInfo(spec['mode'], flags=spec['flags'], index_capacity_hint=spec.get('hint', 0))\n"
        "log = []\n"
        "for f, key in ((0, 'pieces'), (1, 'pieces2')):\n"
        "    ps = spec.get(key)\n"
        "    if ps is None: continue\n"
        "    for i, hx in enumerate(ps):\n"
        "        h.feed(f, bytes.fromhex(hx), last=(i == len(ps) - 1))\n"
        "        log.append(h.memory_stats())\n"
        "rep = h.finish()\n"
        "tr = h.render(rep, 'a.fq', 'b.fq' if spec.get('pieces2') is not None else None)\n"
        "json.dump({'tr': list(tr), 'mem': h.memory_stats(), 'log': log, 'paths': h.path_counts()}, open(sys.argv[2], 'w'))\n")
    import json
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        spec = {"mode": mode, "flags": flags, "pieces": [p.hex() for p in pieces], "pieces2": [p.hex() for p in pieces2] if pieces2 is not None else None}
        json.dump(spec, open(os.path.join(d, "in.json"), "w"))
        e = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
        e.update(env or {})
        out = subprocess.run([sys.executable, "-c", code, os.path.join(d, "in.json"), os.path.join(d, "out.json")], env=e, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-3000:]
        return json.load(open(os.path.join(d, "out.json")))


def _pieces(data, k):
    return [data[i:i + k] for i in range(0, len(data), k)] or [b""]


def _check_streaming(kind, nrec, piece):
    import fastq_utils_b200 as fq
    recs = [_rec(i) for i in range(nrec)]
    data = "".join(recs).encode()
    mates = [_rec(i, 2) for i in range(nrec)]
    random.Random(5).shuffle(mates)
    data2 = "".join(mates).encode()
    # -- clean jobs: every chunk is final when the next one arrives, nothing but names stays
    for mode, argv, d2 in ((fq.MODE_INDEX, ["a.fq"], None), (fq.MODE_SINGLE, ["-r", "a.fq"], None), (fq.MODE_INDEX_PAIR, ["a.fq", "b.fq"], data2)):
        r = _job(kind, mode, _pieces(data, piece), _pieces(d2, piece) if d2 else None)
        assert tuple(r["tr"]) == oracle_run(argv, data, d2), (mode, r["tr"])
        m = r["mem"]
        total = len(data) + (len(d2) if d2 else 0)
        assert m["chunk_bytes_held"] == 0, m
        assert m["chunk_bytes_released"] >= total, m   # (bridge chunks are chunks too)
        assert m["records_final"] == nrec * (2 if d2 else 1), m
        assert max(x["chunk_bytes_held"] for x in r["log"]) <= 2 * piece + 4096, r["log"][:4]
        if mode != fq.MODE_SINGLE:
            per = 16 * ((len(recs[0].split(" ")[0]) - 1 + 15) // 16)
            assert 0 < m["arena_bytes"] <= (nrec * (2 if d2 else 1)) * (per + 16), m
    # -- an invalid base in the middle: the chunk that holds it and everything after it stay (the message needs the record), the
    # chunks in front of it are gone; the transcript is the reference's
    k = nrec // 2
    bad = list(recs); bad[k] = bad[k].replace("ACGTN", "ACXTN", 1)
    dbad = "".join(bad).encode()
    r = _job(kind, fq.MODE_INDEX, _pieces(dbad, piece))
    assert tuple(r["tr"]) == oracle_run(["a.fq"], dbad, None)
    assert 0 < r["mem"]["chunk_bytes_held"] < len(dbad) * 0.7 and r["mem"]["chunk_bytes_released"] > len(dbad) * 0.3, r["mem"]
    # -- a duplicated name far behind its first occurrence (whose chunk is gone): the name in the message comes from the arena
    dup = list(recs); dup[nrec - 3] = dup[5]
    ddup = "".join(dup).encode()
    r = _job(kind, fq.MODE_INDEX, _pieces(ddup, piece))
    assert tuple(r["tr"]) == oracle_run(["a.fq"], ddup, None), r["tr"]
    # -- mate loop: a mate without partner, and a partner claimed twice
    m2 = list(mates); m2[nrec // 3] = _rec(nrec + 77, 2)
    r = _job(kind, fq.MODE_INDEX_PAIR, _pieces(data, piece), _pieces("".join(m2).encode(), piece))
    assert tuple(r["tr"]) == oracle_run(["a.fq", "b.fq"], data, "".join(m2).encode()), r["tr"]
    m3 = list(mates); m3[nrec - 2] = m3[1]
    r = _job(kind, fq.MODE_INDEX_PAIR, _pieces(data, piece), _pieces("".join(m3).encode(), piece))
    assert tuple(r["tr"]) == oracle_run(["a.fq", "b.fq"], data, "".join(m3).encode()), r["tr"]
    # -- a NUL-led header line ends the file quietly (src/fastq.c:248): the statistics of the records in front of it, nothing behind it
    nul = list(recs); nul[k] = "\x00" + nul[k]
    dnul = "".join(nul).encode("latin-1")
    r = _job(kind, fq.MODE_INDEX, _pieces(dnul, piece))
    assert tuple(r["tr"]) == oracle_run(["a.fq"], dnul, None), r["tr"]
    r = _job(kind, fq.MODE_SINGLE, _pieces(dnul, piece))
    assert tuple(r["tr"]) == oracle_run(["-r", "a.fq"], dnul, None), r["tr"]


def test_sim_chunks_are_released_and_names_kept():
    _check_streaming("sim", 3000, 20_000)


@pytest.mark.gpu
def test_gpu_chunks_are_released_and_names_kept():
    _check_streaming("gpu", 40_000, 1_500_000)   # pieces large enough for the clean-data pass


def _check_weak_hash(kind, nrec):
    """12-bit hashes: thousands of different names share a hash.  Unique names must all be indexed, duplicates and unpaired mates
    found, and every mate must claim ITS name, exactly as with the full hash (and as the reference does it, with strcmp)."""
    import fastq_utils_b200 as fq
    env = {"FQG_TEST_WEAK_HASH": "1"}
    recs = [_rec(i) for i in range(nrec)]
    data = "".join(recs).encode()
    mates = [_rec(i, 2) for i in range(nrec)]
    random.Random(6).shuffle(mates)
    data2 = "".join(mates).encode()
    piece = max(20_000, len(data) // 5)
    r = _job(kind, fq.MODE_INDEX, _pieces(data, piece), env=env)
    assert tuple(r["tr"]) == oracle_run(["a.fq"], data, None)
    assert r["mem"]["collisions_walked"] > nrec // 4, r["mem"]
    r = _job(kind, fq.MODE_INDEX_PAIR, _pieces(data, piece), _pieces(data2, piece), env=env)
    assert tuple(r["tr"]) == oracle_run(["a.fq", "b.fq"], data, data2)
    dup = list(recs); dup[nrec - 3] = dup[5]; dup[nrec // 2] = dup[nrec // 2 - 40]
    ddup = "".join(dup).encode()
    r = _job(kind, fq.MODE_INDEX, _pieces(ddup, piece), env=env)
    assert tuple(r["tr"]) == oracle_run(["a.fq"], ddup, None), r["tr"]
    m2 = list(mates); m2[nrec // 3] = _rec(nrec + 77, 2); m2[nrec - 2] = m2[1]
    d2 = "".join(m2).encode()
    r = _job(kind, fq.MODE_INDEX_PAIR, _pieces(data, piece), _pieces(d2, piece), env=env)
    assert tuple(r["tr"]) == oracle_run(["a.fq", "b.fq"], data, d2), r["tr"]
    one_left = "".join(mates[:-1]).encode()
    r = _job(kind, fq.MODE_INDEX_PAIR, _pieces(data, piece), _pieces(one_left, piece), env=env)
    assert tuple(r["tr"]) == oracle_run(["a.fq", "b.fq"], data, one_left), r["tr"]


def test_sim_names_with_equal_hashes():
    _check_weak_hash("sim", 6000)


@pytest.mark.gpu
def test_gpu_names_with_equal_hashes():
    _check_weak_hash("gpu", 60_000)


@pytest.mark.parametrize("seed", range(40))
def test_sim_fuzz_weak_hash(seed, monkeypatch):
    """the differential fuzz of test_host_logic.py once more, under the 12-bit hash and with names that differ in one byte"""
    from test_oracle_fuzz import NAMES, make_file, mutate, render
    monkeypatch.setenv("FQG_TEST_WEAK_HASH", "1")
    rng = random.Random(77_000 + seed)
    style = rng.randrange(len(NAMES))
    n = rng.choice([40, 80, 200])
    two = rng.random() < 0.6
    r1 = make_file(rng, n, style, 1)
    r2 = make_file(rng, n, style, 2) if two else None
    if two and rng.random() < 0.7:
        rng.shuffle(r2)
    for _ in range(rng.choice([0, 1, 1, 2])):
        mutate(rng, r1 if (not two or rng.random() < 0.5) else r2)
    d1 = render(rng, r1, "lf")
    d2 = render(rng, r2, "lf") if two else None
    argv = ["a.fq"] + (["b.fq"] if two else [])
    assert fqg_run(argv, d1, d2, chunk=rng.choice([0, 300, 4000]), kind="sim") == oracle_run(argv, d1, d2), (argv, d1, d2)

"""Host-side twin of csrc/fq_synth.cu (numpy): the same synthetic Illumina records, byte for byte, without a GPU and without
loading libfastq_gpu.so — bench.py's reference arm feeds them to the unmodified reference binary, the tests compare them with
the device generator.  Also the transcript the reference prints for these workloads, by construction."""
import numpy as np

ILL_HDR, ILL_LEN = 55, 150
ILL_REC = ILL_HDR + ILL_LEN + 1 + 2 + ILL_LEN + 1  # 359
NAME_LEN = 38  # "A00123:45:HXXXXXXXX:1:TTTT:XXXXX:YYYYY"
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def _dec(out, col, v, digits):
    for d in range(digits - 1, -1, -1):
        out[:, col + d] = 48 + v % 10
        v = v // 10


def illumina(first, n, seed=42, mate=1, perm_window=0, block=200_000):
    """Records first .. first+n-1 of fqg_synth_illumina (same arguments) as bytes."""
    with np.errstate(over="ignore"):
        parts = []
        for b0 in range(0, n, block):
            k = min(block, n - b0)
            slot = np.arange(first + b0, first + b0 + k, dtype=np.uint64)
            i = slot
            if perm_window and perm_window > 1:
                w = np.uint64(perm_window)
                base = slot // w * w
                i = base + ((slot - base) * np.uint64(7) + np.uint64(3)) % w
            out = np.empty((k, ILL_REC), dtype=np.uint8)
            out[:, :23] = np.frombuffer(b"@A00123:45:HXXXXXXXX:1:", dtype=np.uint8)
            _dec(out, 23, (1101 + i // np.uint64(400000000)).astype(np.int64), 4)
            out[:, 27] = 58
            _dec(out, 28, (10000 + i % np.uint64(20000)).astype(np.int64), 5)
            out[:, 33] = 58
            _dec(out, 34, (10000 + (i // np.uint64(20000)) % np.uint64(20000)).astype(np.int64), 5)
            out[:, 39:55] = np.frombuffer(b" 1:N:0:ACGTACGT\n", dtype=np.uint8)
            out[:, 40] = 50 if mate == 2 else 49
            out[:, ILL_HDR + ILL_LEN] = 10
            out[:, ILL_HDR + ILL_LEN + 1] = 43
            out[:, ILL_HDR + ILL_LEN + 2] = 10
            out[:, ILL_REC - 1] = 10
            kk = np.arange(ILL_LEN, dtype=np.uint64)[None, :]
            h = _splitmix64(np.uint64(seed) ^ ((i[:, None] * np.uint64(0x100000001B3) + kk * np.uint64(2) + np.uint64(mate * 0x51ED27)) & _M))
            b = (h & np.uint64(3)).astype(np.uint8)
            sq = np.frombuffer(b"ACGT", dtype=np.uint8)[b]
            sq[((h >> np.uint64(8)) & np.uint64(1023)) == 0] = 78
            ql = (35 + ((h >> np.uint64(32)) % np.uint64(39))).astype(np.uint8)
            z = np.nonzero(i == 0)[0]
            if len(z):
                ql[z[0], 0], ql[z[0], 1] = 35, 73
            out[:, ILL_HDR:ILL_HDR + ILL_LEN] = sq
            out[:, ILL_HDR + ILL_LEN + 3:ILL_REC - 1] = ql
            parts.append(out.tobytes())
        return b"".join(parts)


def illumina_name(i):
    """normalised read name of record i (what the index holds)"""
    return "A00123:45:HXXXXXXXX:1:%04d:%05d:%05d" % (1101 + i // 400000000, 10000 + i % 20000, 10000 + (i // 20000) % 20000)


def perm_source(slot, w):
    """record number that fqg_synth_illumina(perm_window=w) puts into slot `slot` of the mate file"""
    base = slot // w * w
    return base + ((slot - base) * 7 + 3) % w


# ---------------------------------------------------------------------------------------------- expected transcripts
BANNER = "fastq_utils 0.25.3\n"
_BS = "\b" * 15


def _progress(n, per=100000):
    return "".join(_BS + str(c) for c in range(per, n + 1, per))


def _stats_block(nreads, qmin=35, qmax=73, enc="33", rl=(150, 150, 150)):
    return ("------------------------------------\n"
            f"Number of reads: {nreads}\nQuality encoding range: {qmin} {qmax}\nQuality encoding: {enc}\nRead length: {rl[0]} {rl[1]} {rl[2]}\nOK\n")


def expect_index(n, name1, name_len=NAME_LEN, sniff="CASAVA=1.8\n", stats=None, n2=None, name2=None, tail=None):
    """(rc, stdout, stderr) of `fastq_info name1 [name2]` on n unique clean reads (and n2 mates): src/fastq_info.c:273-395."""
    err = BANNER + "DEFAULT_HASHSIZE=39000001\n" + f"Scanning and indexing all reads from {name1}\n" + sniff + _progress(n)
    err += f"Scanning complete.\n\nReads processed: {n}\nMemory used in indexing: ~{(8 + n * (name_len + 41)) // (1024 * 1024)} MB\n"
    out = ""
    if name2 is not None:
        err += f"File {name1} processed\nNext file {name2}\n" + sniff + _progress(n2)
        out = "\n"
    if tail is not None:  # an error ends the transcript here
        return tail[0], tail[1], err + tail[2]
    return 0, out, err + (stats or _stats_block(n))


def expect_single(n, sniff="CASAVA=1.8\n", stats=None):
    return 0, "\n", BANNER + "Skipping check for duplicated read names\n" + sniff + _progress(n) + (stats or _stats_block(n))

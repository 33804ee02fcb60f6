"""ctypes binding of include/fastq_gpu.h.  Mirrors the reference's fastq_info entry points (src/fastq_info.c:190-396):
`fastq_info(argv, ...)` is main() on inflated streams, `FastqInfo` is the create / feed / finish loop-level API."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfastq_gpu.so")

MODE_SINGLE, MODE_INDEX, MODE_INDEX_PAIR, MODE_INTERLEAVED, MODE_SORTED_PAIR, MODE_READER = range(6)
KERNEL_CLASSES = ["scan", "records", "index", "mate", "pair", "other", "tile", "lanes"]
FLAG_PAIRED_NAMES, FLAG_EXTERNAL_INDEX, FLAG_TWO_PASS, FLAG_BORROW_FOR_CALL = 1, 2, 4, 8


class Config(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("device", ctypes.c_int32), ("index_capacity_hint", ctypes.c_uint64),
                ("flags", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 3)]


class FileReport(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_uint64), ("num_rds", ctypes.c_uint64), ("min_rl", ctypes.c_uint64),
                ("max_rl", ctypes.c_uint64), ("min_qual", ctypes.c_uint64), ("max_qual", ctypes.c_uint64),
                ("sniff_format", ctypes.c_int32), ("color_space", ctypes.c_int32)]


class Error(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("file", ctypes.c_int32), ("msg_file", ctypes.c_int32), ("chr", ctypes.c_int32),
                ("record", ctypes.c_uint64), ("line", ctypes.c_uint64), ("a", ctypes.c_uint64), ("b", ctypes.c_uint64), ("event_key", ctypes.c_uint64),
                ("hdr1_len", ctypes.c_uint32), ("hdr2_len", ctypes.c_uint32), ("name_len", ctypes.c_uint32),
                ("hdr1", ctypes.c_char * 1024), ("hdr2", ctypes.c_char * 1024), ("name", ctypes.c_char * 1024)]


class Report(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int32), ("reserved", ctypes.c_int32), ("file", FileReport * 2),
                ("n_index_entries", ctypes.c_uint64), ("n_index_left", ctypes.c_uint64), ("index_mem", ctypes.c_uint64),
                ("median_rl", ctypes.c_uint64), ("reads_before_error", ctypes.c_uint64 * 2), ("error", Error)]


class RenderOpts(ctypes.Structure):
    _fields_ = [("empty_ok", ctypes.c_int32), ("no_enc_ok", ctypes.c_int32), ("name1", ctypes.c_char_p), ("name2", ctypes.c_char_p)]


class Transcript(ctypes.Structure):
    _fields_ = [("rc", ctypes.c_int32), ("out", ctypes.c_void_p), ("out_len", ctypes.c_size_t),
                ("err", ctypes.c_void_p), ("err_len", ctypes.c_size_t)]


class StreamIO(ctypes.Structure):
    """fqg_stream_io: the caller's open / read / close (include/fastq_gpu.h)"""
    _fields_ = [("user", ctypes.c_void_p),
                ("open", ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p)),
                ("read", ctypes.CFUNCTYPE(ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)),
                ("close", ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p))]


class KernelStat(ctypes.Structure):
    _fields_ = [("ms", ctypes.c_double), ("launches", ctypes.c_uint64), ("bytes", ctypes.c_uint64), ("items", ctypes.c_uint64)]


CHUNK_HOOK = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int)

_lib = None


def lib():
    """Load libfastq_gpu.so (built in-tree by __graft_entry__.build() / csrc/Makefile).  Fails loudly when missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        _lib = bind(ctypes.CDLL(_SO))
    return _lib


def bind(L):
    """Declare the argument types of include/fastq_gpu.h on a loaded library."""
    vp, u64, sz, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    L.fqg_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(vp)]
    L.fqg_destroy.argtypes = [vp]
    L.fqg_destroy.restype = None
    L.fqg_feed.argtypes = [vp, ci, vp, sz, ci]
    L.fqg_feed_device.argtypes = [vp, ci, vp, sz, ci]
    L.fqg_finish.argtypes = [vp, ctypes.POINTER(Report)]
    L.fqg_reset.argtypes = [vp]
    L.fqg_last_error.argtypes = [vp]
    L.fqg_last_error.restype = ctypes.c_char_p
    L.fqg_launch_count.argtypes = [vp]
    L.fqg_launch_count.restype = u64
    L.fqg_path_counts.argtypes = [vp, ctypes.POINTER(u64 * 4)]
    L.fqg_memory_stats.argtypes = [vp, ctypes.POINTER(u64 * 5)]
    L.fqg_device_ms.argtypes = [vp]
    L.fqg_device_ms.restype = ctypes.c_double
    L.fqg_index_records.argtypes = [vp, vp, sz, ctypes.POINTER(u64), sz, ctypes.POINTER(u64)]
    L.fqg_render.argtypes = [ctypes.POINTER(Report), ctypes.POINTER(RenderOpts), ctypes.POINTER(Transcript)]
    L.fqg_transcript_free.argtypes = [ctypes.POINTER(Transcript)]
    L.fqg_transcript_free.restype = None
    L.fqg_fastq_info_mem.argtypes = [ci, ctypes.POINTER(ctypes.c_char_p), vp, sz, vp, sz, ci, sz, ctypes.POINTER(Transcript)]
    L.fqg_reader_tool_mem.argtypes = [ci, ctypes.POINTER(ctypes.c_char_p), vp, sz, ci, sz, ctypes.POINTER(Transcript)]
    L.fqg_trim_poly_at_stream.argtypes = [ci, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(StreamIO), ci, ctypes.POINTER(Transcript),
                                          ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_char_p)]
    L.fqg_buffer_free.argtypes = [vp]
    L.fqg_filterpair_mem.argtypes = [ci, ctypes.POINTER(ctypes.c_char_p), vp, sz, vp, sz, ci, ctypes.POINTER(Transcript), ctypes.POINTER(vp), ctypes.POINTER(sz),
                                     ctypes.POINTER(ctypes.c_int32)]
    L.fqg_buffer_free.restype = None
    L.fqg_kernel_stats.argtypes = [vp, ci, ctypes.POINTER(KernelStat)]
    L.fqg_kernel_stats_reset.argtypes = [vp]
    L.fqg_prescan_device.argtypes = [vp, ci, vp, sz, ci, ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(u64)]
    L.fqg_set_stream_start.argtypes = [vp, ci, ctypes.c_uint32, u64]
    L.fqg_names_count.argtypes = [vp, ci, ctypes.c_uint32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.fqg_names_pack.argtypes = [vp, ci, ctypes.c_uint32, vp, vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.fqg_shard_insert.argtypes = [vp, vp, u64, vp, ctypes.c_uint32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.fqg_shard_result.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(u64)]
    L.fqg_hist_range.argtypes = [vp, ci, u64, u64, ctypes.POINTER(u64)]
    L.fqg_shard_claim.argtypes = [vp, vp, u64, vp, ctypes.c_uint32, ctypes.POINTER(u64), ctypes.POINTER(u64), u64]
    L.fqg_shard_claim_result.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.fqg_sniff_device.argtypes = [vp, ci, vp, sz, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    L.fqg_set_sniff.argtypes = [vp, ci, ctypes.c_int32, ctypes.c_int32]
    L.fqg_set_file_total.argtypes = [vp, ci, u64]
    L.fqg_set_line_hint.argtypes = [vp, ci, ctypes.c_uint32]
    L.fqg_records_fed.argtypes = [vp, ci, ctypes.POINTER(u64)]
    L.fqg_set_hash_seed.argtypes = [vp, ctypes.c_uint32]
    L.fqg_set_chunk_hook.argtypes = [vp, CHUNK_HOOK, vp]
    L.fqg_names_new.argtypes = [vp, ci, ctypes.POINTER(u64)]
    L.fqg_names_pack_slots.argtypes = [vp, ci, ctypes.c_uint32, ctypes.POINTER(vp), u64, ctypes.c_uint32]
    L.fqg_shard_reserve.argtypes = [vp, u64]
    L.fqg_shard_insert_slots.argtypes = [vp, vp, ctypes.c_uint32, sz, ctypes.c_uint32, u64, ctypes.c_uint32, ci, vp, u64]
    L.fqg_shard_claim_slots.argtypes = [vp, vp, ctypes.c_uint32, sz, ctypes.c_uint32, u64, ctypes.c_uint32, ci, vp, u64]
    L.fqg_set_route.argtypes = [vp, ci, ctypes.c_uint32, ctypes.POINTER(vp), sz, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    L.fqg_route_chunks.argtypes = [vp, ci, ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int32)]
    L.fqg_route_blocks.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    L.fqg_side_mark.argtypes = [vp]
    L.fqg_order_after.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
    L.fqg_route_region_bytes.argtypes = [ctypes.c_uint32, u64, ctypes.c_uint32]
    L.fqg_route_region_bytes.restype = sz
    L.fqg_shard_slots_result.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.fqg_side_copy.argtypes = [vp, vp, vp, sz]
    L.fqg_side_sync.argtypes = [vp]
    L.fqg_side_copy_lane.argtypes = [vp, ctypes.c_int, vp, vp, sz]
    L.fqg_ipc_alloc.argtypes = [vp, sz, ctypes.POINTER(vp), ctypes.c_char_p]
    L.fqg_ipc_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.fqg_ipc_close.argtypes = [vp, vp]
    L.fqg_ipc_free.argtypes = [vp, vp]
    if hasattr(L, "fqg_synth_illumina"):  # absent from the test stand-in
        L.fqg_synth_illumina.argtypes = [vp, u64, u64, u64, ci, u64, vp]
        L.fqg_synth_longreads.argtypes = [vp, vp, u64, u64, u64, vp]
    return L


def _check(ctx, st, what):
    if st != 0:
        msg = lib().fqg_last_error(ctx).decode("latin-1") if ctx else ""
        raise RuntimeError(f"{what} failed with status {st}: {msg}")


def _take(tr):
    out = ctypes.string_at(tr.out, tr.out_len).decode("latin-1")
    err = ctypes.string_at(tr.err, tr.err_len).decode("latin-1")
    rc = tr.rc
    lib().fqg_transcript_free(ctypes.byref(tr))
    return rc, out, err


def fastq_info(argv, data1=None, data2=None, chunk=0, device=0):
    """`fastq_info argv...` on in-memory inflated streams → (exit status, stdout, stderr).  None = file could not be opened."""
    full = [b"fastq_info"] + [a.encode("latin-1") if isinstance(a, str) else a for a in argv]
    arr = (ctypes.c_char_p * (len(full) + 1))(*full, None)
    tr = Transcript()
    un = ctypes.c_size_t(-1).value
    st = lib().fqg_fastq_info_mem(len(full), arr, data1, len(data1) if data1 is not None else un,
                                  data2, len(data2) if data2 is not None else un, device, chunk, ctypes.byref(tr))
    if st != 0:
        raise RuntimeError(f"fqg_fastq_info_mem failed with status {st}")
    return _take(tr)


def reader_tool(tool, argv, data1=None, chunk=0, device=0):
    """`fastq_num_reads argv...` / `fastq_not_empty argv...` (src/fastq_num_reads.c, src/fastq_not_empty.c) on an in-memory inflated
    stream → (exit status, stdout, stderr).  data1=None = the file could not be opened."""
    full = [tool.encode("latin-1")] + [a.encode("latin-1") if isinstance(a, str) else a for a in argv]
    arr = (ctypes.c_char_p * (len(full) + 1))(*full, None)
    tr = Transcript()
    un = ctypes.c_size_t(-1).value
    st = lib().fqg_reader_tool_mem(len(full), arr, data1, len(data1) if data1 is not None else un, device, chunk, ctypes.byref(tr))
    if st != 0:
        raise RuntimeError(f"fqg_reader_tool_mem failed with status {st}")
    return _take(tr)


def _read_inflated(path):
    """a file as the reference's gzopen/gzread deliver it: gzip members inflated, anything else as it is"""
    import zlib
    raw = open(path, "rb").read()
    if raw[:2] != b"\x1f\x8b":
        return raw
    out = []
    while raw[:2] == b"\x1f\x8b":
        d = zlib.decompressobj(31)
        out.append(d.decompress(raw))
        raw = d.unused_data
    return b"".join(out)


def filterpair(argv, data1=None, data2=None, device=0, write=False, _lib=None):
    """`fastq_filterpair argv...` (src/fastq_filterpair.c; argv = fastq1 fastq2 paired1 paired2 unpaired [sorted]) on two in-memory inflated
    streams → (exit status, stdout, stderr, created, [paired1, paired2, unpaired] inflated).  data=None: the file could not be opened.
    write=True gzips the three results into the named files (level 3, like the reference)."""
    full = [b"fastq_filterpair"] + [a.encode("latin-1") if isinstance(a, str) else a for a in argv]
    arr = (ctypes.c_char_p * (len(full) + 1))(*full, None)
    un = ctypes.c_size_t(-1).value
    tr, outs, lens, created = Transcript(), (ctypes.c_void_p * 3)(), (ctypes.c_size_t * 3)(), ctypes.c_int32(0)
    L = _lib if _lib is not None else lib()  # (_lib: the tests' stand-in library)
    st = L.fqg_filterpair_mem(len(full), arr, data1, len(data1) if data1 is not None else un, data2, len(data2) if data2 is not None else un, device,
                              ctypes.byref(tr), outs, lens, ctypes.byref(created))
    if st != 0:
        raise RuntimeError(f"fqg_filterpair_mem failed with status {st}")
    bufs = []
    for i in range(3):
        bufs.append(ctypes.string_at(outs[i], lens[i]) if outs[i] else b"")
        L.fqg_buffer_free(outs[i])
    rc, out, err = tr.rc, ctypes.string_at(tr.out, tr.out_len).decode("latin-1"), ctypes.string_at(tr.err, tr.err_len).decode("latin-1")
    L.fqg_transcript_free(ctypes.byref(tr))
    if write and created.value and len(argv) >= 5:
        import gzip
        for name, b in zip(argv[2:5], bufs):
            with gzip.open(name, "wb", compresslevel=3) as fh:
                fh.write(b)
    return rc, out, err, bool(created.value), bufs


def trim_poly_at(argv, files=None, device=0, write=False, _lib=None):
    """`fastq_trim_poly_at argv...` (src/fastq_trim_poly_at.c) → (exit status, stdout, stderr, name of the output file or None, its
    inflated contents).  files: operand name → inflated bytes (a missing name cannot be opened); None reads the named file from disk.
    write=True gzips the result into the output file like the reference does (level 4, src/fastq_trim_poly_at.c:201)."""
    full = [b"fastq_trim_poly_at"] + [a.encode("latin-1") if isinstance(a, str) else a for a in argv]
    arr = (ctypes.c_char_p * (len(full) + 1))(*full, None)
    state = {}

    def _open(user, name):
        try:
            data = files[name.decode("latin-1")] if files is not None else _read_inflated(name.decode("latin-1"))
        except (KeyError, OSError):
            return None
        h = len(state) + 1
        state[h] = [data, 0]
        return h

    def _read(user, h, buf, cap):
        data, pos = state[h]
        k = min(cap, len(data) - pos)
        ctypes.memmove(buf, data[pos:pos + k], k)
        state[h][1] = pos + k
        return k

    def _close(user, h):
        state.pop(h, None)
    io = StreamIO(None, StreamIO._fields_[1][1](_open), StreamIO._fields_[2][1](_read), StreamIO._fields_[3][1](_close))
    tr, buf, n, name = Transcript(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_char_p()
    L = _lib if _lib is not None else lib()  # (_lib: the tests' stand-in library)
    st = L.fqg_trim_poly_at_stream(len(full), arr, ctypes.byref(io), device, ctypes.byref(tr), ctypes.byref(buf), ctypes.byref(n), ctypes.byref(name))
    if st != 0:
        raise RuntimeError(f"fqg_trim_poly_at_stream failed with status {st}")
    data = ctypes.string_at(buf, n.value) if buf else b""
    L.fqg_buffer_free(buf)
    oname = name.value.decode("latin-1") if name.value is not None else None
    rc, out, err = tr.rc, ctypes.string_at(tr.out, tr.out_len).decode("latin-1"), ctypes.string_at(tr.err, tr.err_len).decode("latin-1")
    L.fqg_transcript_free(ctypes.byref(tr))
    if write and oname is not None:
        import gzip
        with gzip.open(oname, "wb", compresslevel=4) as fh:
            fh.write(data)
    return rc, out, err, oname, data


class FastqInfo:
    """One validation run: feed the inflated bytes of the file(s), then finish() for the report."""

    def __init__(self, mode, device=0, index_capacity_hint=0, flags=0):
        cfg = Config(mode=mode, device=device, index_capacity_hint=index_capacity_hint, flags=flags)
        self._ctx = ctypes.c_void_p()
        st = lib().fqg_create(ctypes.byref(cfg), ctypes.byref(self._ctx))
        if st != 0:
            raise RuntimeError(f"fqg_create failed with status {st} (-1 = no CUDA device; there is no CPU fallback)")
        self.mode = mode

    def close(self):
        if self._ctx:
            lib().fqg_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def feed(self, file, data, last=True):
        """Host bytes (bytes / bytearray / anything exposing a pointer via torch/numpy `data_ptr`-like int tuple (ptr, n))."""
        if isinstance(data, tuple):
            ptr, n = data
            ptr = ctypes.c_void_p(ptr)
        else:
            n = len(data)
            ptr = ctypes.cast(ctypes.c_char_p(bytes(data)) if not isinstance(data, bytes) else ctypes.c_char_p(data), ctypes.c_void_p)
        _check(self._ctx, lib().fqg_feed(self._ctx, file, ptr, n, 1 if last else 0), "fqg_feed")

    def feed_device(self, file, ptr, n, last=True):
        _check(self._ctx, lib().fqg_feed_device(self._ctx, file, ctypes.c_void_p(ptr), n, 1 if last else 0), "fqg_feed_device")

    def finish(self):
        rep = Report()
        _check(self._ctx, lib().fqg_finish(self._ctx, ctypes.byref(rep)), "fqg_finish")
        return rep

    def reset(self):
        _check(self._ctx, lib().fqg_reset(self._ctx), "fqg_reset")

    def render(self, rep, name1, name2=None, empty_ok=False, no_enc_ok=False):
        o = RenderOpts(empty_ok=int(empty_ok), no_enc_ok=int(no_enc_ok), name1=name1.encode("latin-1"),
                       name2=name2.encode("latin-1") if name2 is not None else None)
        tr = Transcript()
        _check(self._ctx, lib().fqg_render(ctypes.byref(rep), ctypes.byref(o), ctypes.byref(tr)), "fqg_render")
        return _take(tr)

    def launch_count(self):
        return int(lib().fqg_launch_count(self._ctx))

    def path_counts(self):
        """chunks by validating path: {lanes, lanes_handed_on, tile, two_pass_fallbacks}"""
        out = (ctypes.c_uint64 * 4)()
        _check(self._ctx, lib().fqg_path_counts(self._ctx, ctypes.byref(out)), "fqg_path_counts")
        return dict(zip(("lanes", "lanes_handed_on", "tile", "two_pass_fallbacks"), (int(x) for x in out)))

    def memory_stats(self):
        """{chunk_bytes_held, chunk_bytes_released, arena_bytes, records_final, collisions_walked} (include/fastq_gpu.h: fqg_memory_stats)"""
        out = (ctypes.c_uint64 * 5)()
        _check(self._ctx, lib().fqg_memory_stats(self._ctx, ctypes.byref(out)), "fqg_memory_stats")
        return dict(zip(("chunk_bytes_held", "chunk_bytes_released", "arena_bytes", "records_final", "collisions_walked"), (int(x) for x in out)))

    def device_ms(self):
        return float(lib().fqg_device_ms(self._ctx))

    def kernel_stats(self, reset=False):
        out = {}
        for i, name in enumerate(KERNEL_CLASSES):
            ks = KernelStat()
            _check(self._ctx, lib().fqg_kernel_stats(self._ctx, i, ctypes.byref(ks)), "fqg_kernel_stats")
            out[name] = {"ms": ks.ms, "launches": int(ks.launches), "bytes": int(ks.bytes), "items": int(ks.items)}
        if reset:
            lib().fqg_kernel_stats_reset(self._ctx)
        return out

    # ---- multi-GPU building blocks (see dist.py) ----
    def prescan_device(self, file, ptr, n, at_eof):
        nl, lf, first = ctypes.c_uint64(), ctypes.c_int32(), (ctypes.c_uint64 * 4)()
        _check(self._ctx, lib().fqg_prescan_device(self._ctx, file, ctypes.c_void_p(ptr), n, int(at_eof), ctypes.byref(nl), ctypes.byref(lf), first), "fqg_prescan_device")
        return int(nl.value), bool(lf.value), [int(x) for x in first]

    def set_stream_start(self, file, skip_lines, first_record):
        _check(self._ctx, lib().fqg_set_stream_start(self._ctx, file, skip_lines, first_record), "fqg_set_stream_start")

    def names_count(self, file, world):
        c, b = (ctypes.c_uint64 * world)(), (ctypes.c_uint64 * world)()
        _check(self._ctx, lib().fqg_names_count(self._ctx, file, world, c, b), "fqg_names_count")
        return list(c), list(b)

    def names_pack(self, file, world, meta_ptr, blob_ptr, meta_base, blob_base):
        mb, bb = (ctypes.c_uint64 * world)(*meta_base), (ctypes.c_uint64 * world)(*blob_base)
        _check(self._ctx, lib().fqg_names_pack(self._ctx, file, world, ctypes.c_void_p(meta_ptr), ctypes.c_void_p(blob_ptr), mb, bb), "fqg_names_pack")

    def shard_insert(self, meta_ptr, n, blob_ptr, meta_start, blob_start):
        ns = len(blob_start)
        ms, bs = (ctypes.c_uint64 * (ns + 1))(*meta_start), (ctypes.c_uint64 * ns)(*blob_start)
        _check(self._ctx, lib().fqg_shard_insert(self._ctx, ctypes.c_void_p(meta_ptr), n, ctypes.c_void_p(blob_ptr), ns, ms, bs), "fqg_shard_insert")

    def shard_result(self):
        key, rec, ln, col = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint64()
        name = ctypes.create_string_buffer(1024)
        _check(self._ctx, lib().fqg_shard_result(self._ctx, ctypes.byref(key), ctypes.byref(rec), name, ctypes.byref(ln), ctypes.byref(col)), "fqg_shard_result")
        return int(key.value), int(rec.value), name.raw[:ln.value], int(col.value)

    def shard_claim(self, meta_ptr, n, blob_ptr, meta_start, blob_start, step_base):
        ns = len(blob_start)
        ms, bs = (ctypes.c_uint64 * (ns + 1))(*meta_start), (ctypes.c_uint64 * ns)(*blob_start)
        _check(self._ctx, lib().fqg_shard_claim(self._ctx, ctypes.c_void_p(meta_ptr), n, ctypes.c_void_p(blob_ptr), ns, ms, bs, step_base), "fqg_shard_claim")

    def shard_claim_result(self):
        key, rec, ln, cl, col = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_uint64()
        name = ctypes.create_string_buffer(1024)
        _check(self._ctx, lib().fqg_shard_claim_result(self._ctx, ctypes.byref(key), ctypes.byref(rec), name, ctypes.byref(ln), ctypes.byref(cl), ctypes.byref(col)), "fqg_shard_claim_result")
        return int(key.value), int(rec.value), name.raw[:ln.value], int(cl.value), int(col.value)

    # ---- pipelined routing (dist.py `_route_round`) ----
    def set_chunk_hook(self, fn):
        """fn(file) is called from inside feed_device right after a chunk's clean-data pass was launched; None removes it."""
        self._hook = CHUNK_HOOK((lambda user, file: fn(file))) if fn is not None else CHUNK_HOOK()
        _check(self._ctx, lib().fqg_set_chunk_hook(self._ctx, self._hook, None), "fqg_set_chunk_hook")

    def names_new(self, file):
        n = ctypes.c_uint64()
        _check(self._ctx, lib().fqg_names_new(self._ctx, file, ctypes.byref(n)), "fqg_names_new")
        return int(n.value)

    def names_pack_slots(self, file, region_ptrs, cap, units=0):
        arr = (ctypes.c_void_p * len(region_ptrs))(*region_ptrs)
        _check(self._ctx, lib().fqg_names_pack_slots(self._ctx, file, len(region_ptrs), arr, cap, units), "fqg_names_pack_slots")

    def shard_reserve(self, n_names):
        _check(self._ctx, lib().fqg_shard_reserve(self._ctx, n_names), "fqg_shard_reserve")

    def shard_insert_slots(self, regions_ptr, n_src, region_bytes, nblocks, stride, beside, units=0, flags_ptr=0, expect=0):
        _check(self._ctx, lib().fqg_shard_insert_slots(self._ctx, ctypes.c_void_p(regions_ptr), n_src, region_bytes, nblocks, stride, units, 1 if beside else 0,
                                                       ctypes.c_void_p(flags_ptr), expect), "fqg_shard_insert_slots")

    def shard_claim_slots(self, regions_ptr, n_src, region_bytes, nblocks, stride, beside, units, flags_ptr=0, expect=0):
        _check(self._ctx, lib().fqg_shard_claim_slots(self._ctx, ctypes.c_void_p(regions_ptr), n_src, region_bytes, nblocks, stride, units, 1 if beside else 0,
                                                      ctypes.c_void_p(flags_ptr), expect), "fqg_shard_claim_slots")

    def set_route(self, file, region_ptrs, region_bytes, depth, stride, units):
        """the clean-data pass of every chunk of `file` writes the names into per-owner regions itself (include/fastq_gpu.h); [] switches it off"""
        arr = (ctypes.c_void_p * max(1, len(region_ptrs)))(*region_ptrs)
        _check(self._ctx, lib().fqg_set_route(self._ctx, file, len(region_ptrs), arr, region_bytes, depth, stride, units), "fqg_set_route")

    def route_chunks(self, file):
        n, b = ctypes.c_uint64(), ctypes.c_int32()
        _check(self._ctx, lib().fqg_route_chunks(self._ctx, file, ctypes.byref(n), ctypes.byref(b)), "fqg_route_chunks")
        return int(n.value), bool(b.value)

    def route_blocks(self):
        n = ctypes.c_uint32()
        _check(self._ctx, lib().fqg_route_blocks(self._ctx, ctypes.byref(n)), "fqg_route_blocks")
        return int(n.value)

    def order_after(self, my_side, earlier, their_side):
        """what this context queues from now on (side / main stream) starts after what `earlier` has queued so far"""
        _check(self._ctx, lib().fqg_order_after(self._ctx, int(bool(my_side)), earlier._ctx, int(bool(their_side))), "fqg_order_after")

    def side_mark(self):
        _check(self._ctx, lib().fqg_side_mark(self._ctx), "fqg_side_mark")

    def shard_slots_result(self):
        """(inserted, names already there / equal hashes, overflow, claimed by mates, mates without a fresh partner)"""
        ins, eq, ov, cl, un = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self._ctx, lib().fqg_shard_slots_result(self._ctx, ctypes.byref(ins), ctypes.byref(eq), ctypes.byref(ov), ctypes.byref(cl), ctypes.byref(un)), "fqg_shard_slots_result")
        return int(ins.value), int(eq.value), bool(ov.value), int(cl.value), int(un.value)

    def side_copy(self, dst, src, n):
        _check(self._ctx, lib().fqg_side_copy(self._ctx, ctypes.c_void_p(dst), ctypes.c_void_p(src), n), "fqg_side_copy")

    def side_copy_lane(self, lane, dst, src, n):
        _check(self._ctx, lib().fqg_side_copy_lane(self._ctx, int(lane), ctypes.c_void_p(dst), ctypes.c_void_p(src), n), "fqg_side_copy_lane")

    def side_sync(self):
        _check(self._ctx, lib().fqg_side_sync(self._ctx), "fqg_side_sync")

    def ipc_alloc(self, nbytes):
        p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        _check(self._ctx, lib().fqg_ipc_alloc(self._ctx, nbytes, ctypes.byref(p), h), "fqg_ipc_alloc")
        return int(p.value), h.raw

    def ipc_open(self, handle):
        p = ctypes.c_void_p()
        _check(self._ctx, lib().fqg_ipc_open(self._ctx, ctypes.create_string_buffer(handle, 64), ctypes.byref(p)), "fqg_ipc_open")
        return int(p.value)

    def ipc_close(self, ptr):
        _check(self._ctx, lib().fqg_ipc_close(self._ctx, ctypes.c_void_p(ptr)), "fqg_ipc_close")

    def ipc_free(self, ptr):
        _check(self._ctx, lib().fqg_ipc_free(self._ctx, ctypes.c_void_p(ptr)), "fqg_ipc_free")

    def sniff_device(self, file, ptr, n, skip):
        f, c = ctypes.c_int32(), ctypes.c_int32()
        _check(self._ctx, lib().fqg_sniff_device(self._ctx, file, ctypes.c_void_p(ptr), n, skip, ctypes.byref(f), ctypes.byref(c)), "fqg_sniff_device")
        return int(f.value), int(c.value)

    def set_sniff(self, file, fmt, color):
        _check(self._ctx, lib().fqg_set_sniff(self._ctx, file, fmt, color), "fqg_set_sniff")

    def set_hash_seed(self, seed):
        _check(self._ctx, lib().fqg_set_hash_seed(self._ctx, seed), "fqg_set_hash_seed")

    def records_fed(self, file):
        n = ctypes.c_uint64()
        _check(self._ctx, lib().fqg_records_fed(self._ctx, file, ctypes.byref(n)), "fqg_records_fed")
        return int(n.value)

    def set_line_hint(self, file, seq_line_len):
        _check(self._ctx, lib().fqg_set_line_hint(self._ctx, file, seq_line_len), "fqg_set_line_hint")

    def set_file_total(self, file, total):
        _check(self._ctx, lib().fqg_set_file_total(self._ctx, file, total), "fqg_set_file_total")

    def hist_range(self, file, lo, hi):
        """records by read length lo..hi → numpy uint64 (a long-read file spans 100 000 bins: no Python list)"""
        import numpy as np
        out = np.zeros(hi - lo + 1, dtype=np.uint64)
        _check(self._ctx, lib().fqg_hist_range(self._ctx, file, lo, hi, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))), "fqg_hist_range")
        return out

    def index_records(self, data, cap=0):
        n = ctypes.c_uint64()
        starts = (ctypes.c_uint64 * max(cap, 1))()
        _check(self._ctx, lib().fqg_index_records(self._ctx, data, len(data), starts, cap, ctypes.byref(n)), "fqg_index_records")
        return int(n.value), list(starts[:min(cap, n.value)])


def route_region_bytes(nblocks, stride, units):
    """bytes of one routing region (include/fastq_gpu.h: fqg_route_region_bytes)"""
    return int(lib().fqg_route_region_bytes(nblocks, stride, units))


def feed_chunk_bytes():
    """Largest piece fqg_feed_device hands to the kernels at once (fq_engine.cpp: kMaxChunk, test hook FQG_MAX_CHUNK_BYTES)."""
    e = os.environ.get("FQG_MAX_CHUNK_BYTES")
    v = (int(e) & ~15) if e else 0
    return v if 4096 <= v < (1 << 31) else (1 << 31) - 16


def illumina_record_bytes():
    return int(lib().fqg_synth_illumina_record_bytes())


def synth_illumina(tensor, first_record, n_records, seed=42, mate=1, perm_window=0, stream=0):
    """Fill a CUDA uint8 tensor (≥ n_records*359 bytes) with synthetic Illumina records first_record..+n_records."""
    st = lib().fqg_synth_illumina(ctypes.c_void_p(tensor.data_ptr()), first_record, n_records, seed, mate, perm_window, ctypes.c_void_p(stream))
    if st != 0:
        raise RuntimeError(f"fqg_synth_illumina failed with status {st}")


def synth_longreads(tensor, offsets, first_record, n_records, seed=7, stream=0):
    st = lib().fqg_synth_longreads(ctypes.c_void_p(tensor.data_ptr()), ctypes.c_void_p(offsets.data_ptr()), first_record, n_records, seed, ctypes.c_void_p(stream))
    if st != 0:
        raise RuntimeError(f"fqg_synth_longreads failed with status {st}")

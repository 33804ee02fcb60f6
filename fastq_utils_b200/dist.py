"""Multi-GPU fastq_info: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) for
the plumbing, libfastq_gpu for every byte of work.

Each rank holds a contiguous byte range of the (single) input file.

  1. every rank builds the line index of its range (`fqg_prescan_device`); an all-gather of (line feeds, ends-with-LF,
     first line ends) fixes each range's line phase; the bytes before a range's first record start are sent to the
     previous rank, which appends them to its stream (the library's chunk-bridging joins them with its tail);
  2. validation runs locally with global record indices (`fqg_set_stream_start`);
  3. default mode: read names are routed by hash to their owner rank with two all-to-alls (24-byte tuples + name bytes);
     the owner inserts them into its shard of the index with exact byte compares (`fqg_shard_insert`);
  4. the earliest event (smallest key in the reference's sequential order) wins; statistics are all-reduced and the
     merged report is rendered with `fqg_render`, so the text and exit status equal the reference's.

Supported modes: MODE_SINGLE (-r) and MODE_INDEX (default, one file).  A clean early end of file caused by a NUL-led
header line (src/fastq.c:248) is reported as unsupported in the sharded path (the single-GPU path handles it).
"""
import ctypes

import torch
import torch.distributed as dist

from . import api

KEY_NONE = (1 << 64) - 1
R_STOP, R_NAME = 0, 3
E_DUP = 13
MAX_READ_LENGTH = 2_500_000


def _key(step, rank):
    return (step << 6) | rank


class ShardedFastqInfo:
    def __init__(self, mode, device=0, n_hint=0, tensor_device=None):
        if mode not in (api.MODE_SINGLE, api.MODE_INDEX):
            raise NotImplementedError("sharded runs support MODE_SINGLE and MODE_INDEX")
        self.mode = mode
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.tdev = tensor_device if tensor_device is not None else torch.device("cuda", device)
        flags = api.FLAG_EXTERNAL_INDEX if mode == api.MODE_INDEX else 0
        self.ctx = api.FastqInfo(mode, device=device, flags=flags)
        self.shard = api.FastqInfo(api.MODE_INDEX, device=device, index_capacity_hint=n_hint) if mode == api.MODE_INDEX else None
        self._keep = []

    # ------------------------------------------------------------------ helpers
    def _gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def _a2a(self, send, in_splits, out_splits):
        recv = torch.empty(sum(out_splits) + 64, dtype=torch.uint8, device=self.tdev)
        if self.world == 1:
            recv[:sum(out_splits)] = send[:sum(in_splits)]
        else:
            dist.all_to_all_single(recv[:sum(out_splits)], send[:sum(in_splits)], output_split_sizes=out_splits, input_split_sizes=in_splits)
        return recv

    # ------------------------------------------------------------------ one job
    def run_device(self, ptr, nbytes, name="-", empty_ok=False, no_enc_ok=False):
        """ptr/nbytes: this rank's byte range in device memory (64 readable bytes must follow).  Returns a dict with the
        merged report fields and, on rank 0, the rendered (rc, stdout, stderr)."""
        W, r = self.world, self.rank
        self.ctx.reset()
        if self.shard:
            self.shard.reset()
        self._keep = []
        last_rank = r == W - 1
        # -- 1. line phase
        if nbytes > 0:
            nlines, ends_lf, first = self.ctx.prescan_device(0, ptr, nbytes, at_eof=last_rank)
        else:
            nlines, ends_lf, first = 0, True, [KEY_NONE] * 4
        lfs = nlines - (1 if (last_rank and nbytes > 0 and not ends_lf) else 0)
        info = self._gather((lfs, ends_lf, first, nbytes))
        G = [0] * (W + 1)
        for i in range(W):
            G[i + 1] = G[i] + info[i][0]
        skip, firstrec, cut = [0] * W, [0] * W, [0] * W
        degenerate = False
        for i in range(W):
            prev_lf = True if i == 0 else info[i - 1][1]
            if i == 0 or (prev_lf and G[i] % 4 == 0):
                skip[i] = 0
            else:
                skip[i] = (4 - G[i] % 4) % 4 or 4
            if i > 0 and (info[i][0] < skip[i] or info[i][3] == 0):
                degenerate = True  # a range without a record start of its own (tiny inputs)
                break
            firstrec[i] = (G[i] + skip[i]) // 4
            cut[i] = info[i][2][skip[i] - 1] if skip[i] > 0 else 0
        head = None
        if degenerate:
            # tiny input: everything goes to rank 0, the other ranks hold an empty stream (the collectives below still run)
            sizes = [info[i][3] for i in range(W)]
            if r == 0:
                whole = torch.zeros(sum(sizes) + 64, dtype=torch.uint8, device=self.tdev)
                if sizes[0]:
                    whole[:sizes[0]] = _as_tensor(ptr, sizes[0], self.tdev)
                off = sizes[0]
                for s in range(1, W):
                    if sizes[s]:
                        dist.recv(whole[off:off + sizes[s]], s)
                    off += sizes[s]
                self._keep.append(whole)
                self.ctx.set_stream_start(0, 0, 0)
                if off:
                    self.ctx.feed_device(0, whole.data_ptr(), off, last=True)
                else:
                    self.ctx.feed(0, b"", last=True)
            else:
                if nbytes:
                    dist.send(_as_tensor(ptr, nbytes, self.tdev), 0)
                self.ctx.set_stream_start(0, 0, G[W] // 4)
                self.ctx.feed(0, b"", last=True)
            lfs_local = None
        else:
            # -- head of range r+1 travels to rank r
            reqs = []
            if W > 1:
                if r > 0 and cut[r] > 0:
                    view = _as_tensor(ptr, cut[r], self.tdev)
                    reqs.append(dist.isend(view, r - 1))
                    self._keep.append(view)
                if r < W - 1 and cut[r + 1] > 0:
                    head = torch.zeros(cut[r + 1] + 64, dtype=torch.uint8, device=self.tdev)
                    reqs.append(dist.irecv(head[:cut[r + 1]], r + 1))
                for q in reqs:
                    q.wait()
            # -- 2. local validation with global indices
            self.ctx.set_stream_start(0, skip[r], firstrec[r])
            if last_rank:
                self.ctx.feed_device(0, ptr, nbytes, last=True) if nbytes else self.ctx.feed(0, b"", last=True)
            else:
                if nbytes:
                    self.ctx.feed_device(0, ptr, nbytes, last=False)
                if head is not None:
                    self.ctx.feed_device(0, head.data_ptr(), cut[r + 1], last=True)
                    self._keep.append(head)
                else:
                    self.ctx.feed(0, b"", last=True)
            lfs_local = lfs
        rep = self.ctx.finish()
        local_key = rep.error.event_key if rep.error.code != 0 else KEY_NONE
        expected = 0 if lfs_local is None else (lfs + (1 if (last_rank and nbytes > 0 and not ends_lf) else 0) - skip[r] + (skip[r + 1] if r < W - 1 else 0)) // 4
        if rep.file[0].n_records < expected and rep.error.code == 0:
            raise NotImplementedError("NUL-led header line (early clean EOF) in a sharded run")
        if rep.error.code != 0 and (local_key & 63) == R_STOP:
            raise NotImplementedError("NUL-led header line (early clean EOF) in a sharded run")
        # -- 3. names to their owners
        dup = (KEY_NONE, 0, b"")
        if self.shard is not None:
            counts, nbytes_names = self.ctx.names_count(0, W)
            theirs = self._gather((counts, nbytes_names))
            in_meta = [c * 24 for c in counts]
            in_blob = list(nbytes_names)
            out_cnt = [theirs[s][0][r] for s in range(W)]
            out_meta = [c * 24 for c in out_cnt]
            out_blob = [theirs[s][1][r] for s in range(W)]
            meta_base, blob_base = [0] * W, [0] * W
            for o in range(1, W):
                meta_base[o] = meta_base[o - 1] + counts[o - 1]
                blob_base[o] = blob_base[o - 1] + nbytes_names[o - 1]
            send_meta = torch.empty(sum(in_meta) + 64, dtype=torch.uint8, device=self.tdev)
            send_blob = torch.empty(sum(in_blob) + 64, dtype=torch.uint8, device=self.tdev)
            self.ctx.names_pack(0, W, send_meta.data_ptr(), send_blob.data_ptr(), meta_base, blob_base)
            if self.tdev.type == "cuda":
                torch.cuda.synchronize()
            recv_meta = self._a2a(send_meta, in_meta, out_meta)
            recv_blob = self._a2a(send_blob, in_blob, out_blob)
            if self.tdev.type == "cuda":
                torch.cuda.synchronize()
            ms, bs = [0], []
            acc = 0
            for s in range(W):
                ms.append(ms[-1] + out_cnt[s])
                bs.append(acc)
                acc += out_blob[s]
            self.shard.shard_insert(recv_meta.data_ptr(), ms[-1], recv_blob.data_ptr(), ms, bs)
            dkey, drec, dname, coll = self.shard.shard_result()
            if sum(self._gather(coll)):
                raise NotImplementedError("64-bit name hash collision between different names in a sharded run: rerun with another seed")
            dup = (dkey, drec, dname)
            self._keep += [send_meta, send_blob, recv_meta, recv_blob]
        # -- 4. merge
        f0 = rep.file[0]
        mine = {"key": local_key, "dup": dup, "err": bytes(ctypes.string_at(ctypes.addressof(rep.error), ctypes.sizeof(api.Error))) if local_key != KEY_NONE else None,
                "nrec": int(f0.n_records), "num_rds": int(f0.num_rds), "min_rl": int(f0.min_rl), "max_rl": int(f0.max_rl),
                "min_q": int(f0.min_qual), "max_q": int(f0.max_qual), "names": int(rep.n_index_entries), "mem": int(rep.index_mem) - 8,
                "sniff": (int(f0.sniff_format), int(f0.color_space)), "rbe": int(rep.reads_before_error[0])}
        allr = self._gather(mine)
        best = min(min(a["key"], a["dup"][0]) for a in allr)
        merged = api.Report()
        merged.mode = self.mode
        m0 = merged.file[0]
        m0.n_records = sum(a["nrec"] for a in allr)
        m0.num_rds = sum(a["num_rds"] for a in allr)
        m0.min_rl = min(a["min_rl"] for a in allr)
        m0.max_rl = max(a["max_rl"] for a in allr)
        m0.min_qual = min(a["min_q"] for a in allr)
        m0.max_qual = max(a["max_q"] for a in allr)
        m0.sniff_format, m0.color_space = next((a["sniff"] for a in allr if a["nrec"] > 0), allr[0]["sniff"])  # the rank holding record 0
        merged.file[1].sniff_format = merged.file[1].color_space = -1
        merged.n_index_entries = sum(a["names"] for a in allr)
        merged.n_index_left = merged.n_index_entries
        merged.index_mem = 8 + sum(a["mem"] for a in allr)
        merged.reads_before_error[0] = m0.n_records
        if best != KEY_NONE:
            holder = next(a for a in allr if min(a["key"], a["dup"][0]) == best)
            if holder["key"] == best:
                ctypes.memmove(ctypes.addressof(merged.error), holder["err"], ctypes.sizeof(api.Error))
            else:
                e = merged.error
                e.code, e.file, e.msg_file, e.record = E_DUP, 0, 0, holder["dup"][1]
                e.line = 4 * (holder["dup"][1] + 1)
                e.event_key = best
                nm = holder["dup"][2][:1023]
                e.name = nm
                e.name_len = len(nm)
            merged.reads_before_error[0] = best >> 6
        # median over the merged histogram (src/fastq_info.c:39-55)
        lo, hi = int(m0.min_rl), int(m0.max_rl)
        med = MAX_READ_LENGTH
        if m0.num_rds == 1:
            med = lo
        elif m0.num_rds > 1 and hi >= lo and hi < MAX_READ_LENGTH:
            h = torch.tensor(self.ctx.hist_range(0, lo, hi), dtype=torch.int64, device=self.tdev)
            if W > 1:
                dist.all_reduce(h)
            c = torch.cumsum(h, 0)
            idx = int(torch.nonzero(c > m0.num_rds // 2)[0]) if bool((c > m0.num_rds // 2).any()) else None
            med = lo + idx if idx is not None else MAX_READ_LENGTH
        merged.median_rl = med
        out = {"report": merged, "event_key": best, "n_records": int(m0.n_records), "n_index_entries": int(merged.n_index_entries)}
        if r == 0:
            out["transcript"] = self.ctx.render(merged, name, None, empty_ok=empty_ok, no_enc_ok=no_enc_ok)
        return out


def _as_tensor(ptr, n, device):
    """uint8 tensor view of raw device (or, in the CPU tests, host) memory."""
    if device.type == "cuda":
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return torch.as_tensor(h, device=device)
    buf = (ctypes.c_uint8 * n).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)

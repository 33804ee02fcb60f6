/*
 * fq_device.h — the operations the host engine (fq_engine.cpp) needs from the GPU.
 *
 * The product implementation is FqCudaDevice (fq_cuda.cu): hand-written sm_100a kernels on one CUDA stream.
 * tests/sim/ holds a sequential stand-in with the same interface so that the host logic (chunking, bridging,
 * event ordering, report) can be exercised on a machine without a GPU; it is never linked into libfastq_gpu.
 */
#ifndef FQ_DEVICE_H
#define FQ_DEVICE_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <new>
#include <stdexcept>
#include "fq_types.h"

/* where to find the bytes of a name given the index of its record inside a file: one entry per segment */
typedef struct {
  unsigned long long g0;     /* index of the segment's first record */
  const FqName* names;       /* descriptors of the segment's records */
  const uint8_t* data;       /* chunk the descriptors point into */
} FqDirEntry;

typedef struct {
  const uint8_t* data;       /* chunk bytes (allocation padded by ≥ 64 bytes) */
  const uint32_t* line_end;  /* exclusive end offset of every gz-line of the chunk */
  const FqLine* lines;       /* explicit lines, 4 per record (serially split records); NULL → derive from line_end */
  uint32_t q, j0;            /* the segment's first line starts at byte q and ends at line_end[j0] */
  uint32_t nrec;
  uint64_t span_bytes;       /* FASTQ bytes the records cover (kernel statistics only) */
  uint64_t g0;               /* index of the first record inside its file */
  uint64_t step_base;        /* steps preceding this file's loop (mate loop) */
  FqRecCtx cx;
  FqStats* stats;            /* counters to update (num_rds, n_names, mem_sum) */
  FqStats* stats_range;      /* min/max read length and quality to update: the mate loop's are never reported (fastq_info.c:316-320 snapshots file 1's before it) */
  unsigned long long* hist;  /* read-length histogram, FQ_MAX_READ_LENGTH bins */
  unsigned long long* key;   /* global minimum event key */
  FqName* names;             /* out: nrec descriptors (NULL when the loop has no name step) */
} FqRecordsArgs;

typedef struct {
  const FqName* names; const uint8_t* data; uint32_t nrec; uint64_t g0; uint64_t step_base;
  FqSlot* slots; uint64_t mask;            /* capacity-1 (power of two) */
  const FqDirEntry* dir1; uint32_t ndir1;  /* file-1 directory: resolves a stored idx1 to name bytes */
  unsigned long long* key;
  unsigned long long* counters;            /* [0] hash collisions seen, [1] slots claimed by the mate loop, [2] table full */
} FqTableArgs;

typedef struct {
  const FqName* a; const uint8_t* da; uint32_t stride_a;
  const FqName* b; const uint8_t* db; uint32_t stride_b;
  uint32_t npairs; uint64_t p0;            /* step of the first pair */
  uint32_t rank;                           /* FQ_RI_UNPAIRED or FQ_RS_MISMATCH */
  unsigned long long* key;
} FqPairArgs;

/* a name on its way to the owner of its hash (multi-GPU sharded index); same layout as fqg_packed_name */
typedef struct { unsigned long long hash; unsigned long long record; uint32_t off; uint32_t len; } FqPackedName;
#define FQ_SHARD_POS_BITS 28
#define FQ_SHARD_MAX_SRC 64
typedef struct {
  const FqPackedName* meta; unsigned long long n; const uint8_t* blob;
  uint32_t n_src; unsigned long long meta_start[FQ_SHARD_MAX_SRC + 1]; unsigned long long blob_start[FQ_SHARD_MAX_SRC];
  FqSlot* slots; unsigned long long mask;
  unsigned long long* dup_key;   /* min event key of a duplicate */
  unsigned long long* counters;  /* [0] collisions, [2] table full */
} FqShardArgs;
/* pipelined routing: a name on its way to the owner of its hash is a slot of 16 + 16 * units bytes: this header, then `units`
 * 16-byte units holding the name's bytes, zero padded (units = 0: the tuple travels alone and the owner cannot judge equal hashes).
 * A region holds the names one source has for one owner in one round, written by `nblocks` writers that do not talk to each other
 * (the CTAs of a clean-data pass; or one writer, the pack kernel): header {nblocks, stride, flags}, one count per writer, then
 * nblocks stretches of `stride` slots. */
typedef struct { unsigned long long hash; unsigned long long rec_len; /* record << 12 | length of the name */ } FqRouteSlot;
typedef struct { uint32_t nblocks, stride, flags, pad; } FqRegionHdr;
#define FQ_ROUTE_NAME_TOO_LONG 1u  /* region flag: a name needed more units than the slots have (it travelled cut short) */
#define FQ_ROUTE_MAX_WORLD 16      /* owners a clean-data pass can write to directly */
typedef struct { uint8_t* region[FQ_SHARD_MAX_SRC]; } FqRegionPtrs;
FQ_HD size_t fq_route_slot_bytes(uint32_t units) { return 16u + 16u * (size_t)units; }
FQ_HD size_t fq_route_counts_bytes(uint32_t nblocks) { return ((size_t)nblocks * 4u + 15u) & ~(size_t)15u; }
FQ_HD size_t fq_route_region_bytes(uint32_t nblocks, unsigned long long stride, uint32_t units) {
  return 16u + fq_route_counts_bytes(nblocks) + (size_t)nblocks * stride * fq_route_slot_bytes(units);
}
FQ_HD uint32_t fq_owner_of(uint64_t hash, uint32_t world) { return (uint32_t)((((hash >> 40) & 0xFFFFFFull) * world) >> 24); }

/* fused scan + validate pass over one chunk (FqCudaDevice only): the records of the segment that starts at line j0 */
typedef struct {
  const uint8_t* data; uint32_t n; int virtual_end; uint32_t* line_end; uint32_t cap;
  uint32_t* out5;            /* device: lines, cap overflow, first over-long header line (init ~0), records that did not fit, internal error */
  uint32_t j0; uint32_t max_rec; uint64_t g0; uint64_t step_base; FqRecCtx cx;
  FqStats* stats; FqStats* stats_range; unsigned long long* hist; unsigned long long* key; FqName* names; uint32_t names_cap;
  uint32_t hint_line_len;    /* length of the file's first sequence line (0 = unknown): picks the clean-data pass's mode */
  uint32_t lead;             /* clean-data pass only: the first `lead` (< 16) bytes of data[] are not the chunk's (offsets still count from data) */
  uint8_t* arena;            /* clean-data pass, per-line mode: block for the names' bytes (names[k].off then points in here); NULL: names stay chunk-relative */
  uint32_t arena_units;      /* its capacity in 16-byte units: a pass that needs more hands the chunk on (capacity anomaly) */
  /* clean-data pass, per-line mode, sharded runs: the pass writes every name straight into the region of the rank that owns its
   * hash (no name descriptors, no arena, no pack kernel): route_world owners, regions of lanes_max_blocks() stretches of
   * route_stride slots with route_units name units */
  uint32_t route_world, route_stride, route_units;
  uint8_t* route_region[FQ_ROUTE_MAX_WORLD];
} FqTileArgs;

#define FQ_LANES_OUT_WORDS 32 /* [16..23]: the last 8 line ends of the chunk (fewer when it has fewer lines); [24] arena units used, [25] the per-line
                               * pass accepted the chunk AND committed its statistics, [26] first line whose end the last tiles stored, [27] records staged */
class FqDevice {
 public:
  virtual ~FqDevice() {}
  virtual const char* name() const = 0;
  /* memory: device allocations are zero-padded by the callee's caller; all copies are ordered on one stream */
  virtual void* alloc(size_t n) = 0;
  virtual void release(void* p) = 0;
  virtual void upload(void* dst, const void* src, size_t n) = 0;
  virtual void download(void* dst, const void* src, size_t n) = 0; /* synchronises */
  virtual void copy(void* dst, const void* src, size_t n) = 0;
  virtual void fill(void* dst, int byte, size_t n) = 0;
  /* up to 32 words of device memory set from host values, ordered on the main stream */
  virtual void set_words(uint32_t* dst, const uint32_t* words, uint32_t nwords) { upload(dst, words, nwords * sizeof(uint32_t)); sync_main(); }
  /* fill ordered with the index kernels instead of the main stream's copies and passes (the stream of index_insert, mate_claim
   * and the shard kernels): clearing the table then runs beside the first chunk's pass */
  virtual void fill_index(void* dst, int byte, size_t n) { fill(dst, byte, n); }
  virtual void sync() = 0;
  /* page-locked host memory for pieces on their way to the device (a streaming caller's double buffer) */
  virtual void* host_alloc(size_t n) { void* p = malloc(n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
  virtual void host_release(void* p) { free(p); }
  /* wait for the copies and kernels queued on the main stream only (the index kernels on their own stream keep running) */
  virtual void sync_main() { sync(); }
  /* K1: exclusive end offsets of all lines of data[0,n) in order; a last line without LF counts when virtual_end.
   * out[0] = number of lines, out[1] = 1 if more than cap were found (only the first cap are stored). */
  virtual void scan_lines(const uint8_t* data, uint32_t n, int virtual_end, uint32_t* line_end, uint32_t cap, uint32_t* out2, uint32_t lead = 0) = 0; /* lead (< 16): bytes in front that are not the chunk's own */
  /* number of LF bytes in data[0,n): *out += count (64-bit) */
  virtual void count_lines(const uint8_t* data, uint32_t n, unsigned long long* out) = 0;
  /* first line i in [j0, j0+nlines) whose raw length reaches the gzgets limit of its phase ((i-j0)&3); *out = min(*out, i) */
  /* When tail_from_n != 0 the unterminated bytes after the last counted line (up to n) are judged as line j0+nlines. */
  virtual void find_overlong(const uint32_t* line_end, uint32_t q, uint32_t j0, uint32_t nlines, uint32_t n, int tail_from_n, uint32_t* out) = 0;
  /* gzgets emulation for one record starting at byte q: lines[4], out3 = {next byte, lines obtained (0..4), LFs consumed}.
   * A line cut short by the end of the data only counts when is_eof. */
  virtual void split_serial(const uint8_t* data, uint32_t n, uint32_t q, int is_eof, FqLine* lines4, uint32_t* out3) = 0;
  /* format / colour-space sniff on (hdr1, seq) of a file's first record: out2 = {FQ_SNIFF_*, is_color} */
  virtual void sniff(const uint8_t* data, FqLine hdr1, FqLine seq, int32_t* out2) = 0;
  /* K2: reader flags + validation + statistics + event key + name descriptors */
  virtual void records(const FqRecordsArgs& a) = 0;
  /* K1+K2 in one pass over HBM; false when the device has no such kernel (the engine then uses K1 and K2) */
  virtual bool tile_pass(const FqTileArgs& a) = 0;
  /* K5: the clean-data pass (FqCudaDevice only; false = no such kernel).  Proves the chunk clean and stages line index, name
   * descriptors and statistics; a.out5 holds FQ_LANES_OUT_WORDS words: [0] lines, [1] cap overflow, [2] first over-long header
   * line, [3] anomaly bits of the pass, [4] internal error, [5] index of a final line without LF, [6..9] staged min/max of
   * quality and read length, [10] a record broke a length rule.  The counters / histogram are committed by the pass itself
   * when [1..4] are clear; lanes_commit(undo=false) folds the staged min/max in, lanes_commit(undo=true) takes the counters back. */
  /* *self_judged: the per-line mode ran — records, statistics and the commit are the pass's own business: out5[25] says whether the
   * chunk was accepted (then everything is in place, names in a.arena), no lanes_commit in either case */
  virtual bool lanes_pass(const FqTileArgs& a, bool* self_judged) { (void)a; (void)self_judged; return false; }
  virtual void lanes_commit(const FqTileArgs& a, bool undo) { (void)a; (void)undo; }
  /* K3: index insert (file 1) / K4: mate claim (file 2) / pair compare (interleaved, sorted) */
  virtual void index_insert(const FqTableArgs& a) = 0;
  virtual void mate_claim(const FqTableArgs& a) = 0;
  virtual void pair_compare(const FqPairArgs& a) = 0;
  /* multi-GPU: per-owner counts (out[2*o] names, out[2*o+1] bytes), packing, owner-side insert, lookup of a record's tuple */
  virtual void names_count(const FqName* names, uint32_t nrec, uint32_t world, unsigned long long* out) = 0;
  virtual void names_pack(const FqName* names, const uint8_t* data, uint32_t nrec, uint64_t g0, uint32_t world, FqPackedName* meta,
                          uint8_t* blob, const unsigned long long* base, unsigned long long* cursor) = 0;
  virtual void shard_insert(const FqShardArgs& a) = 0;
  /* owner side of the mate loop: tuples of file 2 claim the slots filled from `inserted` (file 1's tuples); unpaired events
   * go to a.dup_key as FQ_KEY(step_base + record, FQ_R_NAME), first claims are counted in a.counters[1] */
  virtual void shard_claim(const FqShardArgs& a, const FqShardArgs& inserted, unsigned long long step_base) = 0;
  /* pipelined routing (include/fastq_gpu.h: fqg_names_pack_slots / fqg_shard_insert_slots).  These run on the side stream so
   * that they can be called while a clean-data pass occupies the main stream (beside = that pass has been launched and the names
   * to pack were complete before it).  route_begin clears cursors[world]; names_pack_slots appends one segment's tuples to the
   * owners' regions (tuples beyond cap are counted, not stored); route_end writes the counts into the region headers and
   * returns when everything is in place. */
  virtual void route_begin(unsigned long long* cursors, uint32_t world, bool beside) = 0; /* cursors: 2 * FQ_SHARD_MAX_SRC words (counts, then flags) */
  virtual void names_pack_slots(const FqName* names, const uint8_t* arena, uint32_t nrec, uint64_t g0, uint32_t world, const FqRegionPtrs& R, uint64_t cap,
                                uint32_t units, unsigned long long* cursors) = 0;
  virtual void route_end(const unsigned long long* cursors, uint32_t world, const FqRegionPtrs& R, uint64_t cap) = 0;
  /* slots of n_src regions (region_bytes apart, each planned for nblocks stretches of `stride` slots; a header with nblocks = 0 is an
   * empty region) into the table: counters[1] += inserted, counters[0] += names that are in the table already (units > 0: hash AND
   * bytes equal — a duplicated read name; units = 0: equal hashes, which tuples alone cannot judge), counters[2] = 1 when a count
   * exceeds its stretch, a name did not fit its slot, a header contradicts the plan, or the table is full.  A slot whose hash is
   * taken by ANOTHER name walks on (units > 0).  Asynchronous. */
  virtual void shard_insert_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units,
                                  FqSlot* slots, unsigned long long mask, unsigned long long* counters, bool beside,
                                  const unsigned long long* flags, unsigned long long expect) = 0; /* flags != NULL: wait on the device until flags[s] >= expect for every source s */
  /* the mate loop at the owner (src/fastq_info.c:333-350: lookup, then delete): every slot (units > 0) looks its name up by hash and
   * bytes; the first one to find it claims it (counters[8]++), a name that is not there or was claimed before counts in counters[9] */
  virtual void shard_claim_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units,
                                 FqSlot* slots, unsigned long long mask, unsigned long long* counters, bool beside,
                                 const unsigned long long* flags, unsigned long long expect) = 0;
  /* the number of writers (CTAs) a clean-data pass in per-line mode uses at most: the nblocks of the regions it routes into */
  virtual uint32_t lanes_max_blocks() { return 0; }
  /* the main stream's next clean-data pass waits for what was queued on the side stream so far (copies out of the regions that
   * pass will overwrite) */
  virtual void side_mark() {}
  /* what this device queues from now on (side or main stream) starts after what `earlier` has queued so far */
  virtual void order_after(bool /*my_side*/, FqDevice& /*earlier*/, bool /*their_side*/) {}
  virtual void side_copy(void* dst, const void* src, size_t n) = 0;
  virtual void side_sync() = 0;
  virtual void side_copy_lane(int /*lane*/, void* dst, const void* src, size_t n) { side_copy(dst, src, n); }
  /* memory that other processes can map (CUDA IPC); devices without it throw */
  virtual void* ipc_alloc(size_t n, uint8_t handle[64]) { (void)n; (void)handle; throw std::runtime_error("this device has no inter-process memory"); }
  virtual void* ipc_open(const uint8_t handle[64]) { (void)handle; throw std::runtime_error("this device has no inter-process memory"); }
  virtual void ipc_close(void* p) { (void)p; }
  virtual void ipc_free(void* p) { (void)p; }
  virtual void shard_find(const FqPackedName* meta, unsigned long long n, unsigned long long record, unsigned long long* out_pos) = 0;
  /* The name arena (what the reference's new_indexentry keeps of a record, src/fastq.c:590-611: a copy of its name).  Names leave
   * their chunk so that the chunk's bytes can be released once its records are final: names_measure adds the 16-byte units the
   * names of a segment take (*out_units += ...; records without a name step, hash == FQ_HASH_SKIP, take none), names_gather copies
   * every such name into `arena` (one zero-padded run of 16-byte units per name, in any order; *cursor_units counts the units
   * handed out and starts at 0) and rewrites names[k].off to the byte offset of the copy inside `arena`. */
  virtual void names_measure(const FqName* names, uint32_t nrec, unsigned long long* out_units) = 0;
  virtual void names_gather(FqName* names, const uint8_t* data, uint32_t nrec, uint8_t* arena, unsigned long long* cursor_units) = 0;
  /* Statistics of records that are not final yet (a later event may still restrict the file) are kept in a second set, `open`;
   * stats_fold adds the open set of both files into the main set (counters, minima / maxima, the histogram bins between the open
   * minima and maxima of either file) and leaves the open set empty. */
  virtual void stats_fold(FqStats* const main2[2], FqStats* const open2[2], unsigned long long* const hist2[2], unsigned long long* const hist_open2[2]) = 0;
  /* fastq_filter_n's predicate (src/fastq_filter_n.c:77-86) for n sequence lines of a chunk: out2[2k] = 'N' / 'n' bytes before the first LF or
   * NUL of line k, out2[2k+1] = strlen of the line (bytes before the first NUL, terminator included) */
  virtual void count_n(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out2) = 0;
  /* fastq_filterpair: the names of `n` header lines as descriptors (fq_header_name); and for every descriptor the record index the
   * table holds for an equal name (FQ_IDX_NONE: none), names compared byte by byte against the indexed names (a.dir1) */
  virtual void header_names(const uint8_t* data, const FqLine* hdr_lines, uint32_t n, int fmt, int is_pe, uint32_t seed, FqName* out) = 0;
  virtual void names_lookup(const FqTableArgs& a, unsigned long long* out_idx) = 0;
  /* fastq_trim_poly_at: out3[3k..] = read_len, poly-A/N bytes at the line's end, poly-T/N bytes at its start (fq_poly_at) */
  virtual void poly_at(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out3) = 0;
  /* details of one record for the error message */
  virtual void explain(const uint8_t* data, const FqLine* lines4_host, const FqRecCtx& cx, FqRecOut* out_dev) = 0;
  /* device-side stopwatch on the stream (CUDA events) */
  virtual void timer_start() = 0;
  virtual double timer_stop_ms() = 0;
  /* per-kernel-class device time: which = FQG_K_*; returns false for an unknown class */
  virtual bool kernel_stat(int which, double* ms, uint64_t* launches, uint64_t* bytes, uint64_t* items) = 0;
  virtual void kernel_stats_reset() = 0;
  /* how many kernels were launched so far (bench.py's gpu_launches) */
  virtual unsigned long long launches() const = 0;
};

FqDevice* fq_make_cuda_device(int ordinal); /* fq_cuda.cu; throws std::runtime_error when no CUDA device */

#endif

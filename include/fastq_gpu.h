/*
 * fastq_gpu.h — C ABI of libfastq_gpu: the fastq_info hot path (record split, validation, read-name
 * uniqueness, mate matching) of nunofonseca/fastq_utils 0.25.3 on one B200.
 *
 * The reference has no FFI; its boundary for this path is the C API of src/fastq.h:133-158 + src/hash.h:64-78
 * as driven by the four loops of src/fastq_info.c.  A record-at-a-time call cannot be accelerated, so the
 * library replaces the LOOP level and leaves argv parsing, zlib and exit() to the caller:
 *
 *   fqg_create / fqg_feed / fqg_finish   replace   fastq_new + the loop + the statistics it leaves behind
 *        FQG_MODE_SINGLE       validate_single_fastq_file           src/fastq_info.c:155-176   (-r file)
 *        FQG_MODE_INDEX        fastq_index_readnames                src/fastq.c:396-439        (file)
 *        FQG_MODE_INDEX_PAIR   ... followed by main()'s mate loop   src/fastq_info.c:322-362   (file1 file2)
 *        FQG_MODE_INTERLEAVED  validate_interleaved                 src/fastq_info.c:57-106    (file pe)
 *        FQG_MODE_SORTED_PAIR  validate_paired_sorted_fastq_file    src/fastq_info.c:108-152   (-r -s file1 file2)
 *   fqg_render                  replaces   the fprintf/PRINT_ERROR/exit sequence of src/fastq_info.c:190-396
 *   fqg_fastq_info_mem          = main() on already-inflated streams (what the CLI and the parity tests call)
 *   fqg_index_records           exposes the record index for reader-style tools (src/fastq_truncate.c)
 *   fqg_reader_tool_mem         = main() of fastq_num_reads / fastq_not_empty on an inflated stream (FQG_MODE_READER)
 *
 * Nothing here prints or exits.  All functions return 0 on success or a negative FQG_ERR_* (the caller maps
 * these to the reference's SYS_INT_ERROR_EXIT_STATUS = 2, src/fastq.h:79).  A context is used by one host
 * thread.  There is no CPU fallback: without a CUDA device fqg_create fails with FQG_ERR_NO_DEVICE.
 */
#ifndef FASTQ_GPU_H
#define FASTQ_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQG_VERSION "0.25.3-b200.1"

enum { FQG_MODE_SINGLE = 0, FQG_MODE_INDEX = 1, FQG_MODE_INDEX_PAIR = 2, FQG_MODE_INTERLEAVED = 3, FQG_MODE_SORTED_PAIR = 4,
       FQG_MODE_READER = 5 /* fastq_read_next_entry until the end of one file, nothing validated: src/fastq_num_reads.c:43-45 */ };
enum { FQG_ERR_NO_DEVICE = -1, FQG_ERR_CUDA = -2, FQG_ERR_OOM = -3, FQG_ERR_USAGE = -4, FQG_ERR_INTERNAL = -5 };

typedef struct fqg_ctx fqg_ctx;

typedef struct {
  int32_t mode;                  /* FQG_MODE_* */
  int32_t device;                /* CUDA ordinal */
  uint64_t index_capacity_hint;  /* expected number of read names (0 = grow on demand) */
  uint32_t flags;                /* FQG_FLAG_* */
  uint32_t reserved[3];
} fqg_config;
/* FQG_MODE_INDEX only: a second file operand exists but is never read (`-s f1 f2` without -r): the reference still
 * strips the mate digit from default-format names (is_pe, src/fastq_info.c:290) */
#define FQG_FLAG_PAIRED_NAMES 1u
/* the read-name index lives elsewhere (sharded over GPUs): names are exported with fqg_names_* instead of inserted */
#define FQG_FLAG_EXTERNAL_INDEX 2u
/* never use the fused single-pass kernel (A/B measurements; results are identical) */
#define FQG_FLAG_TWO_PASS 4u
/* fqg_feed_device: the caller's buffer is borrowed for the duration of the call only (a streaming caller that refills it).
 * Without the flag the buffer must stay valid until fqg_reset / fqg_destroy: chunks of a job in which an event fired, and all chunks
 * of the interleaved / sorted-pair loops, are read again by fqg_finish. */
#define FQG_FLAG_BORROW_FOR_CALL 8u
/* keep every chunk until fqg_reset (the record-writing reader tools read the line index of the whole stream afterwards) */
#define FQG_FLAG_KEEP_CHUNKS 16u

/* statistics of one FASTQ_FILE as the reference holds them at the end of its loops (src/fastq.h:110-131) */
typedef struct {
  uint64_t n_records;   /* records read by fastq_read_entry                                   */
  uint64_t num_rds;     /* the reference's counter (index loop counts twice, fastq.c:241,344) */
  uint64_t min_rl, max_rl;       /* terminators included; 2500000 / 0 when nothing was read   */
  uint64_t min_qual, max_qual;   /* sign-extension quirk applied (fastq.c:374)                */
  int32_t sniff_format;          /* -1 not sniffed, 0 default, 1 casava 1.8, 2 integer, 3 no suffix */
  int32_t color_space;           /* -1 not sniffed, 0 / 1                                     */
} fqg_file_report;

/* codes: SURVEY.md appendix B */
enum {
  FQG_OK = 0, FQG_E_TRUNC, FQG_E_TRUNC_PE, FQG_E_WRONGHDR, FQG_E_AT, FQG_E_IDLEN, FQG_E_BADCHAR, FQG_E_UT, FQG_E_SHORT,
  FQG_E_PLUS, FQG_E_HDR2, FQG_E_LEN, FQG_E_LEN_CS, FQG_E_DUP, FQG_E_UNPAIRED, FQG_E_LEFTOVER, FQG_E_MISMATCH,
  FQG_E_EOF1, FQG_E_EOF2, FQG_E_EMPTY, FQG_E_ENC
};

typedef struct {
  int32_t code;          /* FQG_E_*; FQG_OK when every record passed                              */
  int32_t file;          /* file whose loop raised it (0/1)                                       */
  int32_t msg_file;      /* file NAME the reference prints (mate loop validates against file 1)   */
  int32_t chr;           /* FQG_E_BADCHAR: the byte                                               */
  uint64_t record;       /* index of the record inside `file`                                     */
  uint64_t line;         /* the line number the reference prints                                  */
  uint64_t a, b;         /* slen / qlen, read number, count of unpaired reads                      */
  uint64_t event_key;    /* position of the event in the reference's sequential order (smaller = earlier); ~0 when none */
  uint32_t hdr1_len, hdr2_len, name_len;
  char hdr1[1024];       /* raw lines as C strings, for the messages that quote them              */
  char hdr2[1024];
  char name[1024];
} fqg_error;

typedef struct {
  int32_t mode;
  int32_t reserved;
  fqg_file_report file[2];
  uint64_t n_index_entries;      /* index->n_entries after the index loop                         */
  uint64_t n_index_left;         /* ... after the mate loop                                       */
  uint64_t index_mem;            /* bytes, as accumulated by fastq.c:609 (+8, fastq_info.c:293)   */
  uint64_t median_rl;            /* median_rl(), fastq_info.c:39-55 (terminator included)         */
  uint64_t reads_before_error[2];/* records fully processed per file before the first error (progress lines) */
  fqg_error error;
} fqg_report;

int fqg_create(const fqg_config* cfg, fqg_ctx** out);
void fqg_destroy(fqg_ctx* ctx);
/* Hand `n` decompressed bytes of file `file` (0 or 1) to the device.  `last` marks the end of that file.
 * INDEX_PAIR: file 1 may only be fed after file 0 was fed with last=1 (the reference's order). */
int fqg_feed(fqg_ctx* ctx, int file, const void* host_bytes, size_t n, int last);
/* Same, for bytes already in device memory.  The buffer must be followed by at least 64 readable bytes, and — the kernels read 16 bytes
 * at a time — the bytes between the 16-byte boundary below `device_bytes` and the pointer must be readable too (true inside any
 * allocation; a caller that hands over the very first byte of a mapped range aligns it).  It is borrowed until fqg_reset /
 * fqg_destroy, or only for the duration of the call with FQG_FLAG_BORROW_FOR_CALL.  Device memory the job keeps: the read names (a
 * copy of each, like the reference's new_indexentry) and the index; chunks whose records raised no event are dropped as the job goes,
 * so files far larger than the device's memory stream through (single-file and default two-file modes). */
int fqg_feed_device(fqg_ctx* ctx, int file, const void* device_bytes, size_t n, int last);
int fqg_finish(fqg_ctx* ctx, fqg_report* out);
/* forget all input and results, keep the device workspace (bench loops) */
int fqg_reset(fqg_ctx* ctx);
const char* fqg_last_error(const fqg_ctx* ctx);
/* kernels launched by this context so far */
uint64_t fqg_launch_count(const fqg_ctx* ctx);
/* device time (ms, CUDA events on the context's stream) spent between the first feed and the end of finish */
double fqg_device_ms(const fqg_ctx* ctx);

/* Record index of a device- or host-resident stream: writes up to cap byte offsets of record starts, returns
 * the number of complete records in *n_records (fastq_num_reads / fastq_truncate style consumers). */
int fqg_index_records(fqg_ctx* ctx, const void* host_bytes, size_t n, uint64_t* starts, size_t cap, uint64_t* n_records);

/* ---- text: the reference's stdout / stderr for a finished run ---- */
typedef struct {
  int32_t empty_ok;      /* -e */
  int32_t no_enc_ok;     /* -q */
  const char* name1;     /* file operands as given on the command line */
  const char* name2;
} fqg_render_opts;
typedef struct { int32_t rc; char* out; size_t out_len; char* err; size_t err_len; } fqg_transcript;
int fqg_render(const fqg_report* rep, const fqg_render_opts* opts, fqg_transcript* t);
void fqg_transcript_free(fqg_transcript* t);

/* fastq_info's main() on inflated streams: argv as the reference receives it; n = (size_t)-1 means "could not
 * be opened" (src/fastq.c:651-655).  chunk_bytes > 0 feeds the streams in pieces of that size. */
int fqg_fastq_info_mem(int argc, const char** argv, const void* f1, size_t n1, const void* f2, size_t n2,
                       int device, size_t chunk_bytes, fqg_transcript* t);

/* The same for streams the caller reads piece by piece (the CLI: zlib inflation of files of any size, `-` for standard input).  The
 * library decides which words of argv are file operands exactly as the reference's main() does (its getopt quirks included) and calls
 * `open` for those, in the reference's order — file 2 of the default two-file mode only after file 1 was indexed without error.  `read`
 * delivers the next inflated bytes (0: end of the stream, < 0: error) and is called from a helper thread of the library, one call at a
 * time per handle: while it fills one page-locked piece, the other is copied to the device and validated.  Host memory: two pieces of
 * piece_bytes (0 = 64 MiB); device memory: the read names and the index, whatever the size of the files. */
typedef struct {
  void* user;
  void* (*open)(void* user, const char* name);   /* NULL result: "Unable to open <name>" */
  long (*read)(void* user, void* handle, void* buf, size_t cap);
  void (*close)(void* user, void* handle);
} fqg_stream_io;
int fqg_fastq_info_stream(int argc, const char** argv, const fqg_stream_io* io, int device, size_t piece_bytes, fqg_transcript* t);

/* The reader-style tools on an already-inflated stream: argv[0] names the tool,
 *   "fastq_num_reads"  src/fastq_num_reads.c:32-50   prints the number of entries fastq_read_next_entry delivers
 *   "fastq_not_empty"  src/fastq_not_empty.c:32-47   exit status 0 when the file holds at least one entry, 1 otherwise
 * with the reference's usage text, "file truncated" / "Unable to open" errors and exit statuses.  n = (size_t)-1: could not be opened. */
int fqg_reader_tool_mem(int argc, const char** argv, const void* f1, size_t n1, int device, size_t chunk_bytes, fqg_transcript* t);

/* fastq_filterpair (src/fastq_filterpair.c:38-240) on two inflated streams below 2 GiB each: argv as the reference
 * receives it (fastq1 fastq2 paired1 paired2 unpaired [sorted]); n = (size_t)-1: the file could not be opened.  File 1 (with "sorted":
 * both files) goes through the index loop of fastq_info — validation and duplicate check included — the mates are found through the
 * index on the device.  outs[0..2] / lens[0..2]: what the reference gzips into paired1, paired2 and unpaired, inflated (release each
 * with fqg_buffer_free); *created != 0 when the run got as far as creating the three files.  After an error (exit status 1 or 3) the
 * reference leaves unfinished gzip files behind: the buffers then hold the records written so far. */
int fqg_filterpair_mem(int argc, const char** argv, const void* f1, size_t n1, const void* f2, size_t n2, int device,
                       fqg_transcript* t, char* outs[3], size_t lens[3], int32_t* created);

/* fastq_trim_poly_at (src/fastq_trim_poly_at.c:121-233): argv as the reference receives it (--file, --outfile,
 * --min_poly_at_len, --min_len, --help, parsed by the C library's getopt_long like the reference does).  The library opens --file
 * through `io` (any size: the stream is taken in windows that start at record starts), delimits the records and scans their poly-A / poly-T ends on the device, and returns what the
 * reference would gzip into --outfile, inflated, in *outfile (release with fqg_buffer_free); *outfile_name is the argv word naming that
 * file, NULL when the run ended before the reference creates it (usage errors, --help, an input that cannot be opened).  After a
 * "file truncated" error (exit status 1) the reference leaves an unfinished gzip file behind: *outfile then holds the records written so far. */
int fqg_trim_poly_at_stream(int argc, const char** argv, const fqg_stream_io* io, int device, fqg_transcript* t,
                            char** outfile, size_t* outfile_len, const char** outfile_name);
void fqg_buffer_free(void* p);

/* ---- multi-GPU building blocks (fastq_utils_b200/dist.py drives them with torch.distributed) ----
 * A rank holds a contiguous byte range of a file.  fqg_prescan_device builds the line index of the range (kept for the
 * following fqg_feed_device of the same pointer) and reports what the ranks exchange to fix each range's line phase. */
int fqg_prescan_device(fqg_ctx* ctx, int file, const void* device_bytes, size_t n, int at_eof,
                       uint64_t* n_lines, int32_t* ends_with_lf, uint64_t first_line_ends[4]);
/* the first record of this context's stream starts after `skip_lines` lines of the first fed buffer and is record
 * number `first_record` of the whole file (event keys, line numbers and name indices become global).  A record offset is taken in
 * FQG_MODE_SINGLE, in the index modes with FQG_FLAG_EXTERNAL_INDEX, and — even offsets only: a range starts with the first mate of a
 * pair — in FQG_MODE_INTERLEAVED. */
int fqg_set_stream_start(fqg_ctx* ctx, int file, uint32_t skip_lines, uint64_t first_record);
/* names of all records fed so far, routed by hash to `world` owners: sizes, then the packed tuples
 * (24-byte fqg_packed_name grouped by owner + the name bytes grouped by owner) into caller-provided device memory.
 * device_blob may be NULL: then only the tuples are packed; fqg_shard_insert with a NULL blob counts every equal hash as a
 * hash collision instead of judging it, and the caller repeats the exchange with the name bytes when any owner reports one
 * (a duplicate read name or a 64-bit collision: both rare) */
typedef struct { uint64_t hash; uint64_t record; uint32_t off; uint32_t len; } fqg_packed_name;
int fqg_names_count(fqg_ctx* ctx, int file, uint32_t world, uint64_t* counts, uint64_t* bytes);
int fqg_names_pack(fqg_ctx* ctx, int file, uint32_t world, void* device_meta, void* device_blob,
                   const uint64_t* meta_base, const uint64_t* blob_base);
/* owner side: insert received tuples (n_src groups, group s = metas [meta_start[s], meta_start[s+1]) whose `off` is
 * relative to blob_start[s]) into this context's index shard; the buffers must stay valid until fqg_shard_result */
int fqg_shard_insert(fqg_ctx* ctx, const void* device_meta, uint64_t n, const void* device_blob, uint32_t n_src,
                     const uint64_t* meta_start, const uint64_t* blob_start);
/* earliest duplicate this shard saw: event key (~0 = none), the later record's index and the name */
int fqg_shard_result(fqg_ctx* ctx, uint64_t* event_key, uint64_t* record, char name[1024], uint32_t* name_len, uint64_t* hash_collisions);
/* mate loop at the owner: tuples of file 2 claim the names inserted by fqg_shard_insert (lookup-then-delete of
 * src/fastq_info.c:333-350); step_base = records of file 1 + 1.  Result: earliest "unpaired read" event, the record's index
 * inside file 2, its name, and how many index entries were claimed (the rest are the "found N unpaired reads"). */
int fqg_shard_claim(fqg_ctx* ctx, const void* device_meta, uint64_t n, const void* device_blob, uint32_t n_src,
                    const uint64_t* meta_start, const uint64_t* blob_start, uint64_t step_base);
int fqg_shard_claim_result(fqg_ctx* ctx, uint64_t* event_key, uint64_t* record, char name[1024], uint32_t* name_len,
                           uint64_t* claimed, uint64_t* hash_collisions);
/* sharded runs: the read-name format / colour space of a file come from ITS first record (src/fastq.c:459-485), which only
 * one rank holds: that rank sniffs (fqg_sniff_device), everyone sets the result before feeding */
int fqg_sniff_device(fqg_ctx* ctx, int file, const void* device_bytes, size_t n, uint32_t skip_lines, int32_t* sniff_format, int32_t* color_space);
int fqg_set_sniff(fqg_ctx* ctx, int file, int32_t sniff_format, int32_t color_space);
/* raw length of a typical sequence line of the file (a rank that does not hold the file's first record cannot see it): picks the
 * mode of the clean-data pass, one thread per line for short lines; 0 = unknown.  A wrong hint costs time, never correctness. */
int fqg_set_line_hint(fqg_ctx* ctx, int file, uint32_t seq_line_len);
/* seed of the read-name hash for the job that starts now (after fqg_reset, before the first feed; fqg_reset goes back to 0).  The
 * hash is unobservable behind the exact compare; a sharded run whose owners met two different names with one 64-bit hash repeats
 * the job with the next seed, as the one-GPU engine does by itself */
int fqg_set_hash_seed(fqg_ctx* ctx, uint32_t seed);
/* records delimited in what was fed so far, whether or not the loop read them all: fqg_finish reports fewer (n_records) when a
 * NUL-led header line ended the file early (src/fastq.c:248) — a sharded run must then forget the ranges behind that rank's */
int fqg_records_fed(fqg_ctx* ctx, int file, uint64_t* n_records);
/* records of file 0 over all ranks: the mate loop's steps and line numbers continue after them */
int fqg_set_file_total(fqg_ctx* ctx, int file, uint64_t total_records);
/* bins [lo, hi] of a file's read-length histogram (terminator included, like the reference's rdlen_ctr) */
int fqg_hist_range(fqg_ctx* ctx, int file, uint64_t lo, uint64_t hi, uint64_t* out);

/* ---- pipelined routing (sharded index runs, one file, tuples only; dist.py `_route_round`) ----
 * Instead of one exchange after the whole range has been validated, the names travel chunk by chunk while the next chunk's
 * clean-data pass runs.  The hook is called once per chunk from inside fqg_feed_device: for a chunk that takes the clean-data
 * pass right after that pass has been LAUNCHED (the GPU is busy with it; the names of the chunks before it are complete), for
 * any other chunk after it has been validated.  The callee packs the names that were not packed yet
 * (fqg_names_pack_slots, on the context's side stream), moves them to their owners and hands finished rounds to the owner's
 * index (fqg_shard_insert_slots).  Whatever is left after the last feed is routed by the caller. */
typedef void (*fqg_chunk_hook)(void* user, int file);
int fqg_set_chunk_hook(fqg_ctx* ctx, fqg_chunk_hook hook, void* user);
/* records whose names were not packed by fqg_names_pack_slots yet */
int fqg_names_new(fqg_ctx* ctx, int file, uint64_t* n_new);
/* A name travels as a slot of 16 + 16 * name_units bytes: {hash, record << 12 | length} and name_units 16-byte units of its bytes,
 * zero padded (name_units = 0: the tuple alone — enough for a one-file job, whose owner only has to notice equal hashes; a two-file
 * job needs the bytes, the mate loop compares every name).  A region holds what one source has for one owner in one round, written
 * by nblocks writers that do not talk to each other: a 16-byte header {uint32 nblocks, stride, flags, 0}, nblocks uint32 counts
 * (padded to 16 bytes), then nblocks stretches of `stride` slots.  A count above `stride` says a stretch overflowed (the surplus is
 * dropped), flag 1 that a name was longer than its slot: the owner reports either and the caller repeats the job through the
 * exact path.  nblocks = 0: the source had nothing in this round. */
size_t fqg_route_region_bytes(uint32_t nblocks, uint64_t stride, uint32_t name_units);
/* (1) The clean-data pass writes the names itself, from shared memory, while it validates: chunk n of `file` (counting the chunks the
 * pass accepted) goes to region_ptrs[o] + (n % depth) * region_bytes for owner o, laid out for fqg_route_blocks() writers of `stride`
 * slots.  No name descriptors, no arena, no pack kernel on this path.  fqg_route_chunks says how many chunks are complete in their
 * regions, and whether a chunk went another way (*broken: an anomaly handed it to the per-record kernels — the caller repeats the
 * job).  The caller moves the regions to their owners (fqg_side_copy) and calls fqg_side_mark so that the pass that reuses a region
 * waits for the copies out of it.  world = 0 switches the routing off. */
int fqg_set_route(fqg_ctx* ctx, int file, uint32_t world, void* const* region_ptrs, size_t region_bytes, uint32_t depth, uint32_t stride, uint32_t name_units);
int fqg_route_chunks(fqg_ctx* ctx, int file, uint64_t* n_chunks, int32_t* broken);
int fqg_route_blocks(fqg_ctx* ctx, uint32_t* nblocks);
int fqg_side_mark(fqg_ctx* ctx);
/* Stream order between two contexts of one process and device (the feeding context and the context that holds the index shard):
 * what `later` queues from now on, on its side stream (later_side != 0) or its main stream, starts after everything `earlier` has
 * queued so far on its side / main stream.  The owner's kernel of a round takes the whole device between two passes with this (a
 * table kernel squeezed in beside a running pass was measured 4x slower, and slowed the pass as well). */
int fqg_order_after(fqg_ctx* later, int later_side, fqg_ctx* earlier, int earlier_side);
/* (2) Records validated by the per-record kernels (the few at the seams of byte ranges; everything on the stand-in device) have name
 * descriptors: this packs those not packed yet by owner into `world` dense regions (one writer: nblocks = 1, stride = region_cap).
 * Region o starts at region_ptrs[o] (device memory, local or a peer's mapped with fqg_ipc_open).  Returns when they are complete. */
int fqg_names_pack_slots(fqg_ctx* ctx, int file, uint32_t world, void* const* region_ptrs, uint64_t region_cap, uint32_t name_units);
/* owner side: room for n_names in the index shard before the first fqg_shard_insert_slots (the table cannot grow between rounds) */
int fqg_shard_reserve(fqg_ctx* ctx, uint64_t n_names);
/* inserts the slots of n_src regions (region_bytes apart, planned for nblocks writers of `stride` slots); asynchronous: the regions
 * must stay valid until fqg_shard_slots_result — the index points at the names inside them.  beside != 0: one block per SM, so that
 * the kernel fits next to a running clean-data pass. */
int fqg_shard_insert_slots(fqg_ctx* ctx, const void* device_regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t name_units, int beside,
                           const void* device_flags, uint64_t expect);
/* device_flags != NULL: n_src 64-bit words; the kernel itself waits until word s is >= expect before it reads source s's region.  The
 * sources write their word behind the region's bytes (fqg_side_copy, same stream): no host takes part in a routing round.  A source
 * that stays silent for about ten seconds is given up (overflow is reported: the caller repeats the job). */
/* the mate loop at the owner (src/fastq_info.c:333-350): the slots of file 2's names (name_units > 0) look their name up by hash and
 * bytes and claim it (the reference's lookup-then-delete) */
int fqg_shard_claim_slots(fqg_ctx* ctx, const void* device_regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t name_units, int beside,
                          const void* device_flags, uint64_t expect);
/* waits for the inserts and claims: names inserted; names that were in the index already (name_units = 0: equal hashes, which
 * tuples alone cannot tell from a duplicate); whether a region, a slot or the table overflowed; names claimed by mates; mates that
 * found no name, or one that had been claimed before.  Anything but inserted == names of file 1, claimed == inserted == mates
 * sends the job through the exact path, which reports what the reference reports. */
int fqg_shard_slots_result(fqg_ctx* ctx, uint64_t* inserted, uint64_t* equal_hashes, int32_t* overflow, uint64_t* claimed, uint64_t* unpaired);
/* device memory other processes of this node can map (CUDA IPC): the owner's regions written by its peers' pack kernels
 * over NVLink instead of an all-to-all.  fqg_ipc_alloc returns the pointer and a 64-byte handle; fqg_ipc_open maps a peer's. */
/* device-to-device copy on the context's side stream (the copy engines move a packed region into a peer's arena while the SMs
 * stay with the clean-data pass), and the wait for it */
int fqg_side_copy(fqg_ctx* ctx, void* device_dst, const void* device_src, size_t bytes);
int fqg_side_sync(fqg_ctx* ctx);
/* the same on one of a few more copy streams (lane 1..7; lane 0 is the side stream itself): the regions for different owners travel
 * side by side on several copy engines instead of one after the other.  A lane keeps the order of its own copies (a region, then its
 * flag word); fqg_side_mark and fqg_side_sync cover every lane, and fqg_side_mark makes the side stream itself wait for the lanes. */
int fqg_side_copy_lane(fqg_ctx* ctx, int lane, void* device_dst, const void* device_src, size_t bytes);
int fqg_ipc_alloc(fqg_ctx* ctx, size_t bytes, void** device_ptr, uint8_t handle[64]);
int fqg_ipc_open(fqg_ctx* ctx, const uint8_t handle[64], void** device_ptr);
int fqg_ipc_close(fqg_ctx* ctx, void* device_ptr);
int fqg_ipc_free(fqg_ctx* ctx, void* device_ptr);

/* ---- per-kernel device timing (CUDA events around every launch on the context's stream) ---- */
enum { FQG_K_SCAN = 0, FQG_K_RECORDS = 1, FQG_K_INDEX = 2, FQG_K_MATE = 3, FQG_K_PAIR = 4, FQG_K_OTHER = 5, FQG_K_TILE = 6 /* fused scan+records, one thread per record */,
       FQG_K_LANES = 7 /* clean-data pass, chunk-parallel */, FQG_K_COUNT = 8 };
typedef struct { double ms; uint64_t launches; uint64_t bytes; uint64_t items; } fqg_kernel_stat;
/* accumulated since fqg_create / the last fqg_kernel_stats_reset; `bytes` = FASTQ bytes the launches covered */
int fqg_kernel_stats(fqg_ctx* ctx, int which, fqg_kernel_stat* out);
int fqg_kernel_stats_reset(fqg_ctx* ctx);

/* which path validated the chunks fed so far (since fqg_create / fqg_reset): out[0] chunks accepted by the clean-data pass,
 * out[1] chunks it handed on (an anomaly: the per-record kernels decided), out[2] chunks validated by the fused per-record
 * kernel, out[3] times the job fell back to the two-pass kernels */
int fqg_path_counts(fqg_ctx* ctx, uint64_t out[4]);

/* what the job holds in device memory (since fqg_create / fqg_reset): out[0] bytes of input chunks still held (owned or borrowed),
 * out[1] bytes of input chunks released because all their records are final, out[2] bytes of the read-name arena (the copy of each
 * name the reference's new_indexentry keeps, src/fastq.c:590-611), out[3] records that are final, out[4] names that met a DIFFERENT
 * name with the same 64-bit hash on their way into / through the index and walked on (src/hash.c:38-45 does the same along its chain) */
int fqg_memory_stats(fqg_ctx* ctx, uint64_t out[5]);

/* ---- synthetic inputs generated on the device (bench.py, large parity tests); see fq_synth.cu ---- */
int fqg_synth_illumina_record_bytes(void);
int fqg_synth_long_header_bytes(void);
int fqg_synth_illumina(void* device_out, uint64_t first_record, uint64_t n_records, uint64_t seed, int mate,
                       uint64_t perm_window, void* cuda_stream);
int fqg_synth_longreads(void* device_out, const uint64_t* device_offsets, uint64_t first_record, uint64_t n_records,
                        uint64_t seed, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif

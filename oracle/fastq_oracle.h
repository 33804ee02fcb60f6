/*
 * oracle/fastq_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's fastq_info hot path (record split, validation, read-name
 * uniqueness / mate matching, report).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this.  The product (libfastq_gpu) never links or calls it.
 *
 * Parity status: PINNED — checked against the transcripts (exit status, stdout, stderr) of the
 * unmodified reference binary (oracle/_ref/fastq_info) over the reference's own test corpus and
 * hand-made edge files (tests/golden/transcripts.json, made by tests/golden/make_golden.py), and
 * differentially fuzzed against that binary in tests/test_oracle_fuzz.py.
 */
#ifndef FASTQ_ORACLE_H
#define FASTQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int rc;          /* process exit status the reference would return          */
  char *out;       /* what the reference would write to stdout (malloc'd)      */
  size_t out_len;
  char *err;       /* what the reference would write to stderr (malloc'd)      */
  size_t err_len;
} oracle_result;

/*
 * Run `fastq_info argv[1..]` on already-decompressed streams.
 *   argv      : exactly what the reference's main() would receive (argv[0] = program name)
 *   f1 / f2   : decompressed bytes of the first / second positional file (f2 may be NULL);
 *               a length of (size_t)-1 means "the caller could not open it" (fastq.c:651-655)
 * Mirrors fastq_info.c:190-396.
 */
int oracle_fastq_info(int argc, const char **argv,
                      const uint8_t *f1, size_t n1,
                      const uint8_t *f2, size_t n2,
                      oracle_result *res);
void oracle_free(oracle_result *res);

/* Stand-alone pieces used by unit tests. */
const char *oracle_qual_range2enc(unsigned int min_qual, unsigned int max_qual); /* fastq.c:274-297 */
/* name normaliser (fastq.c:442-516) for a given format (0 DEFAULT,1 CASAVA18,2 INT/NOP); returns name length,
 * writes *len_p (the value the reference feeds into index_mem). hdr = raw header line (NUL-terminated). */
long oracle_readname(const char *hdr, int format, int is_pe, char *name_out, unsigned long *len_p);
/* format sniff (fastq.c:459-478): returns 0 DEFAULT, 1 CASAVA18, 2 INTEGERNAME, 3 NOP("no suffix") */
int oracle_sniff_format(const char *name_after_at);
/* colour-space sniff (fastq.c:731-754): 1 colour space, 0 sequence space */
int oracle_sniff_colorspace(const char *seq_line);
/* number of gzgets-style lines/records in a stream (for host-logic tests) */
uint64_t oracle_count_newlines(const uint8_t *p, size_t n);

#ifdef __cplusplus
}
#endif
#endif

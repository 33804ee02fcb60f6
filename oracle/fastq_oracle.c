/*
 * oracle/fastq_oracle.c — TEST INFRASTRUCTURE ONLY (see fastq_oracle.h).
 *
 * A sequential, in-memory restatement of what `fastq_info` (reference v0.25.3) does, written from
 * the behaviour described by /root/reference/src/{fastq_info.c,fastq.c,hash.c}.  It is NOT the product:
 * it is the checker the CUDA path is compared against.  Each function cites the reference lines it
 * follows.  The reference exits from deep inside its library; here that is a longjmp back to the entry.
 *
 * Parity: PINNED against the unmodified reference binary (tests/golden/, tests/test_oracle_*.py).
 */
#define _GNU_SOURCE
#include "fastq_oracle.h"
#include <regex.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define O_MAX_READ 2500000L /* fastq.h:30-33  MAX_READ_LENGTH  */
#define O_MAX_LABEL 1000L   /* fastq.h:35-37  MAX_LABEL_LENGTH */
#define O_MAX_PHRED 126u    /* fastq.h:46 */
#define O_HASHSIZE 39000001UL
#define FMT_UNDEF (-1)
#define FMT_DEFAULT 0
#define FMT_CASAVA 1
#define FMT_INT 2 /* INTEGERNAME == NOP == 2 (fastq.h:25-28) */
#define SP_UNDEF (-1)
#define SP_SEQ 0
#define SP_COLOR 1

/* ------------------------------------------------------------------ text sinks */
typedef struct { char *p; size_t n, cap; } sbuf;
static void sb_put(sbuf *s, const char *d, size_t k) {
  if (s->n + k + 1 > s->cap) {
    s->cap = (s->n + k + 1) * 2 + 64;
    s->p = (char *)realloc(s->p, s->cap);
  }
  memcpy(s->p + s->n, d, k);
  s->n += k;
  s->p[s->n] = 0;
}
static void sb_printf(sbuf *s, const char *fmt, ...) {
  va_list ap, ap2;
  va_start(ap, fmt);
  va_copy(ap2, ap);
  int k = vsnprintf(NULL, 0, fmt, ap);
  va_end(ap);
  char *tmp = (char *)malloc((size_t)k + 1);
  vsnprintf(tmp, (size_t)k + 1, fmt, ap2);
  va_end(ap2);
  sb_put(s, tmp, (size_t)k);
  free(tmp);
}

typedef struct {
  sbuf out, err;
  jmp_buf bail;
  int rc;
} orun;

static void o_exit(orun *r, int code) { r->rc = code; longjmp(r->bail, 1); }
/* PRINT_ERROR, fastq.h:69 : blank line, "ERROR: ", message, newline — all on stderr */
#define O_ERROR(r, ...) do { sb_printf(&(r)->err, "\nERROR: "); sb_printf(&(r)->err, __VA_ARGS__); sb_printf(&(r)->err, "\n"); } while (0)

/* ------------------------------------------------------------------ stream = gzFile stand-in */
typedef struct {
  const uint8_t *p;
  size_t n, pos;
  int past; /* zlib's state->past: a read found nothing left */
} ostream;

/* zlib gzgets(): up to max-1 bytes, stops after '\n'; NULL if nothing was read. */
static char *o_gets(ostream *s, char *buf, long max) {
  long left = max - 1, k = 0;
  int eol = 0;
  if (left) do {
    if (s->pos >= s->n) { s->past = 1; break; }
    char c = (char)s->p[s->pos++];
    buf[k++] = c; left--;
    eol = (c == '\n');
  } while (left && !eol);
  if (k == 0) return NULL;
  buf[k] = 0;
  return buf;
}
/* GZ_READ, fastq.c:202-209 */
static void o_read_line(ostream *s, char *buf, long max) { if (!o_gets(s, buf, max)) buf[0] = 0; }

/* ------------------------------------------------------------------ FASTQ_FILE / FASTQ_ENTRY */
typedef struct {
  ostream fd;
  unsigned long cline; /* zero: the reference relies on fresh mmap'd memory (fastq.c:163-188) */
  const char *filename;
  unsigned long max_rl, last_rl, min_rl, min_qual, max_qual, num_rds;
  unsigned long *rdlen_ctr;
  int is_pe, readname_format, is_casava_18, space;
} ofile;

typedef struct {
  char *hdr1, *hdr2, *seq, *qual;
  unsigned long read_len;
} oentry;

#define O_UNOPENABLE ((size_t)-1) /* caller could not open the file: fastq_open fails, fastq.c:651-655 */
static ofile *o_file_new_raw(const char *name, const uint8_t *p, size_t n) { /* fastq_new, fastq.c:163-188 */
  ofile *f = (ofile *)calloc(1, sizeof(ofile));
  f->fd.p = p; f->fd.n = n;
  f->filename = name;
  f->min_rl = O_MAX_READ; f->min_qual = O_MAX_PHRED;
  f->readname_format = FMT_UNDEF; f->is_casava_18 = FMT_UNDEF; f->space = SP_UNDEF;
  f->rdlen_ctr = (unsigned long *)calloc(O_MAX_READ, sizeof(unsigned long));
  return f;
}
static ofile *o_file_open(orun *r, const char *name, const uint8_t *p, size_t n) {
  if (n == O_UNOPENABLE) {
    O_ERROR(r, "Unable to open %s", name);
    o_exit(r, 1);
  }
  return o_file_new_raw(name, p, n);
}
#define o_file_new(name, p, n) o_file_open(r, name, p, n)
static void o_file_free(ofile *f) { if (f) { free(f->rdlen_ctr); free(f); } }
static oentry *o_entry_new(void) {
  oentry *e = (oentry *)calloc(1, sizeof(oentry));
  e->hdr1 = (char *)calloc(1, O_MAX_LABEL); e->hdr2 = (char *)calloc(1, O_MAX_LABEL);
  e->seq = (char *)calloc(1, O_MAX_READ);   e->qual = (char *)calloc(1, O_MAX_READ);
  return e;
}
static void o_entry_free(oentry *e) { if (e) { free(e->hdr1); free(e->hdr2); free(e->seq); free(e->qual); free(e); } }

/* fastq_new_entry_stats, fastq.c:97-110 */
static void o_entry_stats(ofile *f, oentry *e) {
  unsigned long slen = e->read_len;
  if (slen < f->min_rl) f->min_rl = slen;
  if (slen > f->max_rl) f->max_rl = slen;
  ++f->num_rds;
  f->last_rl = slen;
  f->rdlen_ctr[slen]++;
}

/* fastq_read_entry, fastq.c:245-261 */
static int o_read_entry(orun *r, ofile *f, oentry *e) {
  if (f->fd.past) return 0;
  o_read_line(&f->fd, e->hdr1, O_MAX_LABEL);
  if (e->hdr1[0] == '\0') return 0;
  o_read_line(&f->fd, e->seq, O_MAX_READ);
  o_read_line(&f->fd, e->hdr2, O_MAX_LABEL);
  o_read_line(&f->fd, e->qual, O_MAX_READ);
  if (e->seq[0] == '\0' || e->hdr2[0] == '\0' || e->qual[0] == '\0') {
    O_ERROR(r, "Error in file %s: line %lu: file truncated", f->filename, f->cline);
    o_exit(r, 1);
  }
  f->cline += 4;
  e->read_len = strlen(e->seq);
  return 1;
}
/* fastq_read_next_entry, fastq.c:237-243 */
static int o_read_next_entry(orun *r, ofile *f, oentry *e) {
  int k = o_read_entry(r, f, e);
  if (k <= 0) return k;
  o_entry_stats(f, e);
  return 1;
}

/* ------------------------------------------------------------------ sniffers, fastq.c:666-754 */
static int o_regex(const char *pat, int flags, const char *s) {
  regex_t re;
  if (regcomp(&re, pat, flags)) return 0;
  int hit = regexec(&re, s, 0, NULL, 0) == 0;
  regfree(&re);
  return hit;
}
int oracle_sniff_format(const char *rn) {
  if (o_regex("[A-Z0-9:]* [1234]:[YN]:[0-9]*.*", 0, rn)) return 1;            /* :672 */
  if (o_regex("^[0-9]+[\n\r]?$", REG_EXTENDED, rn)) return 2;                   /* :694 */
  if (!o_regex("[# \t/:][0-9abAB][\n\r]?$", REG_EXTENDED, rn)) return 3;        /* :714 no suffix */
  return 0;
}
int oracle_sniff_colorspace(const char *seq) { return o_regex("^[GT]?[0123n\\.NtT]+\n?$", REG_EXTENDED, seq); } /* :737 */

/* fastq_get_readname, fastq.c:442-516 */
static char *o_get_readname(orun *r, ofile *f, oentry *e, char *rn, unsigned long *len_p, int is_header1) {
  unsigned long len = 0;
  char *hdr = is_header1 ? e->hdr1 : e->hdr2;
  if (is_header1 && hdr[0] != '@') {
    O_ERROR(r, "Error in file %s: line %lu: wrong header %s", f->filename, f->cline, hdr);
    o_exit(r, 3);
  }
  strncpy(rn, hdr + 1, O_MAX_LABEL - 1);
  if (f->readname_format == FMT_UNDEF) { /* once per file, :459-478 */
    switch (oracle_sniff_format(rn)) {
      case 1: sb_printf(&r->err, "CASAVA=1.8\n"); f->readname_format = FMT_CASAVA; break;
      case 2: sb_printf(&r->err, "Read name provided as an integer\n"); f->readname_format = FMT_INT; break;
      case 3: sb_printf(&r->err, "Read name provided with no suffix\n"); f->readname_format = FMT_INT; break;
      default: f->readname_format = FMT_DEFAULT;
    }
  }
  if (f->space == SP_UNDEF) { /* :480-485 */
    f->space = oracle_sniff_colorspace(e->seq) ? SP_COLOR : SP_SEQ;
    if (f->space == SP_COLOR) sb_printf(&r->err, "Color space\n");
  }
  switch (f->readname_format) {
    case FMT_DEFAULT: /* :489-495 */
      len = strlen(rn);
      if (f->is_pe) len--;
      if (len >= 1) rn[len - 1] = '\0'; /* len==0 is out-of-bounds in the reference; left untouched here */
      break;
    case FMT_INT: /* :497-501 */
      len = strlen(rn);
      if (len >= 1) rn[len - 1] = '\0';
      break;
    case FMT_CASAVA: /* :502-511 */
      len = 0;
      while (rn[len] != ' ' && rn[len] != '\0') ++len;
      rn[len] = '\0';
      if (len >= 2 && rn[len - 2] == '/') { rn[len - 2] = '\0'; len -= 2; }
      break;
  }
  *len_p = len;
  return rn;
}

long oracle_readname(const char *hdr, int format, int is_pe, char *name_out, unsigned long *len_p) {
  orun r; memset(&r, 0, sizeof r);
  ofile f; memset(&f, 0, sizeof f);
  oentry e; memset(&e, 0, sizeof e);
  f.readname_format = format; f.space = SP_SEQ; f.is_pe = is_pe; f.filename = "";
  e.hdr1 = (char *)hdr; e.seq = (char *)"";
  unsigned long lp = 0;
  if (hdr[0] != '@') return -1;
  o_get_readname(&r, &f, &e, name_out, &lp, 1);
  if (len_p) *len_p = lp;
  free(r.err.p); free(r.out.p);
  return (long)strlen(name_out);
}

/* compare_headers, fastq.c:543-566 */
static int o_compare_headers(const char *h1, const char *h2) {
  unsigned i = 0;
  if (h2[0] == '\n' || h2[0] == '\r' || h2[0] == '\0') return 1;
  while (h1[i] != '\0' && h2[i] != '\0' && h1[i] == h2[i]) i++;
  for (unsigned a = i; h1[a] != '\0'; a++) if (h1[a] != '\r' && h1[a] != '\n') return 0;
  for (unsigned b = i; h2[b] != '\0'; b++) if (h2[b] != '\r' && h2[b] != '\n') return 0;
  return 1;
}

static int o_is_base(char c) { return c && strchr("ACGTUacgtu0123nN.", c) != NULL; }

/* fastq_validate_entry, fastq.c:300-392 : 0 ok, 1 error already printed */
static int o_validate_entry(orun *r, ofile *f, oentry *e) {
  char rn1[O_MAX_LABEL], rn2[O_MAX_LABEL];
  if (e->hdr1[0] != '@') {
    O_ERROR(r, "Error in file %s: line %lu: sequence identifier should start with an @ - %s", f->filename, f->cline, e->hdr1);
    return 1;
  }
  if (e->hdr1[1] == '\0' || e->hdr1[1] == '\n' || e->hdr1[1] == '\r') {
    O_ERROR(r, "Error in file %s: line %lu: sequence identifier should be longer than 1", f->filename, f->cline);
    return 1;
  }
  unsigned long slen = 0;
  int seen_t = 0, seen_u = 0;
  for (; e->seq[slen] != '\0' && e->seq[slen] != '\n' && e->seq[slen] != '\r'; slen++) {
    char c = e->seq[slen];
    if (!o_is_base(c)) {
      O_ERROR(r, "Error in file %s: line %lu: invalid character '%c' (hex. code:'%x'), expected ACGTUacgtu0123nN.", f->filename, f->cline + 1, c, c);
      return 1;
    }
    if (c == 'U' || c == 'u') {
      seen_u = 1;
      if (seen_t) { O_ERROR(r, "Error in file %s: line %lu: read contains both U and T bases", f->filename, f->cline - 2); return 1; }
    } else if (c == 'T' || c == 't') {
      seen_t = 1;
      if (seen_u) { O_ERROR(r, "Error in file %s: line %lu: read contains both U and T bases", f->filename, f->cline - 2); return 1; }
    }
  }
  o_entry_stats(f, e);
  if (slen < 1) {
    O_ERROR(r, "Error in file %s: line %lu: read length too small - %lu", f->filename, f->cline + 1, slen);
    return 1;
  }
  if (e->hdr2[0] != '+') {
    O_ERROR(r, "Error in file %s: line %lu:  header2 wrong. The line should contain only '+' followed by a newline or read name (header1).", f->filename, f->cline + 2);
    return 1;
  }
  unsigned long len;
  if (e->hdr2[0] != '\0' && e->hdr2[0] != '\r') {
    char *a = o_get_readname(r, f, e, rn1, &len, 1);
    char *b = o_get_readname(r, f, e, rn2, &len, 0);
    if (!o_compare_headers(a, b)) {
      O_ERROR(r, "Error in file %s: line %lu:  header2 differs from header1\nheader 1 \"%s\"\nheader 2 \"%s\"", f->filename, f->cline, e->hdr1, e->hdr2);
      return 1;
    }
  }
  unsigned long qlen = 0;
  for (; e->qual[qlen] != '\0' && e->qual[qlen] != '\n' && e->qual[qlen] != '\r'; qlen++) {
    unsigned int x = (unsigned int)e->qual[qlen]; /* signed char → sign extension, :374 */
    if (x < f->min_qual) f->min_qual = x;
    if (x > f->max_qual) f->max_qual = x;
  }
  if (f->space == SP_SEQ && qlen != slen) {
    O_ERROR(r, "Error in file %s: line %lu: sequence and quality don't have the same length %lu!=%lu", f->filename, f->cline, slen, qlen);
    return 1;
  }
  if (f->space == SP_COLOR && (qlen == slen - 1 || qlen == slen)) return 0;
  if (f->space == SP_COLOR) {
    O_ERROR(r, "Error in file %s: line %lu: sequence and quality length don't match %lu!=%lu", f->filename, f->cline, slen, qlen);
    return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------ exact string set (hash.c + fastq.c:529-611) */
typedef struct onode { struct onode *next; char *name; } onode;
typedef struct { onode **b; size_t nb; unsigned long long n_entries; } oset;
static unsigned long long o_strhash(const char *s) { /* any hash: unobservable behind exact strcmp */
  unsigned long long h = 1469598103934665603ULL;
  for (; *s; s++) { h ^= (unsigned char)*s; h *= 1099511628211ULL; }
  return h;
}
static oset *o_set_new(void) { oset *t = (oset *)calloc(1, sizeof *t); t->nb = 1u << 16; t->b = (onode **)calloc(t->nb, sizeof(onode *)); return t; }
static void o_set_grow(oset *t) {
  size_t nb2 = t->nb * 4; onode **b2 = (onode **)calloc(nb2, sizeof(onode *));
  for (size_t i = 0; i < t->nb; i++) for (onode *x = t->b[i]; x;) { onode *nx = x->next; size_t j = o_strhash(x->name) & (nb2 - 1); x->next = b2[j]; b2[j] = x; x = nx; }
  free(t->b); t->b = b2; t->nb = nb2;
}
static onode **o_set_find(oset *t, const char *name) {
  onode **pp = &t->b[o_strhash(name) & (t->nb - 1)];
  for (; *pp; pp = &(*pp)->next) if (!strcmp((*pp)->name, name)) return pp;
  return NULL;
}
static void o_set_add(oset *t, const char *name) {
  if (t->n_entries > t->nb * 2) o_set_grow(t);
  onode *x = (onode *)malloc(sizeof *x); x->name = strdup(name);
  size_t j = o_strhash(name) & (t->nb - 1); x->next = t->b[j]; t->b[j] = x; t->n_entries++;
}
static void o_set_del(oset *t, onode **pp) { onode *x = *pp; *pp = x->next; free(x->name); free(x); t->n_entries--; }
static void o_set_free(oset *t) {
  if (!t) return;
  for (size_t i = 0; i < t->nb; i++) for (onode *x = t->b[i]; x;) { onode *nx = x->next; free(x->name); free(x); x = nx; }
  free(t->b); free(t);
}

/* PRINT_READS_PROCESSED, fastq.h:82 */
static void o_progress(orun *r, unsigned long c) { if (c % 100000 == 0) sb_printf(&r->err, "\b\b\b\b\b\b\b\b\b\b\b\b\b\b\b%lu", c); }

/* ------------------------------------------------------------------ the loops, fastq_info.c:57-176 and fastq.c:396-439 */
typedef struct { ofile *fd1, *fd2; oentry *m1, *m2; oset *index; unsigned long index_mem; } octx;

static ofile *o_validate_interleaved(orun *r, octx *c, const char *name, const uint8_t *p, size_t n) { /* :57-106 */
  sb_printf(&r->err, "Paired-end interleaved\n");
  ofile *f = c->fd1 = o_file_new(name, p, n);
  f->is_pe = 1;
  char rn1[O_MAX_LABEL], rn2[O_MAX_LABEL];
  unsigned long len = 0;
  while (!f->fd.past) {
    if (!o_read_entry(r, f, c->m1)) break;
    if (!o_read_entry(r, f, c->m2)) {
      O_ERROR(r, "Error in file %s: line %lu: file truncated?", name, f->cline);
      o_exit(r, 3);
    }
    char *a = o_get_readname(r, f, c->m1, rn1, &len, 1);
    char *b = o_get_readname(r, f, c->m2, rn2, &len, 1);
    if (strcmp(a, b)) {
      O_ERROR(r, "Error in file %s: line %lu: unpaired read - %s", name, f->cline, a);
      o_exit(r, 3);
    }
    if (o_validate_entry(r, f, c->m1)) o_exit(r, 3);
    if (o_validate_entry(r, f, c->m2)) o_exit(r, 3);
    o_progress(r, f->cline / 4);
  }
  sb_printf(&r->out, "\n");
  return f;
}

static ofile *o_validate_paired_sorted(orun *r, octx *c, const char *n1, const uint8_t *p1, size_t l1,
                                       const char *n2, const uint8_t *p2, size_t l2) { /* :108-152 */
  ofile *f1 = c->fd1 = o_file_new(n1, p1, l1);
  ofile *f2 = c->fd2 = o_file_new(n2, p2, l2);
  f1->is_pe = f2->is_pe = 1;
  char rn1[O_MAX_LABEL], rn2[O_MAX_LABEL];
  unsigned long len1, len2;
  while (!f1->fd.past) {
    if (!o_read_entry(r, f1, c->m1)) break;
    if (o_validate_entry(r, f1, c->m1)) o_exit(r, 3);
    if (!o_read_entry(r, f2, c->m2)) break;
    if (o_validate_entry(r, f2, c->m2)) o_exit(r, 3);
    o_get_readname(r, f1, c->m1, rn1, &len1, 1);
    o_get_readname(r, f2, c->m2, rn2, &len2, 1);
    if (strcmp(rn1, rn2)) {
      O_ERROR(r, "Readnames do not match across files (read #%ld)", (long)(f1->cline / 4 + 1));
      o_exit(r, 3);
    }
    o_progress(r, f1->cline / 2);
  }
  if (o_read_entry(r, f1, c->m1)) { O_ERROR(r, "Premature end of file2"); o_exit(r, 3); }
  if (o_read_entry(r, f2, c->m2)) { O_ERROR(r, "Premature end of file1"); o_exit(r, 3); }
  sb_printf(&r->out, "\n");
  return f1;
}

static ofile *o_validate_single(orun *r, octx *c, const char *name, const uint8_t *p, size_t n) { /* :155-176 */
  ofile *f = c->fd1 = o_file_new(name, p, n);
  f->is_pe = 1;
  while (!f->fd.past) {
    if (!o_read_entry(r, f, c->m1)) break;
    if (o_validate_entry(r, f, c->m1)) o_exit(r, 3);
    o_progress(r, f->cline / 4);
  }
  sb_printf(&r->out, "\n");
  return f;
}

static void o_index_readnames(orun *r, octx *c, ofile *f) { /* fastq.c:396-439 */
  char rn[O_MAX_LABEL];
  unsigned long len;
  while (!f->fd.past) {
    if (!o_read_next_entry(r, f, c->m1)) break;
    char *name = o_get_readname(r, f, c->m1, rn, &len, 1);
    if (o_set_find(c->index, name)) {
      O_ERROR(r, "Error in file %s: line %lu: duplicated sequence %s", f->filename, f->cline, name);
      o_exit(r, 3);
    }
    o_set_add(c->index, name);
    c->index_mem += 16 + len + 1 + 24; /* sizeof(INDEX_ENTRY)+len+1+sizeof(hashnode), fastq.c:609 */
    if (o_validate_entry(r, f, c->m1)) o_exit(r, 3);
    o_progress(r, f->cline / 4);
  }
}

/* median_rl, fastq_info.c:39-55 */
static unsigned int o_median_rl(ofile *f1, ofile *f2) {
  unsigned long long ctr = 0;
  unsigned int crl = 1;
  unsigned long nreads = f1->num_rds;
  if (f1->num_rds == 1 && f2 == NULL) return (unsigned int)f1->min_rl;
  if (f2) nreads += f2->num_rds;
  while (crl < O_MAX_READ) {
    ctr += f1->rdlen_ctr[crl];
    if (f2) ctr += f2->rdlen_ctr[crl];
    if (f1->num_rds > 1 && ctr > nreads / 2) break;
    ++crl;
  }
  return crl;
}

/* fastq_qualRange2enc, fastq.c:274-297 */
const char *oracle_qual_range2enc(unsigned int lo, unsigned int hi) {
  static const char *names[] = {"33", "64", "solexa", "33 *", "sanger"};
  int enc;
  if (lo >= 33 && lo < 59 && hi >= 90) enc = 4;
  else if (lo >= 33 && hi <= 73) enc = 0;
  else if (lo < 59) enc = 0;
  else if (lo >= 64 && hi > 74) enc = 1;
  else if (lo >= 59 && hi > 74) enc = 2;
  else enc = 3;
  if (hi > O_MAX_PHRED) return NULL;
  if (enc != 4 && hi > lo + 60) return NULL;
  return names[enc];
}

static void o_usage(orun *r, int verbose) { /* fastq_info.c:178-188 */
  sb_printf(&r->out, "Usage: fastq_info [-r -e -s -q -h] fastq1 [fastq2 file|pe]\n");
  if (verbose) {
    sb_printf(&r->out, " -h  : print this help message\n");
    sb_printf(&r->out, " -s  : the reads in the two fastq files have the same ordering\n");
    sb_printf(&r->out, " -e  : do not fail with empty files\n");
    sb_printf(&r->out, " -q  : do not fail if quality encoding cannot be determined\n");
    sb_printf(&r->out, " -r  : skip check for duplicated readnames\n");
  }
}

/* main, fastq_info.c:190-396.  getopt("esfrhq") with GNU argv permutation is restated by hand. */
static void o_main(orun *r, octx *c, int argc, const char **argv_in, const uint8_t *p1, size_t l1, const uint8_t *p2, size_t l2) {
  int is_paired = 0, is_interleaved = 0, is_sorted = 0, empty_ok = 0, no_enc_ok = 0, skip_names = 0, nopt = 0;
  sb_printf(&r->err, "fastq_utils %s\n", "0.25.3");
  /* option scan: option words first (in order), then the rest — what GNU getopt leaves in argv */
  const char **argv = (const char **)calloc((size_t)argc + 1, sizeof(char *));
  int k = 1, stop = 0;
  argv[0] = argv_in[0];
  for (int i = 1; i < argc; i++) {
    const char *w = argv_in[i];
    if (!stop && !strcmp(w, "--")) { argv[k++] = w; stop = 1; continue; }
    if (!stop && w[0] == '-' && w[1] != '\0') argv[k++] = w;
  }
  int nflags_words = k;
  stop = 0;
  for (int i = 1; i < argc; i++) {
    const char *w = argv_in[i];
    if (!stop && !strcmp(w, "--")) { stop = 1; continue; }
    if (stop || !(w[0] == '-' && w[1] != '\0')) argv[k++] = w;
  }
  for (int i = 1; i < nflags_words; i++) {
    const char *w = argv[i];
    if (!strcmp(w, "--")) break;
    for (const char *ch = w + 1; *ch; ch++) {
      switch (*ch) {
        case 'q': no_enc_ok = 1; ++nopt; break;
        case 'e': empty_ok = 1; ++nopt; break;
        case 's': is_sorted = 1; ++nopt; break;
        case 'r': skip_names = 1; ++nopt; break;
        case 'h': o_usage(r, 1); free(argv); o_exit(r, 0); break;
        case 'f':
          sb_printf(&r->err, "Fixing (-f) enabled: Replacing . by N (creating .fix.gz files)\n");
          O_ERROR(r, "-f option is no longer valid.");
          free(argv); o_exit(r, 1); break;
        default:
          ++nopt;
          O_ERROR(r, "Option -%c invalid", *ch);
          free(argv); o_exit(r, 1);
      }
    }
  }
  if (argc - nopt < 2 || argc - nopt > 3) {
    O_ERROR(r, "Invalid number of arguments");
    o_usage(r, 0);
    free(argv); o_exit(r, 1);
  }
  const char *a1 = argv[1 + nopt], *a2 = (argc - nopt == 3) ? argv[2 + nopt] : NULL;
  free(argv);
  if (a2) { is_paired = 1; is_interleaved = strncmp(a2, "pe", 2) == 0; }

  unsigned long num_reads1 = 0, num_reads2 = 0;
  ofile *fd1 = NULL, *fd2 = NULL; /* main()'s own fd2: stays NULL except in the default pair mode */
  if (is_interleaved) {
    fd1 = o_validate_interleaved(r, c, a1, p1, l1);
    num_reads1 = fd1->num_rds;
  } else if (is_paired && is_sorted && skip_names) {
    sb_printf(&r->err, "-s option used: assuming that reads have the same ordering in both files\n");
    fd1 = o_validate_paired_sorted(r, c, a1, p1, l1, a2, p2, l2);
    num_reads1 = fd1->num_rds;
  } else if (!is_paired && skip_names) {
    sb_printf(&r->err, "Skipping check for duplicated read names\n");
    fd1 = o_validate_single(r, c, a1, p1, l1);
    num_reads1 = fd1->num_rds;
  } else {
    fd1 = c->fd1 = o_file_new(a1, p1, l1);
    if (is_paired) fd1->is_pe = 1;
    sb_printf(&r->err, "DEFAULT_HASHSIZE=%lu\n", O_HASHSIZE);
    c->index = o_set_new();
    c->index_mem += 8; /* sizeof(hashtable) = a pointer, fastq_info.c:293 */
    sb_printf(&r->err, "Scanning and indexing all reads from %s\n", fd1->filename);
    o_index_readnames(r, c, fd1);
    sb_printf(&r->err, "Scanning complete.\n");
    num_reads1 = c->index->n_entries;
    sb_printf(&r->err, "\n");
    sb_printf(&r->err, "Reads processed: %llu\n", c->index->n_entries);
    sb_printf(&r->err, "Memory used in indexing: ~%ld MB\n", (long)(c->index_mem / 1024 / 1024));
  }
  if (num_reads1 == 0) { /* :304-314 */
    if (empty_ok) {
      sb_printf(&r->out, "Number of reads: %lu\n", 0L);
      sb_printf(&r->out, "Quality encoding range: %lu %lu\n", 0L, 0L);
      sb_printf(&r->out, "Quality encoding: %s\n", "");
      sb_printf(&r->out, "Read length: %lu %lu %u\n", 0L, 0L, 0);
      o_exit(r, 0);
    }
    O_ERROR(r, "No reads found in %s.", a1);
    o_exit(r, 3);
  }
  unsigned long min_rl = fd1->min_rl, max_rl = fd1->max_rl, min_qual = fd1->min_qual, max_qual = fd1->max_qual;
  if (a2 && !is_interleaved && !is_sorted) { /* mate pass, :322-362 */
    sb_printf(&r->err, "File %s processed\n", a1);
    sb_printf(&r->err, "Next file %s\n", a2);
    fd2 = c->fd2 = o_file_new(a2, p2, l2);
    fd2->is_pe = 1;
    unsigned long len;
    char rn[O_MAX_LABEL];
    while (!fd2->fd.past) {
      if (!o_read_entry(r, fd2, c->m2)) break;
      char *name = o_get_readname(r, fd2, c->m2, rn, &len, 1);
      onode **hit = o_set_find(c->index, name);
      if (!hit) {
        O_ERROR(r, "Error in file %s: line %lu: unpaired read - %s", a2, fd2->cline, name);
        o_exit(r, 3);
      }
      o_set_del(c->index, hit);
      if (o_validate_entry(r, fd1, c->m2)) o_exit(r, 3); /* sic: fd1 */
      o_progress(r, fd2->cline / 4);
    }
    sb_printf(&r->out, "\n");
    if (c->index->n_entries > 0) {
      O_ERROR(r, "Error in file %s: found %llu unpaired reads", a1, c->index->n_entries);
      o_exit(r, 3);
    }
    if (fd2->min_rl < min_rl) min_rl = fd2->min_rl;
    if (fd2->max_rl > max_rl) max_rl = fd2->max_rl;
    if (fd2->min_qual < min_qual) min_qual = fd2->min_qual;
    if (fd2->max_qual > max_qual) max_qual = fd2->max_qual;
  }
  sb_printf(&r->err, "------------------------------------\n");
  if (num_reads2 > 0) sb_printf(&r->err, "Number of reads: %lu %lu\n", num_reads1, num_reads2);
  else sb_printf(&r->err, "Number of reads: %lu\n", num_reads1);
  const char *enc = oracle_qual_range2enc((unsigned int)min_qual, (unsigned int)max_qual);
  if (!enc && !no_enc_ok) {
    if (max_qual > O_MAX_PHRED) O_ERROR(r, "Unable to determine quality encoding - unknown range [%lu,>%u]", min_qual, O_MAX_PHRED);
    else O_ERROR(r, "Unable to determine quality encoding - unknown range [%lu,%lu]", min_qual, max_qual);
    o_exit(r, 3);
  }
  sb_printf(&r->err, "Quality encoding range: %lu %lu\n", min_qual, max_qual);
  if (!enc && no_enc_ok) sb_printf(&r->err, "Quality encoding: NA\n");
  else sb_printf(&r->err, "Quality encoding: %s\n", enc);
  sb_printf(&r->err, "Read length: %lu %lu %u\n", min_rl - 1, max_rl - 1, o_median_rl(fd1, fd2) - 1);
  sb_printf(&r->err, "OK\n");
  o_exit(r, 0);
}

int oracle_fastq_info(int argc, const char **argv, const uint8_t *f1, size_t n1, const uint8_t *f2, size_t n2, oracle_result *res) {
  static orun r; /* static: jmp_buf + volatile-free locals */
  static octx c;
  memset(&r, 0, sizeof r);
  memset(&c, 0, sizeof c);
  c.m1 = o_entry_new();
  c.m2 = o_entry_new();
  if (!setjmp(r.bail)) o_main(&r, &c, argc, argv, f1, n1, f2, n2);
  o_entry_free(c.m1); o_entry_free(c.m2);
  o_file_free(c.fd1); o_file_free(c.fd2);
  o_set_free(c.index);
  if (!r.out.p) sb_put(&r.out, "", 0);
  if (!r.err.p) sb_put(&r.err, "", 0);
  res->rc = r.rc;
  res->out = r.out.p; res->out_len = r.out.n;
  res->err = r.err.p; res->err_len = r.err.n;
  return r.rc;
}
void oracle_free(oracle_result *res) { free(res->out); free(res->err); res->out = res->err = NULL; }

uint64_t oracle_count_newlines(const uint8_t *p, size_t n) {
  uint64_t k = 0;
  for (size_t i = 0; i < n; i++) k += p[i] == '\n';
  return k;
}

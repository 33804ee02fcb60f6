/*
 * fq_record.h — per-record semantics of the fastq_info hot path, written once for host and device.
 *
 * Everything here is a pure function of the bytes of one record (four raw gz-lines) and a small context.
 * The CUDA kernels (fq_cuda.cu) call these from one thread per record; the warp-cooperative long-read kernel
 * re-implements only the two bulk scans (sequence alphabet, quality min/max) and defers to the functions
 * here for everything else.  The `explain` kernel calls the same code to produce message details.
 *
 * Reference behaviour restated (no code shared): src/fastq.c:245-261 (reader), :300-392 (validator),
 * :442-516 (read-name normaliser), :543-566 (header compare), :666-754 (format / colour-space sniffers).
 * A gz-line is a C string for the reference: everything after the first NUL byte is invisible to it.
 */
#ifndef FQ_RECORD_H
#define FQ_RECORD_H
#include "fq_types.h"

/* ---------------------------------------------------------------- byte access */
FQ_HD uint32_t fq_ld32(const uint8_t* d, uint32_t aligned_off) { return *(const uint32_t*)(d + aligned_off); }

/* mask selecting the bytes of the aligned word at `a` whose addresses lie in [s, e) (requires a+4 > s, a < e) */
FQ_HD uint32_t fq_bytemask(uint32_t a, uint32_t s, uint32_t e) {
  uint32_t lo = s > a ? s - a : 0u;          /* 0..3 */
  uint32_t hi = e - a < 4u ? e - a : 4u;     /* 1..4 */
  uint32_t m = hi == 4u ? 0xFFFFFFFFu : ((1u << (8u * hi)) - 1u);
  return m & ~((1u << (8u * lo)) - 1u);
}
/* exact "some byte of x is zero" → 0x80 in each zero byte */
FQ_HD uint32_t fq_zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); }

/* length of the C string held by the raw line [off, off+len): bytes before the first NUL */
FQ_HD uint32_t fq_cstrlen(const uint8_t* d, uint32_t off, uint32_t len) {
  if (len == 0) return 0;
  uint32_t e = off + len;
  for (uint32_t a = off & ~3u; a < e; a += 4) {
    uint32_t w = fq_ld32(d, a);
    uint32_t z = fq_zero_bytes(w) & fq_bytemask(a, off, e);
    if (z) { /* first zero byte inside the line */
      for (uint32_t i = (a > off ? a : off); i < e; i++) if (d[i] == 0) return i - off;
    }
  }
  return len;
}

/* ---------------------------------------------------------------- sequence line (src/fastq.c:317-341) */
/* fastq_trim_poly_at's two scans over a sequence line as gzgets left it (src/fastq_trim_poly_at.c:77-115).  read_len = strlen of the
 * line buffer: the bytes up to the first NUL, terminators included (src/fastq.c:259).  tail: bytes from index read_len - 2 downwards
 * that are one of N A n a (the loop starts in front of the line's last byte, which is its LF in a complete line); head: bytes from
 * index 0 upwards that are one of N T n t. */
FQ_HD void fq_poly_at(const uint8_t* s, uint32_t len, uint32_t* read_len, uint32_t* tail, uint32_t* head) {
  uint32_t rl = 0;
  while (rl < len && s[rl] != 0) rl++;
  uint32_t a = 0;
  for (long x = (long)rl - 2; x >= 0; --x) { const uint8_t c = s[x]; if (c != 'N' && c != 'A' && c != 'n' && c != 'a') break; ++a; }
  uint32_t b = 0;
  for (uint32_t x = 0; x < rl; ++x) { const uint8_t c = s[x]; if (c != 'N' && c != 'T' && c != 'n' && c != 't') break; ++b; }
  *read_len = rl; *tail = a; *head = b;
}
FQ_HD bool fq_is_base(uint8_t c) {
  switch (c) {
    case 'A': case 'C': case 'G': case 'T': case 'U': case 'a': case 'c': case 'g': case 't': case 'u':
    case '0': case '1': case '2': case '3': case 'n': case 'N': case '.': return true;
    default: return false;
  }
}
/* bit 0 of every byte lane set iff that byte is one of ACGTN/acgtn (the common alphabet; bit 5 = case is ignored) */
FQ_HD uint32_t fq_common_base_lanes(uint32_t w) {
  uint32_t b0 = w, b1 = w >> 1, b2 = w >> 2, b3 = w >> 3, b4 = w >> 4, b6 = w >> 6, b7 = w >> 7;
  uint32_t g00 = b0 & (~b2 | b1);  /* 00001 00011 00111  : A C G */
  uint32_t g01 = b2 & b1 & ~b0;    /* 01110              : N     */
  uint32_t g10 = b2 & ~b1 & ~b0;   /* 10100              : T     */
  uint32_t f = (~b4 & ~b3 & g00) | (~b4 & b3 & g01) | (b4 & ~b3 & g10);
  return f & b6 & ~b7 & 0x01010101u;
}

typedef struct {
  uint32_t slen;      /* bases before the first NUL / LF / CR                        */
  uint32_t code;      /* 0, FQ_E_BADCHAR or FQ_E_UT                                   */
  uint32_t bad;       /* offending byte                                              */
} FqSeqScan;

/* exact, byte by byte: the reference's loop */
FQ_HD FqSeqScan fq_seq_scan_careful(const uint8_t* d, uint32_t off, uint32_t cl) {
  FqSeqScan r; r.slen = 0; r.code = 0; r.bad = 0;
  bool seen_t = false, seen_u = false;
  for (; r.slen < cl; r.slen++) {
    uint8_t c = d[off + r.slen];
    if (c == '\n' || c == '\r') break;
    if (!fq_is_base(c)) { r.code = FQ_E_BADCHAR; r.bad = c; return r; }
    if (c == 'U' || c == 'u') { seen_u = true; if (seen_t) { r.code = FQ_E_UT; return r; } }
    else if (c == 'T' || c == 't') { seen_t = true; if (seen_u) { r.code = FQ_E_UT; return r; } }
  }
  return r;
}
/* word-at-a-time: lines made only of ACGTN (either case) plus an LF / CRLF ending never leave this function */
FQ_HD FqSeqScan fq_seq_scan(const uint8_t* d, uint32_t off, uint32_t rawlen, uint32_t cl) {
  if (cl == rawlen && rawlen > 0) {
    uint32_t e = off + rawlen;
    if (d[e - 1] == '\n') { e--; if (e > off && d[e - 1] == '\r') e--; }
    uint32_t bad = 0;
    if (e > off)
      for (uint32_t a = off & ~3u; a < e; a += 4) {
        uint32_t w = fq_ld32(d, a);
        bad |= ~fq_common_base_lanes(w) & fq_bytemask(a, off, e) & 0x01010101u;
      }
    if (!bad) { FqSeqScan r; r.slen = e - off; r.code = 0; r.bad = 0; return r; }
  }
  return fq_seq_scan_careful(d, off, cl);
}

/* ---------------------------------------------------------------- quality line (src/fastq.c:373-378) */
typedef struct { uint32_t qlen, qmin, qmax; } FqQualScan;
FQ_HD FqQualScan fq_qual_scan_careful(const uint8_t* d, uint32_t off, uint32_t cl) {
  FqQualScan r; r.qlen = 0; r.qmin = 255; r.qmax = 0;
  for (; r.qlen < cl; r.qlen++) {
    uint32_t c = d[off + r.qlen];
    if (c == '\n' || c == '\r') break;
    if (c < r.qmin) r.qmin = c;
    if (c > r.qmax) r.qmax = c;
  }
  return r;
}
FQ_HD uint32_t fq_min2x16(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vminu2(a, b);
#else
  uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
  uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
  return lo | (hi << 16);
#endif
}
FQ_HD uint32_t fq_max2x16(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vmaxu2(a, b);
#else
  uint32_t lo = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
  uint32_t hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
  return lo | (hi << 16);
#endif
}
FQ_HD FqQualScan fq_qual_scan(const uint8_t* d, uint32_t off, uint32_t rawlen, uint32_t cl) {
  if (cl == rawlen && rawlen > 0) {
    uint32_t e = off + rawlen;
    if (d[e - 1] == '\n') { e--; if (e > off && d[e - 1] == '\r') e--; }
    if (e > off) {
      uint32_t mn = 0x00FF00FFu, mx = 0u; /* two 16-bit lanes holding byte values */
      for (uint32_t a = off & ~3u; a < e; a += 4) {
        uint32_t w = fq_ld32(d, a), m = fq_bytemask(a, off, e);
        uint32_t lo = w | ~m, hi = w & m;
        mn = fq_min2x16(mn, fq_min2x16(lo & 0x00FF00FFu, (lo >> 8) & 0x00FF00FFu));
        mx = fq_max2x16(mx, fq_max2x16(hi & 0x00FF00FFu, (hi >> 8) & 0x00FF00FFu));
      }
      FqQualScan r;
      r.qmin = (mn & 0xFFFFu) < (mn >> 16) ? (mn & 0xFFFFu) : (mn >> 16);
      r.qmax = (mx & 0xFFFFu) > (mx >> 16) ? (mx & 0xFFFFu) : (mx >> 16);
      r.qlen = e - off;
      if (r.qmin > 0x0Du) return r; /* no NUL / LF / CR inside: the string is the whole content */
    }
  }
  return fq_qual_scan_careful(d, off, cl);
}

/* ---------------------------------------------------------------- read names (src/fastq.c:442-516) */
/* `hcl` = C-string length of the header line starting at hoff.  Result: name bytes [noff, noff+nlen) and the
 * `len` the reference adds into index_mem.  The string after the marker byte is rn = line+1. */
FQ_HD void fq_readname(const uint8_t* d, uint32_t hoff, uint32_t hcl, int fmt, int is_pe,
                       uint32_t* noff, uint32_t* nlen, uint64_t* mem_len) {
  uint32_t s = hcl >= 1 ? hcl - 1 : 0; /* strlen(rn) */
  const uint8_t* rn = d + hoff + 1;
  *noff = hoff + 1;
  if (fmt == FQ_FMT_CASAVA) { /* :502-511 cut at the first space, then drop a trailing "/x" */
    uint32_t len = 0;
    while (len < s && rn[len] != ' ') ++len;
    if (len >= 2 && rn[len - 2] == '/') len -= 2;
    *nlen = len; *mem_len = len;
  } else if (fmt == FQ_FMT_INT) { /* :497-501 drop the last byte */
    *nlen = s >= 1 ? s - 1 : 0; *mem_len = s;
  } else { /* DEFAULT :489-495 drop the last byte, and one more when paired */
    uint64_t len = (uint64_t)s - (is_pe ? 1u : 0u); /* wraps like the reference's unsigned long when s == 0 */
    *mem_len = len;
    if (s == 0) *nlen = 0;                 /* reference writes out of bounds here; the visible name is "" */
    else if (len >= 1) *nlen = (uint32_t)len - 1;
    else *nlen = s;                        /* len == 0: nothing is cut */
  }
}

/* The name fastq_get_readname (src/fastq.c:442-516) makes of a header line, as a descriptor: hash, offset and length of the name's
 * bytes.  A line that does not start with '@' is the reference's "wrong header" (:448): hash = FQ_HASH_SKIP, len = 0xFFFFFFFF. */
FQ_HD FqName fq_header_name(const uint8_t* d, uint32_t off, uint32_t len, int fmt, int is_pe, uint32_t seed);

/* compare_headers, src/fastq.c:543-566, on two normalised names */
FQ_HD bool fq_compare_headers(const uint8_t* a, uint32_t alen, const uint8_t* b, uint32_t blen) {
  if (blen == 0 || b[0] == '\n' || b[0] == '\r') return true;
  uint32_t i = 0;
  while (i < alen && i < blen && a[i] == b[i]) i++;
  for (uint32_t k = i; k < alen; k++) if (a[k] != '\r' && a[k] != '\n') return false;
  for (uint32_t k = i; k < blen; k++) if (b[k] != '\r' && b[k] != '\n') return false;
  return true;
}

/* 64-bit hash of a name, defined on its little-endian 32-bit words (last one zero-padded) and its length, so
 * that any implementation (byte loop here, funnel-shifted words in a kernel) produces the same value.  The
 * value itself is unobservable: equality is always confirmed on the bytes (reference: hashit + strcmp).
 * Two independent 32-bit multiplicative lanes per word (each step a bijection of its lane, so names that differ in
 * one word never collide in a lane) keep the per-word cost at five 32-bit instructions; a 64-bit finaliser mixes them. */
typedef struct { uint32_t a, b, weak; } FqHashState;
FQ_HD FqHashState fq_hash_init(uint32_t seed) {
  FqHashState h; h.a = 0x85A308D3u ^ (seed * 0x9E3779B1u); h.b = 0x243F6A88u + seed * 0x85EBCA77u; h.weak = seed & FQ_SEED_WEAK;
  return h;
}
FQ_HD void fq_hash_word(FqHashState* h, uint32_t w) {
  h->a = (h->a ^ w) * 0x9E3779B1u;
  h->b = (((h->b << 13) | (h->b >> 19)) ^ w) * 0xC2B2AE3Du;
}
FQ_HD uint64_t fq_hash_fin(FqHashState s, uint32_t len) {
  uint64_t h = ((uint64_t)s.a << 32) | s.b;
  h = (h ^ len) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  if (s.weak) h = (h & 0xFFFull) * 0x0000010000000001ull; /* the same 12 bits pick the slot and (bits 40..) the owning rank */
  return h >= FQ_HASH_SKIP ? h - 2 : h;
}
FQ_HD uint64_t fq_hash_name(const uint8_t* p, uint32_t len, uint32_t seed) {
  FqHashState h = fq_hash_init(seed);
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) {
    uint32_t w = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    fq_hash_word(&h, w);
  }
  if (i < len) {
    uint32_t w = 0;
    for (uint32_t k = 0; i + k < len; k++) w |= (uint32_t)p[i + k] << (8 * k);
    fq_hash_word(&h, w);
  }
  return fq_hash_fin(h, len);
}
FQ_HD bool fq_bytes_equal(const uint8_t* a, const uint8_t* b, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) if (a[i] != b[i]) return false;
  return true;
}
FQ_HD FqName fq_header_name(const uint8_t* d, uint32_t off, uint32_t len, int fmt, int is_pe, uint32_t seed) {
  FqName nm;
  const uint32_t hcl = fq_cstrlen(d, off, len);
  if (hcl == 0 || d[off] != '@') { nm.hash = FQ_HASH_SKIP; nm.off = off; nm.len = 0xFFFFFFFFu; return nm; }
  uint32_t noff, nlen; uint64_t mem;
  fq_readname(d, off, hcl, fmt, is_pe, &noff, &nlen, &mem);
  nm.hash = fq_hash_name(d + noff, nlen, seed); nm.off = noff; nm.len = nlen;
  return nm;
}

/* ---------------------------------------------------------------- sniffers (src/fastq.c:666-754), first record of a file */
/* rn = C string after the '@' (terminators included), length s */
FQ_HD int fq_sniff_format(const uint8_t* rn, uint32_t s) {
  /* is_casava_1_8_readname: unanchored BRE "[A-Z0-9:]* [1234]:[YN]:[0-9]*.*" == contains " [1-4]:[YN]:" */
  for (uint32_t i = 0; i + 5 <= s; i++)
    if (rn[i] == ' ' && rn[i + 1] >= '1' && rn[i + 1] <= '4' && rn[i + 2] == ':' && (rn[i + 3] == 'Y' || rn[i + 3] == 'N') && rn[i + 4] == ':')
      return FQ_SNIFF_CASAVA;
  /* is_int_readname: ^[0-9]+[\n\r]?$ */
  {
    uint32_t i = 0;
    while (i < s && rn[i] >= '0' && rn[i] <= '9') i++;
    if (i >= 1 && (i == s || (i + 1 == s && (rn[i] == '\n' || rn[i] == '\r')))) return FQ_SNIFF_INT;
  }
  /* is_nosuffix_readname: NOT matching [# \t/:][0-9abAB][\n\r]?$ */
  {
    uint32_t e = s;
    bool hit = false;
    for (int strip = 0; strip < 2 && !hit; strip++) {
      if (strip == 1) { if (e >= 1 && (rn[e - 1] == '\n' || rn[e - 1] == '\r')) e--; else break; }
      if (e >= 2) {
        uint8_t c = rn[e - 1], p = rn[e - 2];
        bool cc = (c >= '0' && c <= '9') || c == 'a' || c == 'b' || c == 'A' || c == 'B';
        bool pp = p == '#' || p == ' ' || p == '\t' || p == '/' || p == ':';
        hit = cc && pp;
      }
    }
    if (!hit) return FQ_SNIFF_NOSUFFIX;
  }
  return FQ_SNIFF_DEFAULT;
}
/* is_color_space: ^[GT]?[0123n\.NtT]+\n?$ on the C string of the sequence line; inside a POSIX bracket
 * expression the backslash is an ordinary member of the class */
FQ_HD bool fq_cs_body(uint8_t c) {
  return c == '0' || c == '1' || c == '2' || c == '3' || c == 'n' || c == '\\' || c == '.' || c == 'N' || c == 't' || c == 'T';
}
FQ_HD int fq_sniff_colorspace(const uint8_t* sq, uint32_t s) {
  uint32_t e = s;
  if (e >= 1 && sq[e - 1] == '\n') e--;
  /* body class */
  uint32_t i = 0;
  if (e == 0) return 0;
  bool first_is_gt = sq[0] == 'G' || sq[0] == 'T';
  bool first_is_body = fq_cs_body(sq[0]);
  for (i = 1; i < e; i++) if (!fq_cs_body(sq[i])) return 0;
  /* bytes 1..e-1 are body characters.  Either the first byte is body too (≥1 body byte in total), or it is
   * the optional [GT] prefix followed by at least one body byte. */
  if (first_is_body) return 1;
  if (first_is_gt && e >= 2) return 1;
  return 0;
}

/* ---------------------------------------------------------------- the record (reader + validator) */
FQ_HD uint8_t fq_first_byte(const uint8_t* d, const FqLine& l) { return l.len ? d[l.off] : 0; }

/* Reader step flags + validator verdict + statistics for one record whose four raw lines are L[0..3]. */
FQ_HD void fq_check_record_careful(const uint8_t* d, const FqLine* L, const FqRecCtx& cx, FqRecOut* o) {
  o->flags = 0; o->vrank = FQ_V_OK; o->code = 0; o->read_len = 0; o->slen = 0; o->qlen = 0;
  o->qmin = 255; o->qmax = 0; o->bad = 0; o->name_off = L[0].off + 1; o->name_len = 0; o->mem_len = 0;
  /* fastq_read_entry, src/fastq.c:245-261 */
  uint8_t h0 = fq_first_byte(d, L[0]);
  if (h0 == 0) { o->flags = FQ_RF_STOP; return; }
  if (fq_first_byte(d, L[1]) == 0 || fq_first_byte(d, L[2]) == 0 || fq_first_byte(d, L[3]) == 0) { o->flags = FQ_RF_TRUNC; return; }
  if (h0 != '@') o->flags |= FQ_RF_NOTAT;
  uint32_t cl0 = fq_cstrlen(d, L[0].off, L[0].len);
  uint32_t cl1 = fq_cstrlen(d, L[1].off, L[1].len);
  o->read_len = cl1;
  /* key name (fastq_get_readname on hdr1 with the record's own file's format) */
  if (h0 == '@') fq_readname(d, L[0].off, cl0, cx.fmt_key, cx.pe_key, &o->name_off, &o->name_len, &o->mem_len);
  /* fastq_validate_entry, src/fastq.c:300-392 */
  if (h0 != '@') { o->vrank = FQ_V_AT; o->code = FQ_E_AT; return; }
  uint8_t h1 = cl0 >= 2 ? d[L[0].off + 1] : 0;
  if (h1 == 0 || h1 == '\n' || h1 == '\r') { o->vrank = FQ_V_IDLEN; o->code = FQ_E_IDLEN; return; }
  FqSeqScan sq = fq_seq_scan(d, L[1].off, L[1].len, cl1);
  o->slen = sq.slen;
  if (sq.code) { o->vrank = FQ_V_SEQ; o->code = sq.code; o->bad = sq.bad; return; }
  if (sq.slen < 1) { o->vrank = FQ_V_SHORT; o->code = FQ_E_SHORT; return; }
  if (d[L[2].off] != '+') { o->vrank = FQ_V_PLUS; o->code = FQ_E_PLUS; return; }
  /* header 2 against header 1; "+\n" and "+\r\n" (the common cases) compare equal without looking at hdr1 */
  if (L[2].len >= 2 && d[L[2].off + 1] != '\n' && d[L[2].off + 1] != '\r') {
    uint32_t cl2 = fq_cstrlen(d, L[2].off, L[2].len);
    uint32_t a_off, a_len, b_off, b_len; uint64_t m;
    fq_readname(d, L[0].off, cl0, cx.fmt_val, cx.pe_val, &a_off, &a_len, &m);
    fq_readname(d, L[2].off, cl2, cx.fmt_val, cx.pe_val, &b_off, &b_len, &m);
    if (!fq_compare_headers(d + a_off, a_len, d + b_off, b_len)) { o->vrank = FQ_V_HDR2; o->code = FQ_E_HDR2; return; }
  }
  uint32_t cl3 = fq_cstrlen(d, L[3].off, L[3].len);
  FqQualScan qs = fq_qual_scan(d, L[3].off, L[3].len, cl3);
  o->qlen = qs.qlen; o->qmin = qs.qmin; o->qmax = qs.qmax;
  if (cx.space == FQ_SPACE_COLOR) {
    if (!(qs.qlen == sq.slen - 1 || qs.qlen == sq.slen)) { o->vrank = FQ_V_LEN; o->code = FQ_E_LEN_CS; }
  } else if (qs.qlen != sq.slen) { o->vrank = FQ_V_LEN; o->code = FQ_E_LEN; }
}

/* does a record of this loop reach the name step (index insert / mate claim / pair compare)? */
FQ_HD bool fq_record_has_name(int loop, const FqRecOut& o) {
  if (o.flags & (FQ_RF_STOP | FQ_RF_TRUNC | FQ_RF_NOTAT)) return false;
  return loop != FQ_LOOP_SINGLE && loop != FQ_LOOP_READER;
}


/* ---------------------------------------------------------------- the common case, 16 bytes at a time
 * A record made of: "@name...\n", bases from ACGTN/acgtn + LF or CRLF, "+\n" or "+\r\n", qualities above 0x0D + LF or
 * CRLF, with matching lengths, is clean; for it the functions below produce exactly what fq_check_record_careful
 * produces (tests/sim/test_record_paths.cpp compares the two on random records).  Anything else — every error, NUL
 * bytes, other alphabets, a header-2 that repeats the name — returns false and the careful path decides. */
typedef struct { uint32_t x, y, z, w; } FqU4;
FQ_HD FqU4 fq_ld128(const uint8_t* d, uint32_t a16) {
  FqU4 r;
#if defined(__CUDA_ARCH__)
  uint4 v = *(const uint4*)(d + a16);
  r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
#else
  const uint8_t* p = d + a16;
  uint32_t t[4];
  for (int i = 0; i < 4; i++) t[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  r.x = t[0]; r.y = t[1]; r.z = t[2]; r.w = t[3];
#endif
  return r;
}
/* bits of the 16-byte chunk at `a` whose byte addresses lie in [s, e)   (a < e, a + 16 > s) */
FQ_HD uint32_t fq_range16(uint32_t a, uint32_t s, uint32_t e) {
  uint32_t lo = s > a ? s - a : 0u, hi = e - a < 16u ? e - a : 16u;
  return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}
/* four words holding a flag in bit 7 of every byte → 16 flags in byte order */
FQ_HD uint32_t fq_gather16(uint32_t l0, uint32_t l1, uint32_t l2, uint32_t l3) {
  uint32_t lo = ((((l0 & 0x80808080u) >> 7) | ((l1 & 0x80808080u) >> 3)) * 0x00204081u) >> 21;
  uint32_t hi = ((((l2 & 0x80808080u) >> 7) | ((l3 & 0x80808080u) >> 3)) * 0x00204081u) >> 21;
  return (lo & 0xFFu) | ((hi & 0xFFu) << 8);
}
/* 4 flag bits → 0xFF in each flagged byte */
FQ_HD uint32_t fq_bytes_of4(uint32_t bits4) { return (((bits4 & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu; }
/* ACGTN / acgtn predicate in bit 7 of every byte (other bits undefined).  The bits of a byte are lined up at bit 7 with
 * multiplications by powers of two: on sm_100a these become IMAD.SHL on the FMA pipe, leaving the ALU pipe the seven LOP3s. */
FQ_HD uint32_t fq_base_pred(uint32_t w) {
  uint32_t b0 = w * 128u, b1 = w * 64u, b2 = w * 32u, b3 = w * 16u, b4 = w * 8u, b6 = w * 2u, b7 = w;
  uint32_t g00 = b0 & (~b2 | b1), g01 = b2 & b1 & ~b0, g10 = b2 & ~b1 & ~b0;
  return ((~b4 & ~b3 & g00) | (~b4 & b3 & g01) | (b4 & ~b3 & g10)) & b6 & ~b7;
}
/* every byte of [s, e) is one of ACGTN/acgtn  (e > s) */
FQ_HD bool fq_seq_fast(const uint8_t* d, uint32_t s, uint32_t e) {
  uint32_t a = s & ~15u, alast = (e - 1) & ~15u;
  FqU4 v = fq_ld128(d, a);
  if (~fq_gather16(fq_base_pred(v.x), fq_base_pred(v.y), fq_base_pred(v.z), fq_base_pred(v.w)) & fq_range16(a, s, e)) return false;
  if (a == alast) return true;
  uint32_t all = 0x80808080u;
  for (a += 16; a < alast; a += 16) {
    v = fq_ld128(d, a);
    all &= fq_base_pred(v.x) & fq_base_pred(v.y) & fq_base_pred(v.z) & fq_base_pred(v.w);
  }
  if ((all & 0x80808080u) != 0x80808080u) return false;
  v = fq_ld128(d, alast);
  return !(~fq_gather16(fq_base_pred(v.x), fq_base_pred(v.y), fq_base_pred(v.z), fq_base_pred(v.w)) & fq_range16(alast, s, e));
}
FQ_HD uint32_t fq_odd_bytes(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(w, 0u, 0x4341u);
#else
  return (w >> 8) & 0x00FF00FFu;
#endif
}
/* unsigned min / max over the bytes of [s, e), kept as two 16-bit lanes until the end  (e > s) */
FQ_HD void fq_qual_fast(const uint8_t* d, uint32_t s, uint32_t e, uint32_t* qmin, uint32_t* qmax) {
  uint32_t mn = 0x00FF00FFu, mx = 0u;
  uint32_t a = s & ~15u, alast = (e - 1) & ~15u;
#define FQ_QW(w_) do { uint32_t ev_ = (w_) & 0x00FF00FFu, od_ = fq_odd_bytes(w_); \
    mn = fq_min2x16(mn, fq_min2x16(ev_, od_)); mx = fq_max2x16(mx, fq_max2x16(ev_, od_)); } while (0)
#define FQ_QWM(w_, bm_) do { uint32_t lo_ = (w_) | ~(bm_), hi_ = (w_) & (bm_); \
    mn = fq_min2x16(mn, fq_min2x16(lo_ & 0x00FF00FFu, fq_odd_bytes(lo_))); mx = fq_max2x16(mx, fq_max2x16(hi_ & 0x00FF00FFu, fq_odd_bytes(hi_))); } while (0)
  {
    FqU4 v = fq_ld128(d, a); uint32_t rm = fq_range16(a, s, e);
    FQ_QWM(v.x, fq_bytes_of4(rm)); FQ_QWM(v.y, fq_bytes_of4(rm >> 4)); FQ_QWM(v.z, fq_bytes_of4(rm >> 8)); FQ_QWM(v.w, fq_bytes_of4(rm >> 12));
  }
  if (a != alast) {
    for (a += 16; a < alast; a += 16) { FqU4 v = fq_ld128(d, a); FQ_QW(v.x); FQ_QW(v.y); FQ_QW(v.z); FQ_QW(v.w); }
    FqU4 v = fq_ld128(d, alast); uint32_t rm = fq_range16(alast, s, e);
    FQ_QWM(v.x, fq_bytes_of4(rm)); FQ_QWM(v.y, fq_bytes_of4(rm >> 4)); FQ_QWM(v.z, fq_bytes_of4(rm >> 8)); FQ_QWM(v.w, fq_bytes_of4(rm >> 12));
  }
#undef FQ_QW
#undef FQ_QWM
  *qmin = (mn & 0xFFFFu) < (mn >> 16) ? (mn & 0xFFFFu) : (mn >> 16);
  *qmax = (mx & 0xFFFFu) > (mx >> 16) ? (mx & 0xFFFFu) : (mx >> 16);
}
/* little-endian 32-bit word at any byte offset, from two aligned words */
FQ_HD uint32_t fq_ldu32(const uint8_t* d, uint32_t off) {
  uint32_t a = off & ~3u, sh = (off & 3u) * 8u;
  uint32_t lo = fq_ld32(d, a), hi = fq_ld32(d, a + 4);
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
FQ_HD uint32_t fq_low_bytes(uint32_t w, uint32_t nbytes) { return nbytes >= 4 ? w : (w & ((1u << (8u * nbytes)) - 1u)); }
/* same value as fq_hash_name, four bytes at a time */
FQ_HD uint64_t fq_hash_name_words(const uint8_t* d, uint32_t off, uint32_t len, uint32_t seed) {
  FqHashState h = fq_hash_init(seed);
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) fq_hash_word(&h, fq_ldu32(d, off + i));
  if (i < len) fq_hash_word(&h, fq_low_bytes(fq_ldu32(d, off + i), len - i));
  return fq_hash_fin(h, len);
}

/* Header line and name hash in ONE walk over the line's words (aligned loads, a carried word and a funnel shift give the
 * words of the name, which starts one byte into the line).  Same results as fq_header_fast + fq_hash_name_words; false
 * where fq_header_fast is false. */
/* 0x80 in every byte of x that is <= 0x20 (exact) */
FQ_HD uint32_t fq_low_bytes_flags(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | x) & 0x80808080u; }
FQ_HD bool fq_header_hash_fast(const uint8_t* d, uint32_t h0, uint32_t hl, int fmt, int is_pe, uint32_t seed, bool want_hash,
                               uint32_t* name_len, uint64_t* mem_len_out, uint64_t* hash) {
  if (hl < 3) return false;
  if (d[h0] != '@' || d[h0 + hl - 1] != '\n') return false;
  { uint8_t c1 = d[h0 + 1]; if (c1 == 0 || c1 == '\n' || c1 == '\r') return false; }
  const uint32_t ns = h0 + 1, s = hl - 1;
  const uint32_t sh = (ns & 3u) * 8u;
  uint32_t a = ns & ~3u;
  uint32_t cur = fq_ld32(d, a);
  FqHashState h = fq_hash_init(seed);
  /* length of the name when no space cuts it: the line without its LF (INT), and without one more byte (DEFAULT: the mate digit
   * or the CR of a CRLF line, exactly as the reference drops it) */
  uint32_t nlen; uint64_t mem_len;
  bool casava = fmt == FQ_FMT_CASAVA;
  if (fmt == FQ_FMT_INT) { nlen = s - 1; mem_len = s; }
  else if (!casava) { uint32_t len = s - (is_pe ? 1u : 0u); mem_len = len; nlen = len >= 1 ? len - 1 : s; }
  else { nlen = s; mem_len = s; }
  uint32_t i = 0;
  bool cut = false; /* casava: the first space was found */
  for (; i < s; i += 4) {
    uint32_t next = fq_ld32(d, a + 4); a += 4;
#if defined(__CUDA_ARCH__)
    uint32_t w = __funnelshift_r(cur, next, sh);
#else
    uint32_t w = sh ? (cur >> sh) | (next << (32u - sh)) : cur;
#endif
    cur = next;
    /* a word without a byte <= 0x20 (NUL, space, terminators, control bytes) needs no closer look: inside the name it is
     * hashed, beyond it skipped */
    if (!fq_low_bytes_flags(w) && i + 4 <= s) {
      if (i + 4 <= nlen) { if (want_hash) fq_hash_word(&h, w); continue; }
      if (i >= nlen) continue;
    }
    uint32_t z = fq_zero_bytes(w);
    if (casava && !cut) z |= fq_zero_bytes(w ^ 0x20202020u);
    if (s - i < 4) z &= (1u << (8u * (s - i))) - 1u;
    if (z) {
      uint32_t k = 0; while (!((z >> (8 * k + 7)) & 1u)) k++;
      if (((w >> (8 * k)) & 0xFFu) == 0) return false; /* NUL: the careful path's strlen would stop here */
      /* a space: the name ends here (then a trailing "/x" goes) */
      uint32_t len = i + k;
      cut = true;
      if (len >= 2 && d[ns + len - 2] == '/') { len -= 2; nlen = len; mem_len = len; if (want_hash) h = fq_hash_init(seed); i = 0; break; }
      nlen = len; mem_len = len;
      if (want_hash && k) fq_hash_word(&h, w & ((1u << (8u * k)) - 1u));
      /* the rest of this word and of the line may still hide a NUL */
      uint32_t z2 = fq_zero_bytes(w) ; if (s - i < 4) z2 &= (1u << (8u * (s - i))) - 1u;
      if (z2) return false;
      i += 4;
      for (; i < s; i += 4) {
        uint32_t nx = fq_ld32(d, a + 4); a += 4;
#if defined(__CUDA_ARCH__)
        uint32_t w2 = __funnelshift_r(cur, nx, sh);
#else
        uint32_t w2 = sh ? (cur >> sh) | (nx << (32u - sh)) : cur;
#endif
        cur = nx;
        if (!fq_low_bytes_flags(w2)) continue;
        uint32_t z3 = fq_zero_bytes(w2);
        if (s - i < 4) z3 &= (1u << (8u * (s - i))) - 1u;
        if (z3) return false;
      }
      *name_len = nlen; *mem_len_out = mem_len;
      if (want_hash) *hash = fq_hash_fin(h, nlen);
      return true;
    }
    if (want_hash) {
      if (i + 4 <= nlen) fq_hash_word(&h, w);
      else if (i < nlen) fq_hash_word(&h, w & ((1u << (8u * (nlen - i))) - 1u));
    }
  }
  if (cut) { /* "/x" in front of the space: hash the shortened name on its own; NULs after the space still matter */
    for (uint32_t j = (nlen & ~3u); j < s; j += 4) {
      uint32_t z = fq_zero_bytes(fq_ldu32(d, ns + j));
      if (s - j < 4) z &= (1u << (8u * (s - j))) - 1u;
      if (z) return false;
    }
    *name_len = nlen; *mem_len_out = mem_len;
    if (want_hash) *hash = fq_hash_name_words(d, ns, nlen, seed);
    return true;
  }
  if (casava && s >= 2 && d[ns + s - 2] == '/') { /* no space at all: the "/x" rule applies to the end of the line */
    nlen = s - 2; mem_len = nlen;
    *name_len = nlen; *mem_len_out = mem_len;
    if (want_hash) *hash = fq_hash_name_words(d, ns, nlen, seed);
    return true;
  }
  *name_len = nlen; *mem_len_out = mem_len;
  if (want_hash) *hash = fq_hash_fin(h, nlen);
  return true;
}

/* Header line [h0, h0+hl) of the common shape — '@', a name byte, no NUL, LF at the end — → length of the normalised
 * name (which starts at h0+1) and the `len` fastq_get_readname reports.  False: let the careful path decide. */
FQ_HD bool fq_header_fast(const uint8_t* d, uint32_t h0, uint32_t hl, int fmt, int is_pe, uint32_t* name_len, uint64_t* mem_len_out) {
  if (hl < 3) return false;
  if (d[h0] != '@' || d[h0 + hl - 1] != '\n') return false;
  { uint8_t c1 = d[h0 + 1]; if (c1 == 0 || c1 == '\n' || c1 == '\r') return false; }
  const uint32_t ns = h0 + 1, s = hl - 1; /* rn = line + 1, strlen(rn) = s when there is no NUL */
  uint32_t nlen; uint64_t mem_len;
  if (fmt == FQ_FMT_CASAVA) { /* first space (or the end), then drop a trailing "/x" */
    uint32_t len = s;
    for (uint32_t i = 0; i < s; i += 4) {
      uint32_t w = fq_ldu32(d, ns + i);
      uint32_t z = (fq_zero_bytes(w) | fq_zero_bytes(w ^ 0x20202020u));
      if (s - i < 4) z &= (1u << (8u * (s - i))) - 1u;
      if (z) {
        uint32_t k = 0; while (!((z >> (8 * k + 7)) & 1u)) k++;
        if (((w >> (8 * k)) & 0xFFu) == 0) return false; /* NUL */
        len = i + k; break;
      }
    }
    /* bytes after the space may still hide a NUL: the careful path's strlen would stop there */
    for (uint32_t i = (len & ~3u); i < s; i += 4) {
      uint32_t z = fq_zero_bytes(fq_ldu32(d, ns + i));
      if (s - i < 4) z &= (1u << (8u * (s - i))) - 1u;
      if (z) return false;
    }
    if (len >= 2 && d[ns + len - 2] == '/') len -= 2;
    nlen = len; mem_len = len;
  } else {
    for (uint32_t i = 0; i < s; i += 4) {
      uint32_t z = fq_zero_bytes(fq_ldu32(d, ns + i));
      if (s - i < 4) z &= (1u << (8u * (s - i))) - 1u;
      if (z) return false;
    }
    if (fmt == FQ_FMT_INT) { nlen = s - 1; mem_len = s; }
    else { uint32_t len = s - (is_pe ? 1u : 0u); mem_len = len; nlen = len >= 1 ? len - 1 : s; }
  }
  *name_len = nlen; *mem_len_out = mem_len;
  return true;
}

/* Clean record of the common shape → fills *o like the careful path would (and *hash with the name hash) and returns
 * true.  Returns false without touching *o otherwise. */
FQ_HD bool fq_check_record_fast(const uint8_t* d, const FqLine* L, const FqRecCtx& cx, FqRecOut* o, uint64_t* hash) {
  const uint32_t h0 = L[0].off, hl = L[0].len, s0 = L[1].off, sl = L[1].len, p0 = L[2].off, pl = L[2].len, q0 = L[3].off, ql = L[3].len;
  if (hl < 3 || sl < 2 || ql < 2) return false;
  /* header 2: "+\n" or "+\r\n" */
  if (!(d[p0] == '+' && ((pl == 2 && d[p0 + 1] == '\n') || (pl == 3 && d[p0 + 1] == '\r' && d[p0 + 2] == '\n')))) return false;
  const uint32_t ns = h0 + 1;
  uint32_t nlen; uint64_t mem_len, hsh = FQ_HASH_SKIP;
  if (!fq_header_hash_fast(d, h0, hl, cx.fmt_key, cx.pe_key, cx.seed, cx.loop != FQ_LOOP_SINGLE, &nlen, &mem_len, &hsh)) return false;
  /* sequence */
  if (d[s0 + sl - 1] != '\n') return false;
  uint32_t se = s0 + sl - 1;
  if (se > s0 && d[se - 1] == '\r') se--;
  if (se == s0) return false;
  if (!fq_seq_fast(d, s0, se)) return false;
  /* quality */
  if (d[q0 + ql - 1] != '\n') return false;
  uint32_t qe = q0 + ql - 1;
  if (qe > q0 && d[qe - 1] == '\r') qe--;
  if (qe == q0) return false;
  uint32_t slen = se - s0, qlen = qe - q0;
  if (cx.space == FQ_SPACE_COLOR) { if (!(qlen == slen - 1 || qlen == slen)) return false; }
  else if (qlen != slen) return false;
  uint32_t qmin, qmax;
  fq_qual_fast(d, q0, qe, &qmin, &qmax);
  if (qmin <= 0x0Du) return false;
  o->flags = 0; o->vrank = FQ_V_OK; o->code = 0; o->read_len = sl; o->slen = slen; o->qlen = qlen;
  o->qmin = qmin; o->qmax = qmax; o->bad = 0; o->name_off = ns; o->name_len = nlen; o->mem_len = mem_len;
  *hash = hsh;
  return true;
}

/* Reader step flags + validator verdict + statistics + (when the record reaches the name step) the name hash. */
FQ_HD void fq_check_record(const uint8_t* d, const FqLine* L, const FqRecCtx& cx, FqRecOut* o, uint64_t* hash) {
  if (fq_check_record_fast(d, L, cx, o, hash)) return;
  fq_check_record_careful(d, L, cx, o);
  *hash = fq_record_has_name(cx.loop, *o) ? fq_hash_name(d + o->name_off, o->name_len, cx.seed) : FQ_HASH_SKIP;
}

/* event key of one record's own failure (reader flags + validator), FQ_KEY_NONE when it is clean.  g = index of
 * the record inside its file, step_base = steps that precede this file's loop (mate loop: after the index loop). */
FQ_HD uint64_t fq_record_key(int loop, uint64_t g, uint64_t step_base, const FqRecOut& o) {
  switch (loop) {
    case FQ_LOOP_INTERLEAVED: {
      uint64_t p = g >> 1;
      if ((g & 1) == 0) {
        if (o.flags & FQ_RF_STOP) return FQ_KEY(p, FQ_RI_STOP1);
        if (o.flags & FQ_RF_TRUNC) return FQ_KEY(p, FQ_RI_TRUNC1);
        if (o.flags & FQ_RF_NOTAT) return FQ_KEY(p, FQ_RI_WRONGHDR1);
        if (o.vrank != FQ_V_OK) return FQ_KEY(p, FQ_RI_V1 + o.vrank);
      } else {
        if (o.flags & FQ_RF_STOP) return FQ_KEY(p, FQ_RI_NOM2);
        if (o.flags & FQ_RF_TRUNC) return FQ_KEY(p, FQ_RI_TRUNC2);
        if (o.flags & FQ_RF_NOTAT) return FQ_KEY(p, FQ_RI_WRONGHDR2);
        if (o.vrank != FQ_V_OK) return FQ_KEY(p, FQ_RI_V2 + o.vrank);
      }
      return FQ_KEY_NONE;
    }
    case FQ_LOOP_SORTED1:
      if (o.flags & FQ_RF_STOP) return FQ_KEY(g, FQ_RS_STOP1);
      if (o.flags & FQ_RF_TRUNC) return FQ_KEY(g, FQ_RS_TRUNC1);
      if (o.vrank != FQ_V_OK) return FQ_KEY(g, FQ_RS_V1 + o.vrank);
      return FQ_KEY_NONE;
    case FQ_LOOP_SORTED2:
      if (o.flags & FQ_RF_STOP) return FQ_KEY(g, FQ_RS_STOP2);
      if (o.flags & FQ_RF_TRUNC) return FQ_KEY(g, FQ_RS_TRUNC2);
      if (o.vrank != FQ_V_OK) return FQ_KEY(g, FQ_RS_V2 + o.vrank);
      return FQ_KEY_NONE;
    case FQ_LOOP_READER: /* the reader alone: a NUL-led header ends the file, a partial record is 'file truncated'; nothing is validated */
      if (o.flags & FQ_RF_STOP) return FQ_KEY(step_base + g, FQ_R_STOP);
      if (o.flags & FQ_RF_TRUNC) return FQ_KEY(step_base + g, FQ_R_TRUNC);
      return FQ_KEY_NONE;
    case FQ_LOOP_SINGLE:
      if (o.flags & FQ_RF_STOP) return FQ_KEY(step_base + g, FQ_R_STOP);
      if (o.flags & FQ_RF_TRUNC) return FQ_KEY(step_base + g, FQ_R_TRUNC);
      if (o.vrank != FQ_V_OK) return FQ_KEY(step_base + g, FQ_R_V0 + o.vrank);
      return FQ_KEY_NONE;
    default: /* index and mate loops: "wrong header" comes from fastq_get_readname before the key step */
      if (o.flags & FQ_RF_STOP) return FQ_KEY(step_base + g, FQ_R_STOP);
      if (o.flags & FQ_RF_TRUNC) return FQ_KEY(step_base + g, FQ_R_TRUNC);
      if (o.flags & FQ_RF_NOTAT) return FQ_KEY(step_base + g, FQ_R_WRONGHDR);
      if (o.vrank != FQ_V_OK) return FQ_KEY(step_base + g, FQ_R_V0 + o.vrank);
      return FQ_KEY_NONE;
  }
}
#endif

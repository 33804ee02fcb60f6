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
 * value itself is unobservable: equality is always confirmed on the bytes (reference: hashit + strcmp). */
FQ_HD uint64_t fq_hash_mix(uint64_t h, uint32_t w) {
  h = (h ^ w) * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
FQ_HD uint64_t fq_hash_fin(uint64_t h, uint32_t len) {
  h = (h ^ len) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h >= FQ_HASH_SKIP ? h - 2 : h;
}
FQ_HD uint64_t fq_hash_name(const uint8_t* p, uint32_t len, uint32_t seed) {
  uint64_t h = 0x243F6A8885A308D3ull ^ ((uint64_t)seed * 0xFF51AFD7ED558CCDull);
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) {
    uint32_t w = (uint32_t)p[i] | ((uint32_t)p[i + 1] << 8) | ((uint32_t)p[i + 2] << 16) | ((uint32_t)p[i + 3] << 24);
    h = fq_hash_mix(h, w);
  }
  if (i < len) {
    uint32_t w = 0;
    for (uint32_t k = 0; i + k < len; k++) w |= (uint32_t)p[i + k] << (8 * k);
    h = fq_hash_mix(h, w);
  }
  return fq_hash_fin(h, len);
}
FQ_HD bool fq_bytes_equal(const uint8_t* a, const uint8_t* b, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) if (a[i] != b[i]) return false;
  return true;
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
FQ_HD void fq_check_record(const uint8_t* d, const FqLine* L, const FqRecCtx& cx, FqRecOut* o) {
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
/* does a record of this loop reach the name step (index insert / mate claim / pair compare)? */
FQ_HD bool fq_record_has_name(int loop, const FqRecOut& o) {
  if (o.flags & (FQ_RF_STOP | FQ_RF_TRUNC | FQ_RF_NOTAT)) return false;
  return loop != FQ_LOOP_SINGLE;
}

#endif

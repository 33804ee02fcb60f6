/* TEST INFRASTRUCTURE: the 16-byte fast path of fq_record.h against the careful byte-wise path on random records,
 * at every alignment.  Whenever the fast path accepts a record, every output field and the name hash must equal the
 * careful path's.  Also counts how often the fast path accepts clean Illumina-like records (must be always). */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../fastq_utils_b200/csrc/fq_record.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

int main() {
  long accepted = 0, clean_total = 0, clean_accepted = 0, bad = 0;
  const char* names[] = {"r%u/1", "M01:5:FC:1:1101:%u:2000 1:N:0:ACGT", "%u", "read%u_x extra words", "S4_01:4:1:%u:16/1 1:Y:0:0", "ab%u:1", "A:1/2:%u 2:Y:0:A"};
  for (int it = 0; it < 400000; it++) {
    std::string rec;
    char nm[128]; snprintf(nm, sizeof nm, names[rnd() % 7], rnd() % 100000);
    uint32_t L = 1 + rnd() % 70;
    std::string seq, qual;
    for (uint32_t i = 0; i < L; i++) { seq += "ACGTNacgtn"[rnd() % 10]; qual += (char)(33 + rnd() % 60); }
    bool crlf = rnd() % 16 == 0;
    const char* nl = crlf ? "\r\n" : "\n";
    int mut = rnd() % 24; /* most records stay clean */
    bool clean = mut >= 12;
    if (mut == 0) seq[rnd() % L] = "XU.0*\r\0"[rnd() % 7];
    if (mut == 1) qual[rnd() % L] = (char)(rnd() % 16);
    if (mut == 2) qual += "I";
    if (mut == 3) nm[rnd() % strlen(nm)] = 0;
    if (mut == 4) qual[rnd() % L] = (char)(0x80 + rnd() % 128);
    if (mut == 5 && L > 1) qual.erase(0, 1);
    std::string plus = "+";
    if (mut == 6) plus += nm;
    if (mut == 7) plus = "-";
    if (mut == 8) seq.clear();
    std::string hdr = std::string(mut == 9 ? "" : "@") + nm;
    if (mut == 10) hdr = "@";
    if (mut == 11) { clean = true; } /* high-bit qualities are clean too when mut==4? no: keep 11 as a plain clean case */
    uint32_t pad = rnd() % 16;
    std::vector<uint8_t> buf(pad, 'Z');
    uint32_t off[5];
    const std::string parts[4] = {hdr, seq, plus, qual};
    for (int i = 0; i < 4; i++) {
      off[i] = (uint32_t)buf.size();
      buf.insert(buf.end(), parts[i].begin(), parts[i].end());
      if (mut == 3 && i == 0) { /* the NUL written into nm cut the C string: re-add the tail as raw bytes */ }
      buf.insert(buf.end(), nl, nl + strlen(nl));
    }
    off[4] = (uint32_t)buf.size();
    buf.resize(buf.size() + 64, 'Q');
    FqLine Ls[4];
    for (int i = 0; i < 4; i++) { Ls[i].off = off[i]; Ls[i].len = off[i + 1] - off[i]; }
    for (int fmt = 0; fmt < 3; fmt++) for (int pe = 0; pe < 2; pe++) for (int sp = 0; sp < 2; sp++) {
      FqRecCtx cx; memset(&cx, 0, sizeof cx);
      cx.loop = FQ_LOOP_INDEX; cx.fmt_key = cx.fmt_val = fmt; cx.pe_key = cx.pe_val = pe; cx.space = sp; cx.weight = 1; cx.seed = 3;
      FqRecOut a, b; uint64_t ha = 0;
      memset(&a, 0xEE, sizeof a); memset(&b, 0xEE, sizeof b);
      fq_check_record_careful(buf.data(), Ls, cx, &b);
      uint64_t hb = fq_record_has_name(cx.loop, b) ? fq_hash_name(buf.data() + b.name_off, b.name_len, cx.seed) : FQ_HASH_SKIP;
      bool ok = fq_check_record_fast(buf.data(), Ls, cx, &a, &ha);
      bool b_clean = b.flags == 0 && b.vrank == FQ_V_OK;
      if (clean && sp == 0) { clean_total++; clean_accepted += ok; }
      if (ok) {
        accepted++;
        bool same = b_clean && a.read_len == b.read_len && a.slen == b.slen && a.qlen == b.qlen && a.qmin == b.qmin && a.qmax == b.qmax &&
                    a.name_off == b.name_off && a.name_len == b.name_len && a.mem_len == b.mem_len && ha == hb && a.flags == b.flags && a.code == b.code;
        if (!same) { if (bad++ < 5) fprintf(stderr, "MISMATCH it=%d mut=%d fmt=%d pe=%d sp=%d name_len %u/%u hash %llx/%llx q %u-%u / %u-%u\n", it, mut, fmt, pe, sp, a.name_len, b.name_len, (unsigned long long)ha, (unsigned long long)hb, a.qmin, a.qmax, b.qmin, b.qmax); }
      }
    }
  }
  printf("accepted=%ld clean=%ld clean_accepted=%ld mismatches=%ld\n", accepted, clean_total, clean_accepted, bad);
  return bad == 0 && clean_accepted == clean_total ? 0 : 1;
}

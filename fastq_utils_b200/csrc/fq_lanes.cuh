/*
 * fq_lanes.cuh — the clean-data pass of libfastq_gpu (included by fq_cuda.cu inside its anonymous namespace).
 *
 *   K5  fq_lanes_kernel          one pass over HBM; every phase is parallel over 16-byte chunks or over lines, never over records
 *   K5r fq_lanes_records_kernel  per-record length rules + statistics from the line index the pass wrote (20 B per record)
 *   K5c fq_lanes_commit_kernel   folds the staged minima / maxima of an accepted chunk into the file statistics
 *
 * What it replaces: src/fastq.c:245-261 (4×gzgets splitter), :300-392 (validator), :97-110 (statistics), :442-516 (names).
 *
 * The pass never decides an error.  It proves that a chunk is CLEAN — header lines "@name...\n" without NUL, sequence lines
 * made of ACGTN/acgtn + LF, "+\n", quality bytes above 0x0D + LF, equal lengths, no over-long line — and for clean chunks
 * produces exactly what the per-record kernels produce (line index, name descriptors, statistics).  Anything else raises an
 * anomaly bit in out[3] / out[10]; nothing of the chunk is committed then and the engine hands the chunk to the per-record
 * kernels (fq_tile_kernel / K1+K2), which own the reference's first-error semantics.
 *
 * Per 32 KiB tile (persistent CTAs claim tiles in order; one bulk async copy brings tile + margin into shared memory):
 *   A  LF flags of every 16-byte chunk (3 integer ops per word, IDP.4A gathers the 16 flag bits in byte order)
 *   B  block prefix of the LF counts + decoupled look-back over tiles → global line number of every chunk → its line class
 *      (line number mod 4: header, sequence, plus, quality)
 *   C  each thread walks its 8 consecutive chunks: line ends → shared list; sequence / quality byte ranges → four work lists
 *      (whole chunks as 16-bit ids, partial chunks with their byte range)
 *   D  the work lists are consumed class by class, so a warp executes ONE predicate for 32 chunks: alphabet (bit-sliced LOP3,
 *      shifts as IMAD on the FMA pipe) or quality minimum / maximum (VIMNMX.U16x2)
 *   E  line ends → global line index (coalesced); header lines → '@' syntax, name slice, 64-bit hash → name descriptors;
 *      plus lines → "+\n"
 */

constexpr int LN_THREADS = 256, LN_WARPS = LN_THREADS / 32;
#ifndef LN_MAXREG
#define LN_MAXREG 56 /* 4 CTAs of 256 threads leave 8 K registers per SM: room for a block of the index kernel beside the pass */
#endif
constexpr int LN_TILE = 32768, LN_CHUNKS = LN_TILE / 16, LN_CPT = LN_CHUNKS / LN_THREADS; /* 8 chunks per thread */
constexpr int LN_LEFT = 16, LN_MARGIN = 1024, LN_WIN = LN_LEFT + LN_TILE + LN_MARGIN;
/* per-line mode: the LF scan covers tile + margin (a line is judged by the tile it starts in, whole), so the tile is a margin shorter */
constexpr int LS_TILE = LN_TILE - LN_MARGIN, LS_TILE_CHUNKS = LS_TILE / 16;
constexpr int LN_LMAX = 2048;                 /* line ends kept per tile */
constexpr int LN_SMAX = 256;                  /* header lines (name descriptors) staged per tile */
constexpr int LN_OFF_MASK = LN_WIN;
constexpr int LN_OFF_PURE = LN_OFF_MASK + LN_CHUNKS * 2;
constexpr int LN_OFF_LEND = LN_OFF_PURE + LN_CHUNKS * 2;      /* chunk-parallel mode: two buffers (the line ends of a tile are written out one round later);
                                                                 per-line mode: one buffer, then the CTA's read-length histogram (lengths below LS_HBINS) */
constexpr int LS_HBINS = LN_LMAX * 2 / 4;                     /* 1024 bins of 32 bits in the second buffer's place */
constexpr int LS_STAGE_HIST = 36864;                          /* bins of the chunk's staged histogram: a line of the per-line mode lies inside one window */
constexpr int LN_OFF_STAGE = LN_OFF_LEND + 2 * LN_LMAX * 2;
constexpr int LN_OFF_LUT = LN_OFF_STAGE + LN_SMAX * 16;
constexpr int LN_SMEM = LN_OFF_LUT + 32 * 16;
static_assert(LN_CPT == 8, "a thread's masks are one 16-byte load");
static_assert(LN_SMAX <= LN_THREADS, "one staged name per thread");
static_assert(LN_WIN % 16 == 0 && LN_OFF_LEND % 16 == 0 && LN_OFF_PURE % 16 == 0 && LN_OFF_STAGE % 16 == 0 && LN_OFF_LUT % 16 == 0, "alignment");

/* anomaly bits (out[3]) */
enum { LN_A_BASE = 1, LN_A_QUAL = 2, LN_A_HEADER = 4, LN_A_PLUS = 8, LN_A_CAPACITY = 16, LN_A_PHASE = 32 };
/* what a per-line pass found, staged per chunk (zeroed before the launch) and folded into the file statistics by fq_lanes_post_kernel
 * only when the WHOLE chunk turned out clean: nothing of a chunk that is handed on to the per-record kernels is ever counted */
struct LanesStage {
  unsigned long long nrec;       /* complete records that keep the length rules */
  unsigned long long mem;        /* sum of the `len` fastq_get_readname reports for them (index_mem, src/fastq.c:609) */
  unsigned int rl_min_inv, rl_max; /* ~minimum and maximum of the read length (terminator included) */
  unsigned int arena_used;       /* 16-byte units of the name arena the fullest tile asked for */
  unsigned int pad;
  unsigned int hist[LS_STAGE_HIST];
};
/* out words */
enum { LN_O_ARENA_USED = 24, LN_O_ACCEPT = 25, LN_O_INDEX_FROM = 26, LN_O_STAGED = 27 };
enum { LN_O_LINES = 0, LN_O_CAPOVF = 1, LN_O_OVERLONG = 2, LN_O_ANOMALY = 3, LN_O_INTERNAL = 4, LN_O_VIRTUAL = 5, LN_O_QMIN = 6, LN_O_QMAX = 7,
       LN_O_RLMIN = 8, LN_O_RLMAX = 9, LN_O_RECBAD = 10, LN_O_SPINS = 11 /* diagnostics: polls of unpublished tile states */,
       LN_O_ROUNDS = 12 /* look-back rounds */, LN_O_WAITED = 13 /* tiles that had to wait for the sum before their scans */, LN_O_WORDS = 16 };

struct LanesParams {
  const uint8_t* data; /* 16-byte aligned (bulk copies) */
  uint32_t lead;       /* the first `lead` (< 16) bytes are not the chunk's own: its first line starts at data[lead] */
  uint32_t n; int virtual_end; uint32_t* line_end; uint32_t cap;
  unsigned long long* tile_state; uint32_t* ticket; uint32_t ntiles;
  uint32_t* out;
  uint32_t j0; FqRecCtx cx; FqName* names; uint32_t names_cap;
  LanesStage* stage;   /* per-line mode */
  uint8_t* arena; uint32_t arena_units; /* per-line mode: where the names go (16-byte units); NULL: the names stay in the chunk */
  /* per-line mode, sharded runs: every name goes straight into the region of the rank that owns its hash (fq_device.h: FqRegionHdr);
   * this CTA writes stretch blockIdx.x of every region */
  uint32_t route_world, route_stride, route_units; uint8_t* route_region[FQ_ROUTE_MAX_WORLD];
  uint32_t tune; /* experiment switch (FQG_LANES_TUNE): 1 = no L2 prefetch of the next tile */
};

/* 0x80 in every byte of x that is LF (exact): three integer instructions */
__device__ __forceinline__ uint32_t ln_lf_flags(uint32_t x) {
  uint32_t t = (x ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
  uint32_t s = t + 0x7F7F7F7Fu;
  return ~(s | x) & 0x80808080u;
}
/* 16 LF flags of a chunk in byte order: the flag bytes (0x80) are weighted 1,2,4,...,128 by two dot products per half */
__device__ __forceinline__ uint32_t ln_lf_mask16(uint4 v) {
  uint32_t lo = __dp4a(ln_lf_flags(v.x), 0x08040201u, 0u);
  lo = __dp4a(ln_lf_flags(v.y), 0x80402010u, lo);
  uint32_t hi = __dp4a(ln_lf_flags(v.z), 0x08040201u, 0u);
  hi = __dp4a(ln_lf_flags(v.w), 0x80402010u, hi);
  return (hi * 256u + lo) >> 7;
}
/* ACGTN / acgtn in bit 7 of every byte (see fq_base_pred): 7 IMAD shifts + 8 LOP3 */
__device__ __forceinline__ uint32_t ln_pred4(uint4 v) {
  return fq_base_pred(v.x) & fq_base_pred(v.y) & fq_base_pred(v.z) & fq_base_pred(v.w);
}
__device__ __forceinline__ void ln_minmax_word(uint32_t w, uint32_t& mn, uint32_t& mx) {
  uint32_t ev = w & 0x00FF00FFu, od = __byte_perm(w, 0u, 0x4341u);
  mn = __vminu2(mn, __vminu2(ev, od));
  mx = __vmaxu2(mx, __vmaxu2(ev, od));
}

/* Alphabet check by reconstruction: bits 1..3 of a byte tell which of A C G T N it could be ((b >> 1) & 7 = 0 1 3 2 7); the byte is
 * valid iff it EQUALS that letter.  The letter is rebuilt arithmetically in all four byte lanes at once,
 *     0x41 + 2 c0 + 0x13 c1 - 0x0F c0 c1 + 7 c0 c1 c2      (A C T G . . . N for codes 0..7; codes 4..6 rebuild a letter of another code)
 * with the right shifts as IMAD.HI and the sums as IMAD: 5 LOP3 on the ALU pipe, 7 on the FMA pipe per word (the bit-sliced
 * predicate needs 8 + 6).  Result: 0 iff all four bytes are one of ACGTN (upper case). */
__device__ __forceinline__ uint32_t ln_diff(uint32_t w) {
  uint32_t t1, t2, t3;
  asm("mul.hi.u32 %0, %1, 0x80000000;" : "=r"(t1) : "r"(w)); /* w >> 1 on the FMA pipe */
  asm("mul.hi.u32 %0, %1, 0x40000000;" : "=r"(t2) : "r"(w));
  asm("mul.hi.u32 %0, %1, 0x20000000;" : "=r"(t3) : "r"(w));
  const uint32_t c0 = t1 & 0x01010101u, c1 = t2 & 0x01010101u, c01 = t1 & t2 & 0x01010101u, c012 = c01 & t3;
  uint32_t e = 0x41414141u + 2u * c0; e += 0x13u * c1; e -= 0x0Fu * c01; e += 7u * c012;
  return w ^ e;
}
__device__ __forceinline__ uint32_t ln_diff4(uint4 v) { return ln_diff(v.x) | ln_diff(v.y) | ln_diff(v.z) | ln_diff(v.w); }
/* per-line mode: every byte of [s, e) (window offsets, e > s) is one of ACGTN → 0 */
__device__ __forceinline__ uint32_t ls_seq_line(const uint8_t* win, const uint4* lut, uint32_t s, uint32_t e) {
  uint32_t a = s & ~15u;
  const uint32_t alast = (e - 1u) & ~15u;
  uint4 m = lut[s & 15u];
  if (a == alast) { const uint4 h = lut[15u + (e - a)]; m.x &= h.x; m.y &= h.y; m.z &= h.z; m.w &= h.w; }
  uint4 v = *(const uint4*)(win + a);
  uint32_t bad = (ln_diff(v.x) & m.x) | (ln_diff(v.y) & m.y) | (ln_diff(v.z) & m.z) | (ln_diff(v.w) & m.w);
  if (a != alast) {
    for (a += 16; a < alast; a += 16) bad |= ln_diff4(*(const uint4*)(win + a));
    m = lut[15u + (e - alast)];
    v = *(const uint4*)(win + alast);
    bad |= (ln_diff(v.x) & m.x) | (ln_diff(v.y) & m.y) | (ln_diff(v.z) & m.z) | (ln_diff(v.w) & m.w);
  }
  return bad;
}
/* per-line mode: unsigned minimum / maximum over the bytes of [s, e) folded into two 16-bit lanes each */
__device__ __forceinline__ void ls_qual_line(const uint8_t* win, const uint4* lut, uint32_t s, uint32_t e, uint32_t& mn, uint32_t& mx) {
  uint32_t a = s & ~15u;
  const uint32_t alast = (e - 1u) & ~15u;
  const uint32_t fill = (uint32_t)win[s] * 0x01010101u; /* a byte of the line stands in for the bytes outside it */
  uint4 m = lut[s & 15u];
  if (a == alast) { const uint4 h = lut[15u + (e - a)]; m.x &= h.x; m.y &= h.y; m.z &= h.z; m.w &= h.w; }
  uint4 v = *(const uint4*)(win + a);
  ln_minmax_word((v.x & m.x) | (fill & ~m.x), mn, mx); ln_minmax_word((v.y & m.y) | (fill & ~m.y), mn, mx);
  ln_minmax_word((v.z & m.z) | (fill & ~m.z), mn, mx); ln_minmax_word((v.w & m.w) | (fill & ~m.w), mn, mx);
  if (a != alast) {
    for (a += 16; a < alast; a += 16) {
      v = *(const uint4*)(win + a);
      ln_minmax_word(v.x, mn, mx); ln_minmax_word(v.y, mn, mx); ln_minmax_word(v.z, mn, mx); ln_minmax_word(v.w, mn, mx);
    }
    m = lut[15u + (e - alast)];
    v = *(const uint4*)(win + alast);
    ln_minmax_word((v.x & m.x) | (fill & ~m.x), mn, mx); ln_minmax_word((v.y & m.y) | (fill & ~m.y), mn, mx);
    ln_minmax_word((v.z & m.z) | (fill & ~m.z), mn, mx); ln_minmax_word((v.w & m.w) | (fill & ~m.w), mn, mx);
  }
}

/* Block-wide look-back: the number of lines in front of `tile` = the counts of the tiles before it, summed back to the nearest
 * one whose inclusive count is known.  Every thread reads the states of four predecessors per round (1024 per round: hundreds of
 * tiles are in flight under persistent CTAs, and a round costs an L2 round trip plus a barrier).  Uniform over the block. */
/* an anomaly was raised somewhere: the chunk goes to the per-record kernels whatever this pass still does, so everybody winds down */
__device__ __forceinline__ bool ln_chunk_lost(const uint32_t* out) {
  uint32_t a, o;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(a) : "l"(out + LN_O_ANOMALY));
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(o) : "l"(out + LN_O_OVERLONG));
  return a != 0 || o != 0xFFFFFFFFu;
}
__device__ __forceinline__ unsigned long long ln_wait_state(const unsigned long long* p, uint32_t* out) {
  unsigned long long v64;
  uint32_t spins = 0;
  while (((v64 = ld_volatile64(p)) >> 62) == 0) {
    if (++spins > (1u << 24)) { atomicExch(out + LN_O_INTERNAL, 2u); return ST_INCL; }
    if ((spins & 63u) == 0 && ln_chunk_lost(out)) return ST_INCL; /* the tile we wait for may never be scanned: CTAs stop claiming tiles */
  }
  if (spins) atomicAdd(out + LN_O_SPINS, spins);
  return v64;
}
__device__ __forceinline__ bool ln_lookback_reduce(uint32_t part, bool has, int lane, int warp, uint32_t* s_sum, uint32_t* s_has, uint32_t* base, uint32_t* out) {
  const uint32_t incl_mask = __ballot_sync(FULL, has);
  const int first = incl_mask ? __ffs(incl_mask) - 1 : 31;
  part = __reduce_add_sync(FULL, lane <= first ? part : 0u);
  if (lane == 0) { s_sum[warp] = part; s_has[warp] = incl_mask ? 1u : 0u; }
  if (lane == 0 && warp == 0) atomicAdd(out + LN_O_ROUNDS, 1u);
  __syncthreads();
  bool found = false;
#pragma unroll
  for (int w = 0; w < LN_WARPS; w++) if (!found) { *base += s_sum[w]; found = s_has[w] != 0; }
  return found;
}
/* `pre` = state of tile (tile - 1 - tid) read earlier (0 when it was not published yet, or not read): the first round looks at
 * the 256 tiles right in front; only if none of them has its inclusive count do rounds of 1024 (four independent loads per thread)
 * follow. */
__device__ __forceinline__ uint32_t ln_lookback(const unsigned long long* tile_state, uint32_t tile, unsigned long long pre, int tid, int lane, int warp,
                                                uint32_t* s_sum, uint32_t* s_has, uint32_t* out) {
  uint32_t base = 0;
  {
    const int idx = (int)tile - 1 - tid;
    unsigned long long v64 = ST_INCL;
    if (idx >= 0) v64 = (pre >> 62) != 0 ? pre : ln_wait_state(tile_state + idx, out);
    if (ln_lookback_reduce((uint32_t)(v64 & ST_VALUE), (v64 >> 62) == 2, lane, warp, s_sum, s_has, &base, out)) return base;
    __syncthreads(); /* s_sum / s_has are rewritten by the next round */
  }
  for (int look = (int)tile - 1 - LN_THREADS;; look -= 4 * LN_THREADS) {
    unsigned long long v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { const int idx = look - 4 * tid - j; v[j] = idx >= 0 ? ld_volatile64(tile_state + idx) : ST_INCL; }
    uint32_t part = 0; bool has = false;
#pragma unroll
    for (int j = 0; j < 4; j++) { /* nearest first */
      if ((v[j] >> 62) == 0) v[j] = ln_wait_state(tile_state + (look - 4 * tid - j), out);
      if (!has) { part += (uint32_t)(v[j] & ST_VALUE); has = (v[j] >> 62) == 2; }
    }
    if (ln_lookback_reduce(part, has, lane, warp, s_sum, s_has, &base, out)) break;
    __syncthreads();
  }
  return base;
}

/* LINES = false: chunk-parallel everywhere (lines of any length).  LINES = true: one thread per line for the bulk scans, warps of
 * one line class each — a third of the instructions when lines are short (they must end within the margin: < 1 KiB). */
template <bool LINES>
__global__ void __maxnreg__(LN_MAXREG)
fq_lanes_kernel(const LanesParams P) {
  constexpr int TILE = LINES ? LS_TILE : LN_TILE;          /* bytes a tile owns */
  constexpr int SCAN = LN_TILE;                            /* bytes whose LFs are flagged: the tile, and in per-line mode the margin too */
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* win = smem;
  uint16_t* maskbuf = (uint16_t*)(smem + LN_OFF_MASK);
  uint16_t* pure = (uint16_t*)(smem + LN_OFF_PURE);   /* whole chunks, one region of 256 entries per warp: sequence from its front, quality from its back */
  uint16_t* lend2 = (uint16_t*)(smem + LN_OFF_LEND);  /* line ends (window offsets), [2][LN_LMAX] */
  FqName* stage = (FqName*)(smem + LN_OFF_STAGE);     /* name descriptors of the tile's header lines, written out one round later */
  uint4* lut = (uint4*)(smem + LN_OFF_LUT);           /* [lo] bytes >= lo, [16 + h] bytes <= h */
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_next, s_fbase, s_fstate, s_w1[LN_WARPS], s_w3[LN_WARPS], s_w4[LN_WARPS];
  __shared__ uint32_t s_cnt, s_mem, s_rlmin, s_rlmax;  /* per-line mode: this CTA's records, index_mem bytes, read lengths */
  __shared__ uint32_t s_arena;                           /* per-line mode: arena units handed out to the names of the current tile */
  __shared__ uint32_t s_own[FQ_ROUTE_MAX_WORLD];         /* routing: names this CTA has written for each owner so far */
  __shared__ uint8_t* s_reg[FQ_ROUTE_MAX_WORLD];         /* ... and where this CTA's stretch of the owner's region starts */
  uint32_t* s_hist = (uint32_t*)(smem + LN_OFF_LEND + LN_LMAX * 2); /* per-line mode: records by read length (below LS_HBINS) */
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_next = atomicAdd(P.ticket, 1u);
    s_cnt = 0; s_mem = 0; s_rlmin = 0xFFFFFFFFu; s_rlmax = 0; s_arena = 0;
  }
  /* every tile has its own stretch of the arena (no global cursor: an atomic with a return value in a header group's path costs a
   * round trip to L2, and several when the index kernel beside the pass keeps the atomic units busy) */
  const uint32_t arena_stride = P.arena ? P.arena_units / P.ntiles : 0u;
  uint32_t arena_max = 0;
  uint32_t route_tile_max = 0; /* (thread 0) the most header lines a tile of this CTA has held so far */
  if (LINES) for (int l = threadIdx.x; l < LS_HBINS; l += LN_THREADS) s_hist[l] = 0;
  const uint32_t route_slot_bytes = 16u + 16u * P.route_units;
  if (LINES && P.route_world && tid < (int)P.route_world) {
    s_own[tid] = 0;
    s_reg[tid] = P.route_region[tid] + 16 + fq_route_counts_bytes(gridDim.x) + (size_t)blockIdx.x * P.route_stride * route_slot_bytes;
  }
  if (tid < 32) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t x = 0;
      for (int b = 0; b < 4; b++) {
        int byte = 4 * k + b;
        bool on = tid < 16 ? byte >= tid : byte <= tid - 16;
        if (on) x |= 0xFFu << (8 * b);
      }
      w[k] = x;
    }
    lut[tid] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  uint32_t parity = 0, buf = 0;
  uint32_t seq_ok = 0x80808080u;           /* AND of the alphabet predicate over everything this thread checked (chunk-parallel mode) */
  uint32_t seq_bad = 0;                    /* OR of the alphabet differences (per-line mode) */
  uint32_t qmn = 0x00FF00FFu, qmx = 0u;    /* quality minimum / maximum, two 16-bit lanes */
  uint32_t anomaly = 0, told = 0;
  /* the tile of the previous round: its line ends and names still wait for the number of lines in front of it */
  bool pend = false, p_have_base = false, p_no_final_lf = false;
  uint32_t p_tile = 0, p_cnt = 0, p_phi = 0, p_base = 0, p_nstage = 0, p_rl0 = 0;

  for (;;) {
    __syncthreads(); /* everyone is done with the previous window and lists; s_next holds the tile claimed for this round */
    const uint32_t tile = s_next;
    const bool active = tile < P.ntiles;
    const unsigned long long t0 = (unsigned long long)tile * TILE;
    if (LINES && tid == 0) { arena_max = max(arena_max, s_arena); s_arena = 0; } /* (the barrier of phase B lies between this and the header groups) */
    if (active && tid == 0) {
      unsigned long long src = tile ? t0 - LN_LEFT : 0;
      uint32_t dst_off = tile ? 0 : LN_LEFT;
      unsigned long long want = (unsigned long long)(LN_LEFT + TILE + LN_MARGIN) - dst_off, have = ((unsigned long long)P.n - src + 15) & ~15ull;
      uint32_t bytes = (uint32_t)(want < have ? want : have);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(win + dst_off)), "l"(P.data + src), "r"(bytes), "r"(bar) : "memory");
    }

    /* has the chunk been lost already (asked for now, looked at when the next tile is claimed: no waiting for the answer) */
    uint32_t lost_a = 0, lost_o = 0xFFFFFFFFu;
    if (tid == 0) {
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(lost_a) : "l"(P.out + LN_O_ANOMALY));
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(lost_o) : "l"(P.out + LN_O_OVERLONG));
    }
    /* the state of one tile in front of the previous tile: asked for now, looked at in F */
    unsigned long long pre64 = 0;
    if (pend && !p_have_base && (int)p_tile - 1 - tid >= 0) pre64 = ld_volatile64(P.tile_state + ((int)p_tile - 1 - tid));

    /* ---- F: finish the previous tile: lines in front of it (its predecessors published their counts a round ago, and nothing that
     * can wait runs before a tile publishes its own), the check of the line class we assumed, line ends and names → global memory */
    /* the staged name this thread writes out: taken now, before this round stages its own */
    FqName my_nm; my_nm.len = 0xFFFFFFFFu; my_nm.off = 0; my_nm.hash = 0;
    if (pend && (uint32_t)tid < p_nstage) my_nm = stage[tid];
    auto finish_previous = [&](uint32_t base, bool looked_up) {
      if (looked_up) {
        if (((base + 4u - P.j0) & 3u) != p_phi) anomaly |= LN_A_PHASE; /* the plus lines of the tile misled us: hand the chunk on */
        if (tid == 0) st_volatile64(P.tile_state + p_tile, ST_INCL | ((unsigned long long)base + p_cnt));
      }
      if (tid == 0 && p_tile == P.ntiles - 1) {
        uint32_t cnt = base + p_cnt;
        if (p_no_final_lf) { if (cnt < P.cap) P.line_end[cnt] = P.n; P.out[LN_O_VIRTUAL] = cnt; cnt++; }
        P.out[LN_O_LINES] = cnt; P.out[LN_O_CAPOVF] = cnt > P.cap ? 1u : 0u;
      }
      if (!LINES) { /* (the per-line mode judges the records itself: nobody reads its line index but the host, who wants the chunk's first and last lines) */
        const uint16_t* pl = lend2 + (buf ^ 1u) * LN_LMAX;
        const uint32_t gofs = (uint32_t)((unsigned long long)p_tile * TILE - LN_LEFT); /* window offset → offset inside the chunk */
        if (p_cnt <= (uint32_t)LN_LMAX)
          for (uint32_t r = tid; r < p_cnt; r += LN_THREADS) { const uint32_t gi = base + r; if (gi < P.cap) P.line_end[gi] = gofs + pl[r]; }
      }
      if (P.names) {
        /* record of a staged name: its number inside the tile plus the records in front of the tile; names of the record cut
         * by the start of the chunk (lines before j0) get a negative number and are dropped */
        const uint32_t rec0 = p_rl0 + ((base + 4u - P.j0) >> 2) - 2u;
        const uint32_t rec = rec0 + tid; /* p_nstage <= LN_SMAX <= LN_THREADS: one name per thread */
        if (my_nm.len != 0xFFFFFFFFu && rec < P.names_cap) P.names[rec] = my_nm;
      }
      pend = false;
    };
    if (!active) { /* no tile left for this CTA: finish the last one and leave */
      if (pend) {
        uint32_t base = p_base;
        if (!p_have_base) base = ln_lookback(P.tile_state, p_tile, pre64, tid, lane, warp, s_w3, s_w4, P.out);
        finish_previous(base, !p_have_base);
      }
      break;
    }

    const uint32_t left = (uint32_t)min((unsigned long long)(TILE + LN_MARGIN), (unsigned long long)P.n - t0); /* data bytes from the tile start */
    const uint32_t nv = min(left, (uint32_t)TILE);      /* valid bytes of the tile itself */
    const uint32_t ns = min(left, (uint32_t)SCAN);      /* valid bytes of the scanned range */
    const uint32_t nloc = LN_LEFT + left;               /* window offsets below this hold data */
    const bool full = ns == (uint32_t)SCAN;
    uint16_t* lend = LINES ? lend2 : lend2 + buf * LN_LMAX;
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) { if (tid == 0) atomicExch(P.out + LN_O_INTERNAL, 1u); break; } }
      parity ^= 1;
    }
    if (tile == 0 && tid == 0) win[LN_LEFT - 1] = '\n';
    const uint32_t lead0 = tile == 0 ? P.lead : 0u; /* line 0 of the first tile starts behind the lead-in */

    /* ---- A: LF flags; lane ↔ adjacent chunks (conflict-free 128-bit shared loads) */
    {
      const uint32_t cbase = warp * (LN_CHUNKS / LN_WARPS) + lane;
#pragma unroll
      for (int i = 0; i < LN_CPT; i++) {
        const uint32_t c = cbase + i * 32;
        uint32_t m = ln_lf_mask16(*(const uint4*)(win + LN_LEFT + 16 * c));
        if (!full) { uint32_t valid = ns > 16 * c ? min(16u, ns - 16 * c) : 0u; m &= (1u << valid) - 1u; }
        if (c == 0) m &= ~((1u << lead0) - 1u);
        maskbuf[c] = (uint16_t)m;
      }
    }
    __syncwarp(); /* a thread's 8 consecutive chunks were flagged by its own warp */
    const uint4 mm = *(const uint4*)(maskbuf + LN_CPT * tid);
    /* LFs in front of each of the thread's chunks, four 8-bit counters per register */
    uint32_t cumA, cumB, tot;
    {
      const uint32_t n0 = __popc(mm.x & 0xFFFFu), n01 = __popc(mm.x), n2 = __popc(mm.y & 0xFFFFu), n23 = __popc(mm.y);
      const uint32_t n4 = __popc(mm.z & 0xFFFFu), n45 = __popc(mm.z), n6 = __popc(mm.w & 0xFFFFu), n67 = __popc(mm.w);
      const uint32_t h = n01 + n23;
      cumA = (n0 << 8) | (n01 << 16) | ((n01 + n2) << 24);                       /* chunks 0..3: 0, n0, n01, n01+n2 */
      cumB = h | ((h + n4) << 8) | ((h + n45) << 16) | ((h + n45 + n6) << 24);   /* chunks 4..7 */
      tot = h + n45 + n67;
    }

    /* ---- B: prefix of the LF counts inside the tile; the tile's count is published for the tiles behind us; line ends.
     * Low half: LFs of the scanned range (ranks of the line ends); high half: only those of the tile itself (its published count). */
    const uint32_t c0 = LN_CPT * tid;
    const uint32_t tot2 = tot | ((!LINES || c0 < (uint32_t)LS_TILE_CHUNKS) ? tot << 16 : 0u);
    uint32_t incl = tot2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) s_w1[warp] = incl;
    /* F, first half: the states of the 256 tiles in front of the previous tile were asked for at the top of the round; each warp
     * reduces its 32 now and the barrier of the prefix carries the partial sums along (no barrier of its own) */
    const bool f_lookup = pend && !p_have_base;
    if (f_lookup) {
      const bool ahead = (int)p_tile - 1 - tid >= 0;
      const unsigned long long v64 = ahead ? pre64 : ST_INCL;
      const uint32_t unpub = __ballot_sync(FULL, (v64 >> 62) == 0);
      const uint32_t incl_mask = __ballot_sync(FULL, (v64 >> 62) == 2);
      const int first = incl_mask ? __ffs(incl_mask) - 1 : 31;
      const uint32_t part = __reduce_add_sync(FULL, lane <= first ? (uint32_t)(v64 & ST_VALUE) : 0u);
      /* 1: an inclusive count is among the 32; 2: a state in front of it was not published yet (read it again, rarely needed) */
      const uint32_t lowmask = first == 31 ? FULL : (2u << first) - 1u;
      if (lane == 0) { s_w3[warp] = part; s_w4[warp] = (unpub & lowmask) ? 2u : incl_mask ? 1u : 0u; }
    }
    __syncthreads();
    uint32_t excl = incl - tot2, cnt2 = 0;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) { uint32_t x = s_w1[w]; cnt2 += x; if (w < warp) excl += x; }
    excl &= 0xFFFFu;
    const uint32_t cntW = cnt2 & 0xFFFFu, cntT = cnt2 >> 16; /* line ends inside the scanned range / inside the tile */
    if (f_lookup && tid == 32) { /* one thread folds the warps' look-back partials; everybody reads the result after the next barrier */
      uint32_t state = 0, base = 0;
#pragma unroll
      for (int w = 0; w < LN_WARPS; w++) if (state == 0) { base += s_w3[w]; state = s_w4[w]; }
      s_fbase = base; s_fstate = state;
    }
    if (tid == 0) {
      if (tile > 0) st_volatile64(P.tile_state + tile, ST_AGG | cntT);
      /* Per-line mode claims the next tile now (and pulls it into L2).  The chunk-parallel mode claims at the end of the round: many
       * of its tiles (long lines: no plus line inside) wait for the counts of the tiles in front, and a tile claimed a round before
       * it is scanned would keep every tile behind it waiting that long. */
      /* routing: this CTA's stretch of an owner's region must hold the names of the tile in hand and of the next one — otherwise the
       * CTA stops claiming tiles (the others take them: the stretches hold a third more than the chunk's names) */
      bool route_full = false;
      if (LINES && P.route_world) {
        uint32_t most = 0;
        for (uint32_t o = 0; o < P.route_world; o++) most = max(most, s_own[o]);
        route_full = most + 2u * (min((uint32_t)LN_SMAX, max(route_tile_max, 64u)) / P.route_world + 16u) > P.route_stride; /* (an owner's share of a tile's names, generously) */
      }
      const uint32_t nxt = LINES ? ((lost_a != 0 || lost_o != 0xFFFFFFFFu || route_full) ? 0xFFFFFFFFu : atomicAdd(P.ticket, 1u)) : 0xFFFFFFFFu;
      if (LINES) s_next = nxt; /* read after the barrier at the top of the next round */
      if (LINES && nxt < P.ntiles && !(P.tune & 1u)) { /* pull the next tile into L2 now: its bulk copy, issued when this round is over, then finds it there */
        const unsigned long long src = (unsigned long long)nxt * TILE - LN_LEFT;
        const unsigned long long have = ((unsigned long long)P.n - src + 15) & ~15ull;
        const uint32_t bytes = (uint32_t)(have < (unsigned long long)(LN_LEFT + TILE + LN_MARGIN) ? have : (unsigned long long)(LN_LEFT + TILE + LN_MARGIN));
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.data + src), "r"(bytes) : "memory");
      }
    }
    const bool too_many = cntW > (uint32_t)LN_LMAX;
    if (too_many) anomaly |= LN_A_CAPACITY;
    else { /* window offsets of this thread's line ends, in order.  The first two LFs of every 32-byte span are stored by predicated
            * instructions on a running shared-memory address (no branches); a loop only for a third LF in 32 bytes. */
      uint32_t addr = smem_u32(lend) + 2 * excl;
      const uint32_t e0 = LN_LEFT + 16 * c0 + 1;
#define LN_EMIT1(h_) asm volatile("{ .reg .pred p; .reg .u32 t, b;\n" \
                                  "  setp.ne.u32 p, %1, 0; brev.b32 t, %1; bfind.shiftamt.u32 b, t; add.u32 b, b, %2;\n" \
                                  "  @p st.shared.u16 [%0], b; @p add.u32 %0, %0, 2;\n" \
                                  "  add.u32 t, %1, -1; and.b32 %1, %1, t; }" : "+r"(addr), "+r"(w) : "r"(e0 + 32 * (h_)) : "memory");
#define LN_EMIT(w_, h_) { uint32_t w = (w_); LN_EMIT1(h_) LN_EMIT1(h_) while (w) LN_EMIT1(h_) }
      LN_EMIT(mm.x, 0) LN_EMIT(mm.y, 1) LN_EMIT(mm.z, 2) LN_EMIT(mm.w, 3)
#undef LN_EMIT
#undef LN_EMIT1
    }
    __syncthreads();
    if (pend) { /* F, second half (after this tile's own count went out: nothing that can wait runs before that) */
      uint32_t base = p_base;
      if (f_lookup) {
        const uint32_t state = s_fstate; base = s_fbase;
        if (state != 1u) { /* not among the 256 in front, or one of them was late: the look-back with its own barriers */
          __syncthreads();
          base = ln_lookback(P.tile_state, p_tile, 0ull, tid, lane, warp, s_w3, s_w4, P.out);
        }
      }
      finish_previous(base, f_lookup);
    }

    /* ---- C0: line class of the tile's first byte.  The true value needs the number of lines in front of the tile; that sum is
     * looked up one round later (F), when the tiles in front have long published their counts.  Until then the tile's own lines
     * answer: a line that is exactly "+\n" is a plus line.  All such lines among the first 32 must agree, or we wait for the sum. */
    uint32_t phi = 0, base_line = 0;
    bool have_base = false;
    {
      const uint32_t k = lane + 1; /* line k of the tile starts at lend[k-1] */
      bool plus = false;
      if (!too_many && k < cntW) { const uint32_t s = lend[k - 1], e = lend[k]; plus = e - s == 2u && win[s] == '+'; }
      const uint32_t pm = __ballot_sync(FULL, plus);
      /* line k is a plus line ⇒ class of the tile's first line = (2 - k) & 3; bit (k-1) of pm ↔ line k */
      const uint32_t w0 = pm & 0x11111111u, w1 = pm & 0x22222222u, w2 = pm & 0x44444444u, w3 = pm & 0x88888888u;
      const uint32_t votes = (w0 ? 1u : 0u) + (w1 ? 1u : 0u) + (w2 ? 1u : 0u) + (w3 ? 1u : 0u);
      if (tile == 0) { /* nothing in front of the first tile */
        if (tid == 0) st_volatile64(P.tile_state, ST_INCL | (unsigned long long)cntT);
        have_base = true;
        phi = (4u - P.j0) & 3u;
      } else if (votes == 1u && !(LINES && tile + 2u >= P.ntiles)) phi = w0 ? 1u : w1 ? 0u : w2 ? 3u : 2u; /* k = 1, 5, .. → 1;  k = 2, 6, .. → 0;  k = 3, .. → 3;  k = 4, .. → 2 */
      else { /* no witness (long lines, or plus lines that repeat the name), or witnesses that disagree: wait for the sum now */
        __syncthreads(); /* the look-back scratch may still be read by the previous tile's look-back */
        if (tid == 0) atomicAdd(P.out + LN_O_WAITED, 1u);
        base_line = ln_lookback(P.tile_state, tile, 0ull, tid, lane, warp, s_w3, s_w4, P.out);
        if (tid == 0) st_volatile64(P.tile_state + tile, ST_INCL | ((unsigned long long)base_line + cntT));
        have_base = true;
        phi = (base_line + 4u - P.j0) & 3u;
      }
    }
    if (LINES && (tile == 0 || tile + 2u >= P.ntiles) && !too_many) {
      /* the host wants the first and the last line ends of the chunk (where its first record starts, where the last complete one
       * ends): these tiles know the lines in front of them already and write their line ends out themselves */
      const uint32_t gofs = (uint32_t)(t0 - LN_LEFT);
      const bool all = tile + 2u >= P.ntiles;
      for (uint32_t r = tid; r < cntT; r += LN_THREADS) { const uint32_t gi = base_line + r; if (gi < P.cap && (all || gi < 8u)) P.line_end[gi] = gofs + lend[r]; }
      if (tid == 0 && (tile + 2u == P.ntiles || P.ntiles == 1u)) P.out[LN_O_INDEX_FROM] = base_line;
    }
    /* line numbers relative to the tile: class = number & 3, record of the tile = (number >> 2) - 1 */
    const uint32_t gbr = 4u + phi;
    const uint32_t g0t = gbr + excl;                /* line number at this thread's first byte */

    if (LINES) {
      /* ---- L: one thread per line.  Lines are taken by the tile they start in; line k of the window (0..) starts at lend[k-1]
       * and ends at lend[k] (tile or margin).  Every fourth line has the same class, so the i-th line of a class is found by
       * arithmetic; warps take groups of 32 lines of ONE class: sequence lines first, then quality, then header lines. */
      p_nstage = 0; p_rl0 = 0;
      if (!too_many) {
        const uint32_t kmin = win[LN_LEFT - 1] == '\n' ? 0u : 1u;
        const uint32_t tile_end = LN_LEFT + nv;          /* lines starting at or beyond belong to the next tile (or do not exist) */
        const uint32_t gofs = (uint32_t)(t0 - LN_LEFT);
        uint32_t kH = kmin + ((0u - (gbr + kmin)) & 3u), kS = kmin + ((1u - (gbr + kmin)) & 3u), kQ = kmin + ((3u - (gbr + kmin)) & 3u);
        if (tile == 0) { /* lines of the record cut by the start of the chunk are judged with their record */
          if (kH < P.j0) kH += 4; if (kS < P.j0) kS += 4; if (kQ < P.j0) kQ += 4;
        }
        const uint32_t nS = kS <= cntT ? (cntT - kS) / 4 + 1 : 0, nQ = kQ <= cntT ? (cntT - kQ) / 4 + 1 : 0;
        uint32_t nH = kH <= cntT ? (cntT - kH) / 4 + 1 : 0;
        if (nH > (uint32_t)LN_SMAX) { anomaly |= LN_A_CAPACITY; nH = 0; }
        p_nstage = nH; p_rl0 = (gbr + kH) >> 2;
        route_tile_max = max(route_tile_max, nH);
        /* groups of 32 lines of one class, handed to the warps round-robin in the order quality, sequence, header: the warp that
         * gets a second group gets the light ones (measured: 0.89 ms against 0.94 for sequence-first, 0.96 for a shared work counter
         * with sequence lines split in halves — extra instructions cost more than balance gains) */
        const uint32_t GQ = (nQ + 31) >> 5, GS = (nS + 31) >> 5, GH = (nH + 31) >> 5;
        const bool snake = (P.tune & 2u) == 0u;
        for (uint32_t rr = 0;; rr++) {
          /* groups in the order of their cost — sequence lines (the alphabet), header lines (walk, hash, name copy, record rules), quality
           * lines — dealt to the warps back and forth, so that the warp with a heavy first group gets a light second one */
          uint32_t g;
          if (snake) g = (rr & 1u) ? LN_WARPS * rr + (LN_WARPS - 1u - warp) : LN_WARPS * rr + warp;
          else g = LN_WARPS * rr + warp;
          if (LN_WARPS * rr >= GQ + GS + GH) break;
          if (g >= GQ + GS + GH) continue;
          uint32_t cls, gi;
          if (snake) { cls = g < GS ? 1u : g < GS + GH ? 0u : 3u; gi = cls == 1u ? g : cls == 0u ? g - GS : g - GS - GH; }
          else { cls = g < GQ ? 3u : g < GQ + GS ? 1u : 0u; gi = cls == 3u ? g : cls == 1u ? g - GQ : g - GQ - GS; } /* uniform over the warp */
          const uint32_t i = 32 * gi + lane;
          const uint32_t n = cls == 1u ? nS : cls == 3u ? nQ : nH;
          const uint32_t k = (cls == 1u ? kS : cls == 3u ? kQ : kH) + 4 * i;
          if (cls != 0u) {
            if (i >= n) continue;
            const uint32_t s = k == 0 ? (uint32_t)LN_LEFT + lead0 : (uint32_t)lend[k - 1];
            if (s >= tile_end) continue;
            uint32_t e; /* one past the line's last byte; has_lf: that byte is its LF */
            bool has_lf = true;
            if (k < cntW) e = lend[k];
            else if (nloc < (uint32_t)(LN_LEFT + TILE + LN_MARGIN)) { /* the data ends inside the window */
              if (!P.virtual_end) continue;                             /* more follows: the rest of the line comes with the next chunk */
              e = nloc; has_lf = false;                                 /* last line of the file, without LF */
            } else { anomaly |= LN_A_CAPACITY; continue; }              /* longer than the margin: not for this mode */
            const uint32_t ce = has_lf ? e - 1u : e;                    /* content: [s, ce) */
            if (cls == 1u) { if (ce > s) seq_bad |= ls_seq_line(win, lut, s, ce); }
            else {
              if (!(win[s - 1] == '\n' && win[s - 2] == '+' && win[s - 3] == '\n')) anomaly |= LN_A_PLUS; /* the line in front must be "+\n" */
              if (ce > s) ls_qual_line(win, lut, s, ce, qmn, qmx);
            }
            continue;
          }
          /* ---- header lines: one RECORD per lane.  The '@' syntax, the name slice and its hash in one walk; the name's bytes go
           * to the arena; the record's length rules (src/fastq.c:346, :380; src/fastq.h:30-37) and its statistics (src/fastq.c:97-110)
           * from the ends of its four lines.  No lane leaves early: the warp reserves arena space and reduces the statistics together. */
          const bool window_full = nloc == (uint32_t)(LN_LEFT + TILE + LN_MARGIN);
          bool live = i < n;
          uint32_t s = 0, e = 0;
          if (live) { stage[i].len = 0xFFFFFFFFu; s = k == 0 ? (uint32_t)LN_LEFT + lead0 : (uint32_t)lend[k - 1]; live = s < tile_end; } /* (nothing to write out unless the header is judged below) */
          if (live) {
            if (k < cntW) e = lend[k];
            else { live = false; if (window_full) anomaly |= LN_A_CAPACITY; else if (P.virtual_end) anomaly |= LN_A_HEADER; } /* no LF in the window: longer than the margin / the file's last line / the rest comes with the next chunk */
          }
          const uint32_t hl = e - s;
          if (live && hl >= FQ_MAX_LABEL_LENGTH) { atomicMin(P.out + LN_O_OVERLONG, tile); live = false; }
          if (live && hl < 3u) { anomaly |= LN_A_HEADER; live = false; }
          uint32_t nlen = 0; uint64_t mem_len = 0, hsh = FQ_HASH_SKIP;
          if (live && !fq_header_hash_fast(win, s, hl, P.cx.fmt_key, P.cx.pe_key, P.cx.seed, P.names != nullptr || P.route_world != 0, &nlen, &mem_len, &hsh)) { anomaly |= LN_A_HEADER; live = false; }
          /* the record's other three lines: is it complete in this chunk, does it keep the length rules? */
          bool rec_ok = false; uint32_t sl = 0;
          if (live) {
            uint32_t e1 = 0, e2 = 0, e3 = 0; bool complete = false;
            if (k + 3u < cntW) { e1 = lend[k + 1]; e2 = lend[k + 2]; e3 = lend[k + 3]; complete = true; }
            else if (window_full) anomaly |= LN_A_CAPACITY; /* a record longer than the margin */
            else if (P.virtual_end && k + 3u == cntW && win[nloc - 1] != '\n') { e1 = lend[k + 1]; e2 = lend[k + 2]; e3 = nloc + 1u; complete = true; } /* the file's last line has no LF: compare contents */
            /* otherwise the end of the chunk cut the record: it is judged with the next chunk, or by the host at the end of the file */
            if (complete) {
              sl = e1 - e;
              if (e2 - e1 != 2u || sl < 2u || e3 - e2 != sl) atomicOr(P.out + LN_O_RECBAD, 1u);
              else rec_ok = true;
            }
          }
          /* arena space for the names of the group, inside the tile's own stretch */
          const uint32_t units = (live && P.arena) ? (nlen + 15u) >> 4 : 0u;
          uint32_t incl_u = units;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const uint32_t a = __shfl_up_sync(FULL, incl_u, d); if (lane >= d) incl_u += a; }
          uint32_t ubase = 0;
          { const uint32_t tot = __shfl_sync(FULL, incl_u, 31); if (lane == 0 && tot) ubase = atomicAdd(&s_arena, tot); }
          ubase = __shfl_sync(FULL, ubase, 0) + incl_u - units;
          if (units && ubase + units > arena_stride) { anomaly |= LN_A_CAPACITY; live = false; }
          const uint32_t my_unit = tile * arena_stride + ubase;
          uint4* dst = (uint4*)(P.arena + (size_t)my_unit * 16u);
          uint32_t copy_units = units;
          if (P.route_world) { /* the owner of the name's hash gets it at once: the next free slot of this CTA's stretch of its region */
            const bool go = live && rec_ok; /* (a record the end of the chunk cut travels with the chunk that holds all of it) */
            const uint32_t o = go ? fq_owner_of(hsh, P.route_world) : 0xFFFFFFFFu;
            const uint32_t m = __match_any_sync(FULL, o);
            const int leader = __ffs(m) - 1;
            uint32_t pos = 0;
            if (go && lane == leader) pos = atomicAdd(&s_own[o], (uint32_t)__popc(m));
            pos = __shfl_sync(FULL, pos, leader) + __popc(m & ((1u << lane) - 1u));
            if (go && pos >= P.route_stride) { anomaly |= LN_A_CAPACITY; live = false; }
            copy_units = 0;
            if (go && live) {
              uint4* slot = (uint4*)(s_reg[o] + (size_t)pos * route_slot_bytes);
              const unsigned long long rl = ((unsigned long long)tile << 24 | (unsigned long long)(i & 0xFFFu) << 12) | nlen; /* (a record number of its own kind: unique, nothing more is asked of it) */
              slot[0] = make_uint4((uint32_t)hsh, (uint32_t)(hsh >> 32), (uint32_t)rl, (uint32_t)(rl >> 32));
              dst = slot + 1; copy_units = (nlen + 15u) >> 4;
              if (copy_units > P.route_units) { copy_units = P.route_units; if (P.route_units) atomicOr(&((FqRegionHdr*)P.route_region[o])->flags, FQ_ROUTE_NAME_TOO_LONG); } /* (no units: the tuple travels alone by design) */
              for (uint32_t u = copy_units; u < P.route_units; u++) dst[u] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          if (live) {
            FqName nm; nm.off = gofs + s + 1; nm.len = nlen; nm.hash = hsh;
            if (copy_units) { /* the name's bytes, 16 at a time from the window (any alignment), zero padded */
              if (units) nm.off = my_unit * 16u;
              const uint32_t a0 = (s + 1u) & ~3u, sh = ((s + 1u) & 3u) * 8u;
              for (uint32_t u = 0; u < copy_units; u++) {
                const uint32_t* wp = (const uint32_t*)(win + a0 + 16u * u);
                const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
                const uint32_t left = nlen - 16u * u;
                uint4 v;
                v.x = fq_low_bytes(__funnelshift_r(w0, w1, sh), left);
                v.y = left > 4u ? fq_low_bytes(__funnelshift_r(w1, w2, sh), left - 4u) : 0u;
                v.z = left > 8u ? fq_low_bytes(__funnelshift_r(w2, w3, sh), left - 8u) : 0u;
                v.w = left > 12u ? fq_low_bytes(__funnelshift_r(w3, w4, sh), left - 12u) : 0u;
                dst[u] = v;
              }
            }
            stage[i] = nm;
          }
          const uint32_t okm = __ballot_sync(FULL, rec_ok);
          if (okm) {
            const uint32_t memw = __reduce_add_sync(FULL, rec_ok ? (uint32_t)mem_len : 0u);
            const uint32_t mn = __reduce_min_sync(FULL, rec_ok ? sl : 0xFFFFFFFFu), mx = __reduce_max_sync(FULL, rec_ok ? sl : 0u);
            if (lane == 0) { atomicAdd(&s_cnt, (uint32_t)__popc(okm)); atomicAdd(&s_mem, memw); atomicMin(&s_rlmin, mn); atomicMax(&s_rlmax, mx); }
            if (mn == mx && mx < (uint32_t)LS_HBINS) { if (lane == 0) atomicAdd(&s_hist[mx], (uint32_t)__popc(okm)); } /* one read length: the common case */
            else if (rec_ok) { if (sl < (uint32_t)LS_HBINS) atomicAdd(&s_hist[sl], 1u); else atomicAdd(&P.stage->hist[sl < (uint32_t)LS_STAGE_HIST ? sl : (uint32_t)LS_STAGE_HIST - 1u], 1u); }
          }
        }
      }
    } else {
      /* ---- C1: whole chunks of sequence / quality lines: classify (all lanes in step), then one list per warp */
      uint32_t n_seq_w, n_qual_w;
      {
        uint32_t is_seq = 0, is_qual = 0; /* bit i: chunk i of this thread is a whole chunk of that class */
#pragma unroll
        for (int i = 0; i < LN_CPT; i++) {
          const uint32_t m = ((i < 2 ? mm.x : i < 4 ? mm.y : i < 6 ? mm.z : mm.w) >> (16 * (i & 1))) & 0xFFFFu;
          const uint32_t cum = ((i < 4 ? cumA : cumB) >> (8 * (i & 3))) & 0xFFu;
          const uint32_t cls = (g0t + cum) & 3u;
          bool whole = m == 0;
          if (!full) whole = whole && nv >= 16 * (c0 + i) + 16;
          if (i == 0) whole = whole && !(tid == 0 && lead0);
          is_seq |= (whole && cls == 1u) ? 1u << i : 0u;
          is_qual |= (whole && cls == 3u) ? 1u << i : 0u;
        }
        const uint32_t n_pure = __popc(is_seq) | (__popc(is_qual) << 16);
        uint32_t ip = n_pure;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, ip, d); if (lane >= d) ip += a; }
        const uint32_t tw = __shfl_sync(FULL, ip, 31), xp = ip - n_pure;
        n_seq_w = tw & 0xFFFFu; n_qual_w = tw >> 16;
        uint32_t as = smem_u32(pure) + 2 * (warp * 256 + (xp & 0xFFFFu)), aq = smem_u32(pure) + 2 * (warp * 256 + 255 - (xp >> 16));
#pragma unroll
        for (int i = 0; i < LN_CPT; i++) {
          asm volatile("{ .reg .pred p, q; setp.ne.u32 p, %2, 0; setp.ne.u32 q, %3, 0;\n"
                       "  @p st.shared.u16 [%0], %4; @p add.u32 %0, %0, 2;\n"
                       "  @q st.shared.u16 [%1], %4; @q sub.u32 %1, %1, 2; }"
                       : "+r"(as), "+r"(aq) : "r"(is_seq & (1u << i)), "r"(is_qual & (1u << i)), "h"((uint16_t)(c0 + i)) : "memory");
        }
      }
      __syncwarp();

      if (!too_many) {
        /* ---- D: sequence alphabet, quality range; one predicate per warp instruction.
         * Whole chunks come from the warp's lists.  Partial chunks hang on LFs: the bytes after a header / plus LF and before a
         * sequence / quality LF; the LFs of one kind are every fourth line end, so item i of a kind is found by arithmetic. */
        const uint16_t* pw = pure + warp * 256;
        for (uint32_t i = lane; i < n_seq_w; i += 32)
          seq_ok &= ln_pred4(*(const uint4*)(win + LN_LEFT + 16 * (uint32_t)pw[i]));
        for (uint32_t i = lane; i < n_qual_w; i += 32) {
          const uint4 v = *(const uint4*)(win + LN_LEFT + 16 * (uint32_t)pw[255 - i]);
          ln_minmax_word(v.x, qmn, qmx); ln_minmax_word(v.y, qmn, qmx); ln_minmax_word(v.z, qmn, qmx); ln_minmax_word(v.w, qmn, qmx);
        }
        const uint32_t n_items = 2u * ((cntT + 3u) >> 2); /* per kind: two per started group of four lines */
#pragma unroll
        for (int kind = 0; kind < 2; kind++) { /* 0: sequence side (classes 0, 1), 1: quality side (classes 2, 3) */
          const uint32_t ra = ((kind ? 2u : 0u) - gbr) & 3u, rb = ((kind ? 3u : 1u) - gbr) & 3u; /* first line ends of those classes */
          for (uint32_t i = tid; i < n_items; i += LN_THREADS) {
            const bool before = i & 1u;                     /* odd items: the bytes before a sequence / quality LF */
            const uint32_t r = 4u * (i >> 1) + (before ? rb : ra);
            if (r < cntT) {
              const uint32_t e = lend[r], q = e - (LN_LEFT + 1), c = q >> 4, p = q & 15u;
              const uint32_t mc = maskbuf[c];               /* all LFs of this chunk */
              const uint32_t below = mc & ((1u << p) - 1u), above = mc >> (p + 1);
              uint32_t nvc = 16u;
              if (!full) nvc = min(16u, nv - 16 * c);       /* the chunk holds an LF, so it starts inside the data */
              const uint32_t lo = before ? (below ? 32u - __clz(below) : 0u) : p + 1;
              const uint32_t hi = before ? p : (above ? p + __ffs(above) : nvc);
              if (hi > lo) {
                const uint4 v = *(const uint4*)(win + LN_LEFT + 16 * c);
                const uint4 a = lut[lo], b = lut[15u + hi];
                if (kind == 0) /* bytes outside [lo, hi) always pass */
                  seq_ok &= (fq_base_pred(v.x) | ~(a.x & b.x)) & (fq_base_pred(v.y) | ~(a.y & b.y)) & (fq_base_pred(v.z) | ~(a.z & b.z)) & (fq_base_pred(v.w) | ~(a.w & b.w));
                else {
                  const uint32_t fill = (uint32_t)win[LN_LEFT + 16 * c + lo] * 0x01010101u; /* a byte of the range stands in for the bytes outside it */
                  uint32_t m;
                  m = a.x & b.x; ln_minmax_word((v.x & m) | (fill & ~m), qmn, qmx);
                  m = a.y & b.y; ln_minmax_word((v.y & m) | (fill & ~m), qmn, qmx);
                  m = a.z & b.z; ln_minmax_word((v.z & m) | (fill & ~m), qmn, qmx);
                  m = a.w & b.w; ln_minmax_word((v.w & m) | (fill & ~m), qmn, qmx);
                }
              }
            }
          }
        }
        /* the chunk cut by the end of the data, when no LF of its own bounds it */
        if (!full && (nv & 15u) && (nv >> 4) >= c0 && (nv >> 4) < c0 + LN_CPT && maskbuf[nv >> 4] == 0) {
          const uint32_t i = (nv >> 4) - c0, c = nv >> 4, hi = nv & 15u;
          const uint32_t cum = ((i < 4 ? cumA : cumB) >> (8 * (i & 3))) & 0xFFu, cls = (g0t + cum) & 3u;
          for (uint32_t k = 0; k < hi; k++) {
            const uint32_t ch = win[LN_LEFT + 16 * c + k];
            if (cls == 1u) { if (!((fq_base_pred(ch) >> 7) & 1u)) seq_ok = 0; }
            else if (cls == 3u) { ln_minmax_word(ch * 0x01010101u, qmn, qmx); }
          }
        }

        /* ---- E: header and plus lines that start in this tile.  Line k of the tile (0..cntT) starts at lend[k-1].  Names are staged
         * in shared memory under their number inside the tile and written out next round (F). */
        {
          const uint32_t kmin = win[LN_LEFT - 1] == '\n' ? 0u : 1u;
          const uint32_t tile_end = LN_LEFT + nv; /* lines starting at or beyond belong to the next tile (or do not exist) */
          const uint32_t gofs = (uint32_t)(t0 - LN_LEFT);
          const uint32_t kh0 = kmin + ((0u - (gbr + kmin)) & 3u), kp0 = kmin + ((2u - (gbr + kmin)) & 3u);
          uint32_t nH = kh0 <= cntT ? (cntT - kh0) / 4 + 1 : 0;
          const uint32_t nP = kp0 <= cntT ? (cntT - kp0) / 4 + 1 : 0;
          if (nH > (uint32_t)LN_SMAX) { anomaly |= LN_A_CAPACITY; nH = 0; }
          p_nstage = nH; p_rl0 = (gbr + kh0) >> 2;
          for (uint32_t u = tid; u < nH + nP; u += LN_THREADS) {
            const bool is_hdr = u < nH;
            const uint32_t k = is_hdr ? kh0 + 4 * u : kp0 + 4 * (u - nH);
            const uint32_t s = k == 0 ? (uint32_t)LN_LEFT + lead0 : (uint32_t)lend[k - 1];
            if (is_hdr) stage[u].len = 0xFFFFFFFFu; /* nothing to write out unless the header is judged below */
            if (s >= tile_end) continue;
            if (tile == 0 && k < P.j0) continue; /* lines of the record cut by the start of the chunk: judged with their record */
            if (!is_hdr) { /* "+\n" */
              if (s + 1 >= nloc) continue; /* cut by the end of the data: the record is completed (or judged) elsewhere */
              if (!(win[s] == '+' && win[s + 1] == '\n')) anomaly |= LN_A_PLUS;
              continue;
            }
            uint32_t e = 0;
            if (k < cntT) e = lend[k];
            else { /* the tile's last line: its LF lies in the margin */
              uint32_t p = LN_LEFT + nv;
              for (; p < nloc; p++) if (win[p] == '\n') { e = p + 1; break; }
              if (!e) {
                if (nloc == (uint32_t)LN_WIN) atomicMin(P.out + LN_O_OVERLONG, tile); /* no LF within 1 KiB: a line gzgets would split */
                continue; /* otherwise cut by the end of the data */
              }
            }
            const uint32_t hl = e - s;
            if (hl >= FQ_MAX_LABEL_LENGTH) { atomicMin(P.out + LN_O_OVERLONG, tile); continue; }
            uint32_t nlen; uint64_t mem_len, hsh = FQ_HASH_SKIP;
            if (!fq_header_hash_fast(win, s, hl, P.cx.fmt_key, P.cx.pe_key, P.cx.seed, P.names != nullptr, &nlen, &mem_len, &hsh)) { anomaly |= LN_A_HEADER; continue; }
            FqName nm; nm.off = gofs + s + 1; nm.len = nlen; nm.hash = hsh;
            stage[u] = nm;
          }
        }
      } else { p_nstage = 0; p_rl0 = 0; }
    }
    if (!LINES && tid == 0) s_next = (lost_a != 0 || lost_o != 0xFFFFFFFFu) ? 0xFFFFFFFFu : atomicAdd(P.ticket, 1u); /* everyone read s_next before the barriers of this round */
    if (seq_bad || (seq_ok & 0x80808080u) != 0x80808080u) anomaly |= LN_A_BASE;
    if (anomaly & ~told) { atomicOr(P.out + LN_O_ANOMALY, anomaly); told |= anomaly; } /* known at once: the other CTAs stop early */
    pend = true; p_tile = tile; p_cnt = cntT; p_phi = phi; p_have_base = have_base; p_base = base_line;
    p_no_final_lf = P.virtual_end && tile == P.ntiles - 1 && P.n > 0 && win[nloc - 1] != '\n';
    buf ^= 1u;
  }

  if (LINES) { /* this CTA's records → the chunk's stage */
    __syncthreads();
    for (int l = tid; l < LS_HBINS; l += LN_THREADS) { const uint32_t v = s_hist[l]; if (v) atomicAdd(&P.stage->hist[l], v); }
    if (tid == 0) {
      if (s_cnt) {
        atomicAdd(&P.stage->nrec, (unsigned long long)s_cnt); atomicAdd(&P.stage->mem, (unsigned long long)s_mem);
        atomicMax(&P.stage->rl_min_inv, ~s_rlmin); atomicMax(&P.stage->rl_max, s_rlmax);
      }
      atomicMax(&P.stage->arena_used, max(arena_max, s_arena)); /* the fullest tile's */
    }
    if (P.route_world && tid < (int)P.route_world) { /* this CTA's stretch of every region: how many slots it filled */
      ((uint32_t*)(P.route_region[tid] + 16))[blockIdx.x] = s_own[tid];
      if (blockIdx.x == 0) { FqRegionHdr* h = (FqRegionHdr*)P.route_region[tid]; h->nblocks = gridDim.x; h->stride = P.route_stride; }
    }
  }
  /* ---- results of this thread → one set of atomics per warp */
  if ((seq_ok & 0x80808080u) != 0x80808080u || seq_bad) anomaly |= LN_A_BASE;
  uint32_t mn = min(qmn & 0xFFFFu, qmn >> 16), mx = max(qmx & 0xFFFFu, qmx >> 16);
  mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
  if (mn <= 0x0Du) anomaly |= LN_A_QUAL; /* NUL / LF / CR (or another control byte) inside a quality line: let the careful path look */
  anomaly = __reduce_or_sync(FULL, anomaly);
  if (lane == 0) {
    if (anomaly) atomicOr(P.out + LN_O_ANOMALY, anomaly);
    if (mn <= mx) { atomicMin(P.out + LN_O_QMIN, mn); atomicMax(P.out + LN_O_QMAX, mx); }
  }
}

/* ------------------------------------------------------------------------------------------------ K5p: the verdict of a per-line pass
 * One block after the pass: is the chunk clean?  Then (and only then) what the pass staged becomes part of the file statistics —
 * fastq_new_entry_stats (src/fastq.c:97-110) `weight` times per record, the index bookkeeping (n_entries, index_mem), the quality
 * range — and the chunk's last line ends travel to the host with the result words. */
struct LanesPostParams {
  uint32_t* out; const uint32_t* line_end; const LanesStage* stage; uint32_t j0; FqRecCtx cx;
  FqStats* stats; FqStats* stats_range; unsigned long long* hist; uint32_t names_cap; /* 0: the loop has no names */
};
__global__ void __launch_bounds__(1024)
fq_lanes_post_kernel(const LanesPostParams P) {
  const uint32_t nlines = P.out[LN_O_LINES];
  const uint32_t nrec = nlines > P.j0 ? (nlines - P.j0) / 4 : 0;
  const bool consistent = (P.stage->nrec == nrec && (!P.names_cap || nrec <= P.names_cap)) || P.out[LN_O_RECBAD] != 0; /* (a name descriptor beyond the array's capacity was dropped) */
  const bool accept = !P.out[LN_O_CAPOVF] && P.out[LN_O_OVERLONG] == 0xFFFFFFFFu && !P.out[LN_O_ANOMALY] && !P.out[LN_O_INTERNAL] && !P.out[LN_O_RECBAD] && consistent;
  __syncthreads();
  if (threadIdx.x == 0) {
    P.out[LN_O_ACCEPT] = accept ? 1u : 0u; P.out[LN_O_ARENA_USED] = P.stage->arena_used; P.out[LN_O_STAGED] = (uint32_t)P.stage->nrec; /* (units per tile: the host scales) */
    if (!consistent) P.out[LN_O_INTERNAL] = 3u; /* every complete record must have been judged exactly once */
  }
  if (!accept) return;
  if (threadIdx.x < 8) { /* the last line ends travel to the host with the result words */
    const uint32_t from = nlines > 8 ? nlines - 8 : 0;
    if (from + threadIdx.x < nlines) P.out[16 + threadIdx.x] = P.line_end[from + threadIdx.x];
  }
  const unsigned long long w = P.cx.weight;
  const uint32_t lmax = min(P.stage->rl_max, (uint32_t)LS_STAGE_HIST - 1u); /* (no bin above the longest read is set) */
  for (uint32_t l = threadIdx.x; l <= lmax; l += blockDim.x) { const uint32_t v = P.stage->hist[l]; if (v) atomicAdd(P.hist + l, w * v); }
  if (threadIdx.x == 0) {
    const unsigned long long n = P.stage->nrec;
    if (n) {
      atomicAdd(&P.stats->num_rds, n * w);
      if (P.cx.loop == FQ_LOOP_INDEX) { atomicAdd(&P.stats->n_names, n); atomicAdd(&P.stats->mem_sum, P.stage->mem); }
      atomicMin(&P.stats_range->min_rl, ~P.stage->rl_min_inv); atomicMax(&P.stats_range->max_rl, P.stage->rl_max);
    }
    if (P.out[LN_O_QMIN] <= P.out[LN_O_QMAX]) { atomicMin(&P.stats_range->min_q, P.out[LN_O_QMIN]); atomicMax(&P.stats_range->max_q, P.out[LN_O_QMAX]); }
    P.out[LN_O_RLMIN] = ~P.stage->rl_min_inv; P.out[LN_O_RLMAX] = P.stage->rl_max;
  }
}

/* ------------------------------------------------------------------------------------------------ K5r: record rules + statistics
 * One thread per record over the line index: over-long lines (src/fastq.h:30-37), read length >= 1 (src/fastq.c:346), equal
 * sequence / quality lengths (:380), fastq_new_entry_stats (:97-110) and the index bookkeeping (n_entries, index_mem).
 * sign = -1 takes the same counts back (a chunk that failed one of these rules is not committed). */
struct LanesRecParams {
  const uint32_t* line_end; uint32_t* out; uint32_t j0; uint32_t lead; const FqName* names; FqRecCtx cx;
  FqStats* stats; unsigned long long* hist; int undo;
};
__device__ __forceinline__ void hist_flush_signed(unsigned long long* hist, uint32_t len, uint32_t count, int undo) {
  unsigned active = __ballot_sync(FULL, count > 0);
  if (!active) return;
  if (count > 0) {
    unsigned peers = __match_any_sync(active, len);
    unsigned long long sum = 0;
    for (unsigned p = peers; p; p &= p - 1) sum += __shfl_sync(peers, count, __ffs(p) - 1);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(hist + len, undo ? 0ull - sum : sum);
  }
}
__global__ void __launch_bounds__(256)
fq_lanes_records_kernel(const LanesRecParams P) {
  /* a chunk the pass itself rejected is never looked at */
  if (P.out[LN_O_CAPOVF] || P.out[LN_O_OVERLONG] != 0xFFFFFFFFu || P.out[LN_O_ANOMALY] || P.out[LN_O_INTERNAL]) return;
  const uint32_t nlines = P.out[LN_O_LINES], virt = P.out[LN_O_VIRTUAL];
  const uint32_t nrec = nlines > P.j0 ? (nlines - P.j0) / 4 : 0;
  if (!P.undo && blockIdx.x == 0 && threadIdx.x < 8) { /* the last line ends travel to the host with the result words */
    const uint32_t from = nlines > 8 ? nlines - 8 : 0;
    if (from + threadIdx.x < nlines) P.out[16 + threadIdx.x] = P.line_end[from + threadIdx.x];
  }
  const int lane = threadIdx.x & 31;
  unsigned long long my_rds = 0, my_names = 0, my_mem = 0;
  uint32_t mn_rl = 0xFFFFFFFFu, mx_rl = 0, run_len = 0, run_cnt = 0, bad = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < nrec; base += stride) {
    const uint32_t r = base + threadIdx.x;
    uint32_t flush_len = 0, flush_cnt = 0;
    if (r < nrec) {
      const uint32_t j = P.j0 + 4 * r;
      const uint32_t s0 = j ? P.line_end[j - 1] : P.lead, e0 = P.line_end[j], e1 = P.line_end[j + 1], e2 = P.line_end[j + 2], e3 = P.line_end[j + 3];
      const uint32_t hl = e0 - s0, sl = e1 - e0, pl = e2 - e1;
      uint32_t ql = e3 - e2;
      if (j + 3 == virt) ql += 1; /* last line of the file without LF: compare contents */
      if (hl >= FQ_MAX_LABEL_LENGTH || pl != 2u || sl >= FQ_MAX_READ_LENGTH || ql >= FQ_MAX_READ_LENGTH || sl < 2u || ql != sl) bad = 1;
      else {
        my_rds += P.cx.weight;
        mn_rl = min(mn_rl, sl); mx_rl = max(mx_rl, sl);
        if (P.cx.loop == FQ_LOOP_INDEX) { my_names++; my_mem += P.names[r].len + (P.cx.fmt_key == FQ_FMT_CASAVA ? 0u : 1u); }
        if (run_cnt && sl != run_len) { flush_len = run_len; flush_cnt = run_cnt; run_cnt = 0; }
        run_len = sl; run_cnt += P.cx.weight;
      }
    }
    hist_flush_signed(P.hist, flush_len, flush_cnt, P.undo);
  }
  hist_flush_signed(P.hist, run_len, run_cnt, P.undo);
  my_rds = warp_sum64(my_rds); my_names = warp_sum64(my_names); my_mem = warp_sum64(my_mem);
  mn_rl = __reduce_min_sync(FULL, mn_rl); mx_rl = __reduce_max_sync(FULL, mx_rl);
  bad = __reduce_or_sync(FULL, bad);
  if (lane == 0) {
    if (my_rds) atomicAdd(&P.stats->num_rds, P.undo ? 0ull - my_rds : my_rds);
    if (my_names) { atomicAdd(&P.stats->n_names, P.undo ? 0ull - my_names : my_names); atomicAdd(&P.stats->mem_sum, P.undo ? 0ull - my_mem : my_mem); }
    if (!P.undo) {
      if (mx_rl) { atomicMin(P.out + LN_O_RLMIN, mn_rl); atomicMax(P.out + LN_O_RLMAX, mx_rl); }
      if (bad) atomicOr(P.out + LN_O_RECBAD, 1u);
    }
  }
}

__global__ void fq_lanes_commit_kernel(const uint32_t* out, FqStats* stats_range) {
  if (threadIdx.x || blockIdx.x) return;
  if (out[LN_O_RLMAX]) { atomicMin(&stats_range->min_rl, out[LN_O_RLMIN]); atomicMax(&stats_range->max_rl, out[LN_O_RLMAX]); }
  if (out[LN_O_QMIN] <= out[LN_O_QMAX]) { atomicMin(&stats_range->min_q, out[LN_O_QMIN]); atomicMax(&stats_range->max_q, out[LN_O_QMAX]); }
}

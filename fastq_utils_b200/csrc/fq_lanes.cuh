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
constexpr int LN_TILE = 32768, LN_CHUNKS = LN_TILE / 16, LN_CPT = LN_CHUNKS / LN_THREADS; /* 8 chunks per thread */
constexpr int LN_LEFT = 16, LN_MARGIN = 1024, LN_WIN = LN_LEFT + LN_TILE + LN_MARGIN;
constexpr int LN_LMAX = 2048;                 /* line ends kept per tile */
constexpr int LN_OFF_MASK = LN_WIN;
constexpr int LN_OFF_LEND = LN_OFF_MASK + LN_CHUNKS * 2;
constexpr int LN_OFF_PURE = LN_OFF_LEND + LN_LMAX * 2;
constexpr int LN_OFF_LUT = LN_OFF_PURE + LN_CHUNKS * 2;
constexpr int LN_SMEM = LN_OFF_LUT + 32 * 16;
static_assert(LN_CPT == 8, "a thread's masks are one 16-byte load");
static_assert(LN_WIN % 16 == 0 && LN_OFF_LEND % 16 == 0 && LN_OFF_PURE % 16 == 0 && LN_OFF_LUT % 16 == 0, "alignment");

/* anomaly bits (out[3]) */
enum { LN_A_BASE = 1, LN_A_QUAL = 2, LN_A_HEADER = 4, LN_A_PLUS = 8, LN_A_CAPACITY = 16, LN_A_PHASE = 32 };
/* out words */
enum { LN_O_LINES = 0, LN_O_CAPOVF = 1, LN_O_OVERLONG = 2, LN_O_ANOMALY = 3, LN_O_INTERNAL = 4, LN_O_VIRTUAL = 5, LN_O_QMIN = 6, LN_O_QMAX = 7,
       LN_O_RLMIN = 8, LN_O_RLMAX = 9, LN_O_RECBAD = 10, LN_O_WORDS = 12 };

struct LanesParams {
  const uint8_t* data; /* 16-byte aligned (bulk copies) */
  uint32_t n; int virtual_end; uint32_t* line_end; uint32_t cap;
  unsigned long long* tile_state; uint32_t* ticket; uint32_t ntiles;
  uint32_t* out;
  uint32_t j0; FqRecCtx cx; FqName* names; uint32_t names_cap;
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

/* One round of the block-wide look-back: every thread holds the state of one predecessor tile (tile-1-tid, ...); returns true
 * when an inclusive count was among them.  *sum accumulates the counts between that tile and ours.  Contains a barrier. */
__device__ __forceinline__ bool ln_lookback_round(unsigned long long v64, int lane, int warp, uint32_t* s_sum, uint32_t* s_has, uint32_t* sum) {
  const uint32_t incl_mask = __ballot_sync(FULL, (v64 >> 62) == 2);
  const int first = incl_mask ? __ffs(incl_mask) - 1 : 31;
  uint32_t part = lane <= first ? (uint32_t)(v64 & ST_VALUE) : 0u;
  part = __reduce_add_sync(FULL, part);
  if (lane == 0) { s_sum[warp] = part; s_has[warp] = incl_mask ? 1u : 0u; }
  __syncthreads();
  bool found = false;
#pragma unroll
  for (int w = 0; w < LN_WARPS; w++) if (!found) { *sum += s_sum[w]; found = s_has[w] != 0; }
  return found;
}
__device__ __forceinline__ unsigned long long ln_wait_state(const unsigned long long* p, uint32_t* out) {
  unsigned long long v64;
  uint32_t spins = 0;
  while (((v64 = ld_volatile64(p)) >> 62) == 0)
    if (++spins > (1u << 24)) { atomicExch(out + LN_O_INTERNAL, 2u); return ST_INCL; }
  return v64;
}

__global__ void __launch_bounds__(LN_THREADS, 4)
fq_lanes_kernel(const LanesParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* win = smem;
  uint16_t* maskbuf = (uint16_t*)(smem + LN_OFF_MASK);
  uint16_t* lend = (uint16_t*)(smem + LN_OFF_LEND);
  uint16_t* pure = (uint16_t*)(smem + LN_OFF_PURE);   /* whole chunks, one region of 256 entries per warp: sequence from its front, quality from its back */
  uint4* lut = (uint4*)(smem + LN_OFF_LUT);           /* [lo] bytes >= lo, [16 + h] bytes <= h */
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_next, s_guess, s_w1[LN_WARPS], s_w3[LN_WARPS], s_w4[LN_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_next = atomicAdd(P.ticket, 1u);
    s_guess = 0;
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
  uint32_t parity = 0;
  uint32_t seq_ok = 0x80808080u;           /* AND of the alphabet predicate over everything this thread checked */
  uint32_t qmn = 0x00FF00FFu, qmx = 0u;    /* quality minimum / maximum, two 16-bit lanes */
  uint32_t anomaly = 0;

  for (;;) {
    __syncthreads(); /* everyone is done with the previous window and lists; s_next holds the tile claimed for this round */
    const uint32_t tile = s_next;
    if (tile >= P.ntiles) break;
    const unsigned long long t0 = (unsigned long long)tile * LN_TILE;
    if (tid == 0) {
      unsigned long long src = tile ? t0 - LN_LEFT : 0;
      uint32_t dst_off = tile ? 0 : LN_LEFT;
      unsigned long long want = (unsigned long long)LN_WIN - dst_off, have = ((unsigned long long)P.n - src + 15) & ~15ull;
      uint32_t bytes = (uint32_t)(want < have ? want : have);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(win + dst_off)), "l"(P.data + src), "r"(bytes), "r"(bar) : "memory");
      s_guess = 0;
    }
    const uint32_t left = (uint32_t)min((unsigned long long)(LN_TILE + LN_MARGIN), (unsigned long long)P.n - t0); /* data bytes from the tile start */
    const uint32_t nv = min(left, (uint32_t)LN_TILE);   /* valid bytes of the tile itself */
    const uint32_t nloc = LN_LEFT + left;               /* window offsets below this hold data */
    const bool full = nv == (uint32_t)LN_TILE;
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) { if (tid == 0) atomicExch(P.out + LN_O_INTERNAL, 1u); break; } }
      parity ^= 1;
    }
    if (tile == 0 && tid == 0) win[LN_LEFT - 1] = '\n';

    /* ---- A: LF flags; lane ↔ adjacent chunks (conflict-free 128-bit shared loads) */
    {
      const uint32_t cbase = warp * (LN_CHUNKS / LN_WARPS) + lane;
#pragma unroll
      for (int i = 0; i < LN_CPT; i++) {
        const uint32_t c = cbase + i * 32;
        uint32_t m = ln_lf_mask16(*(const uint4*)(win + LN_LEFT + 16 * c));
        if (!full) { uint32_t valid = nv > 16 * c ? min(16u, nv - 16 * c) : 0u; m &= (1u << valid) - 1u; }
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

    /* ---- B: prefix of the LF counts inside the tile; this tile's count is published for the tiles behind us */
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) s_w1[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - tot, cntT = 0;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) { uint32_t x = s_w1[w]; cntT += x; if (w < warp) excl += x; }
    if (tid == 0 && tile > 0) st_volatile64(P.tile_state + tile, ST_AGG | cntT);
    /* line ends of this thread's LFs: window offsets, in order (needs only the rank inside the tile).  On the way: every
     * "\n+\n" says that the line ending at its second LF is a plus line, i.e. proposes the line class of the tile's first byte. */
    const uint32_t c0 = LN_CPT * tid;
    const bool too_many = cntT > (uint32_t)LN_LMAX;
    if (too_many) anomaly |= LN_A_CAPACITY;
    else {
      uint32_t rank = excl, prop = 0;
      const uint32_t e0 = LN_LEFT + 16 * c0 + 1;
/* the first two LFs of a 32-byte span without a branch, a loop for the rare rest */
#define LN_EMIT1(w_, h_) { const uint32_t b = __ffs(w) - 1; const uint32_t e = e0 + 32 * (h_) + b; \
        if (w) { lend[rank] = (uint16_t)e; if (((w_) >> b) & 4u) { if (win[e] == '+') prop |= 1u << ((1u - rank) & 3u); } rank++; } w &= w - 1; }
#define LN_EMIT(w_, h_) { uint32_t w = (w_); LN_EMIT1(w_, h_) LN_EMIT1(w_, h_) while (w) LN_EMIT1(w_, h_) }
      LN_EMIT(mm.x, 0) LN_EMIT(mm.y, 1) LN_EMIT(mm.z, 2) LN_EMIT(mm.w, 3)
#undef LN_EMIT1
#undef LN_EMIT
      prop = __reduce_or_sync(FULL, prop);
      if (lane == 0 && prop) atomicOr(&s_guess, prop);
    }
    __syncthreads();

    /* ---- C0: line class of the tile's first byte.  The true value needs the number of lines in front of the tile (the sum of
     * the counts of all tiles before ours); when the plus lines of the tile agree on it we go on with their answer and check it
     * against the sum at the end of the tile, when the tiles in front have long published theirs. */
    uint32_t base_line = 0;
    bool have_base = false;
    uint32_t phi;
    {
      const uint32_t g = s_guess;
      if (g == 1u || g == 2u || g == 4u || g == 8u) phi = 31u - __clz(g);
      else { /* no witness (or witnesses that disagree): wait for the sum now */
        for (int look = (int)tile - 1;; look -= LN_THREADS) { /* trip count is uniform over the block */
          const int idx = look - tid;
          const unsigned long long v64 = idx >= 0 ? ln_wait_state(P.tile_state + idx, P.out) : ST_INCL;
          if (ln_lookback_round(v64, lane, warp, s_w3, s_w4, &base_line)) break;
          __syncthreads(); /* s_w3 / s_w4 are rewritten by the next round */
        }
        have_base = true;
        phi = (base_line + 4u - P.j0) & 3u;
      }
    }
    /* line numbers relative to the tile: class = number & 3, record of the tile = (number >> 2) - 1 */
    const uint32_t gbr = 4u + phi;
    const uint32_t g0t = gbr + excl;                /* line number at this thread's first byte */

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

    /* the state of one tile in front of ours, asked for now and looked at after D */
    unsigned long long pre64 = ST_INCL;
    if (!have_base && (int)tile - 1 - tid >= 0) pre64 = ld_volatile64(P.tile_state + ((int)tile - 1 - tid));
    __syncthreads();

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
    }

    /* ---- E0: lines in front of the tile (block-wide look-back, normally one round over states read before D) */
    if (!have_base) {
      bool first_round = true;
      for (int look = (int)tile - 1;; look -= LN_THREADS) {
        const int idx = look - tid;
        unsigned long long v64 = ST_INCL;
        if (idx >= 0) v64 = (first_round && (pre64 >> 62) != 0) ? pre64 : ln_wait_state(P.tile_state + idx, P.out);
        first_round = false;
        if (ln_lookback_round(v64, lane, warp, s_w3, s_w4, &base_line)) break;
        __syncthreads();
      }
      if (((base_line + 4u - P.j0) & 3u) != phi) anomaly |= LN_A_PHASE; /* the plus lines of the tile misled us: hand the chunk on */
    }
    if (tid == 0) {
      st_volatile64(P.tile_state + tile, ST_INCL | ((unsigned long long)base_line + cntT));
      s_next = atomicAdd(P.ticket, 1u);
      if (tile == P.ntiles - 1) {
        uint32_t cnt = base_line + cntT;
        if (P.virtual_end && P.n > 0 && win[nloc - 1] != '\n') { if (cnt < P.cap) P.line_end[cnt] = P.n; P.out[LN_O_VIRTUAL] = cnt; cnt++; }
        P.out[LN_O_LINES] = cnt; P.out[LN_O_CAPOVF] = cnt > P.cap ? 1u : 0u;
      }
    }
    if (!too_many) {
      /* line numbers shifted so that the chunk's first record starts at line 4: class = number & 3, record = (number >> 2) - 1 (j0 <= 4) */
      const uint32_t gb = base_line + 4u - P.j0;
      const uint32_t gofs = (uint32_t)(t0 - LN_LEFT); /* window offset → offset inside the chunk */
      /* ---- E1: line ends → global line index */
      for (uint32_t r = tid; r < cntT; r += LN_THREADS) { const uint32_t gi = base_line + r; if (gi < P.cap) P.line_end[gi] = gofs + lend[r]; }

      /* ---- E2: header and plus lines that start in this tile.  Line k of the tile (0..cntT) starts at lend[k-1]. */
      {
        const uint32_t kmin = win[LN_LEFT - 1] == '\n' ? 0u : 1u;
        const uint32_t tile_end = LN_LEFT + nv; /* lines starting at or beyond belong to the next tile (or do not exist) */
        uint32_t kh0 = kmin + ((0u - (gb + kmin)) & 3u);
        if (gb + kh0 < 4u) kh0 += 4;
        uint32_t kp0 = kmin + ((2u - (gb + kmin)) & 3u);
        if (gb + kp0 < 4u) kp0 += 4;
        const uint32_t nH = kh0 <= cntT ? (cntT - kh0) / 4 + 1 : 0, nP = kp0 <= cntT ? (cntT - kp0) / 4 + 1 : 0;
        for (uint32_t u = tid; u < nH + nP; u += LN_THREADS) {
          const bool is_hdr = u < nH;
          const uint32_t k = is_hdr ? kh0 + 4 * u : kp0 + 4 * (u - nH);
          const uint32_t s = k == 0 ? (uint32_t)LN_LEFT : (uint32_t)lend[k - 1];
          if (s >= tile_end) continue;
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
              if (nloc == (uint32_t)LN_WIN) atomicMin(P.out + LN_O_OVERLONG, base_line + k); /* no LF within 1 KiB: a line gzgets would split */
              continue; /* otherwise cut by the end of the data */
            }
          }
          const uint32_t hl = e - s;
          if (hl >= FQ_MAX_LABEL_LENGTH) { atomicMin(P.out + LN_O_OVERLONG, base_line + k); continue; }
          uint32_t nlen; uint64_t mem_len, hsh = FQ_HASH_SKIP;
          if (!fq_header_hash_fast(win, s, hl, P.cx.fmt_key, P.cx.pe_key, P.cx.seed, P.names != nullptr, &nlen, &mem_len, &hsh)) { anomaly |= LN_A_HEADER; continue; }
          const uint32_t rec = ((gb + k) >> 2) - 1u;
          if (P.names && rec < P.names_cap) {
            FqName nm; nm.off = gofs + s + 1; nm.len = nlen; nm.hash = hsh;
            P.names[rec] = nm;
          }
        }
      }
    }
  }

  /* ---- results of this thread → one set of atomics per warp */
  if ((seq_ok & 0x80808080u) != 0x80808080u) anomaly |= LN_A_BASE;
  uint32_t mn = min(qmn & 0xFFFFu, qmn >> 16), mx = max(qmx & 0xFFFFu, qmx >> 16);
  mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
  if (mn <= 0x0Du) anomaly |= LN_A_QUAL; /* NUL / LF / CR (or another control byte) inside a quality line: let the careful path look */
  anomaly = __reduce_or_sync(FULL, anomaly);
  if (lane == 0) {
    if (anomaly) atomicOr(P.out + LN_O_ANOMALY, anomaly);
    if (mn <= mx) { atomicMin(P.out + LN_O_QMIN, mn); atomicMax(P.out + LN_O_QMAX, mx); }
  }
}

/* ------------------------------------------------------------------------------------------------ K5r: record rules + statistics
 * One thread per record over the line index: over-long lines (src/fastq.h:30-37), read length >= 1 (src/fastq.c:346), equal
 * sequence / quality lengths (:380), fastq_new_entry_stats (:97-110) and the index bookkeeping (n_entries, index_mem).
 * sign = -1 takes the same counts back (a chunk that failed one of these rules is not committed). */
struct LanesRecParams {
  const uint32_t* line_end; uint32_t* out; uint32_t j0; const FqName* names; FqRecCtx cx;
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
  const int lane = threadIdx.x & 31;
  unsigned long long my_rds = 0, my_names = 0, my_mem = 0;
  uint32_t mn_rl = 0xFFFFFFFFu, mx_rl = 0, run_len = 0, run_cnt = 0, bad = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < nrec; base += stride) {
    const uint32_t r = base + threadIdx.x;
    uint32_t flush_len = 0, flush_cnt = 0;
    if (r < nrec) {
      const uint32_t j = P.j0 + 4 * r;
      const uint32_t s0 = j ? P.line_end[j - 1] : 0u, e0 = P.line_end[j], e1 = P.line_end[j + 1], e2 = P.line_end[j + 2], e3 = P.line_end[j + 3];
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

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
constexpr int LN_LMAX = 2048;  /* line ends kept per tile */
constexpr int LN_EMAX = 2048;  /* partial-chunk work items per tile */
constexpr int LN_OFF_MASK = LN_WIN;
constexpr int LN_OFF_LEND = LN_OFF_MASK + (LN_CHUNKS + 8) * 2;
constexpr int LN_OFF_PURE = LN_OFF_LEND + LN_LMAX * 2;
constexpr int LN_OFF_EDGE = LN_OFF_PURE + LN_CHUNKS * 2;
constexpr int LN_OFF_LUT = LN_OFF_EDGE + LN_EMAX * 4;
constexpr int LN_SMEM = LN_OFF_LUT + 32 * 16;
static_assert(LN_CPT == 8, "a thread's masks are one 16-byte load");
static_assert(LN_WIN % 16 == 0 && LN_OFF_LEND % 16 == 0 && LN_OFF_PURE % 16 == 0 && LN_OFF_EDGE % 16 == 0 && LN_OFF_LUT % 16 == 0, "alignment");

/* anomaly bits (out[3]) */
enum { LN_A_BASE = 1, LN_A_QUAL = 2, LN_A_HEADER = 4, LN_A_PLUS = 8, LN_A_CAPACITY = 16 };
/* out words */
enum { LN_O_LINES = 0, LN_O_CAPOVF = 1, LN_O_OVERLONG = 2, LN_O_ANOMALY = 3, LN_O_INTERNAL = 4, LN_O_VIRTUAL = 5, LN_O_QMIN = 6, LN_O_QMAX = 7,
       LN_O_RLMIN = 8, LN_O_RLMAX = 9, LN_O_RECBAD = 10, LN_O_WORDS = 12 };

struct LanesParams {
  const uint8_t* data;  /* 16-byte aligned (bulk copies); the chunk's first byte is data[lead] */
  uint32_t lead;        /* 0..15 bytes in front of the chunk that are not its data */
  uint32_t n;           /* lead + bytes of the chunk */
  int virtual_end; uint32_t* line_end; uint32_t cap;
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

/* Walk the byte ranges between the LFs of one chunk.  `m` = LF flags, `nvc` = valid bytes of the chunk (16 except at the end of
 * the data), `lo` = first valid byte (0 except at the start of an unaligned chunk), `cls` = line class there (updated).  F(lo, hi, cls) for every non-empty range, G(p) for every LF. */
template <typename FSeg, typename FLf>
__device__ __forceinline__ void ln_walk_chunk(uint32_t m, uint32_t lo, uint32_t nvc, uint32_t& cls, FSeg seg, FLf lf) {
  while (m) {
    uint32_t p = __ffs(m) - 1; m &= m - 1;
    if (p > lo) seg(lo, p, cls);
    lf(p);
    cls = (cls + 1u) & 3u;
    lo = p + 1;
  }
  if (lo < nvc) seg(lo, nvc, cls);
}

__global__ void __launch_bounds__(LN_THREADS, 4)
fq_lanes_kernel(const LanesParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* win = smem;
  uint16_t* maskbuf = (uint16_t*)(smem + LN_OFF_MASK);
  uint16_t* lend = (uint16_t*)(smem + LN_OFF_LEND);
  uint16_t* pure = (uint16_t*)(smem + LN_OFF_PURE);   /* sequence chunks from the front, quality chunks from the back */
  uint32_t* edge = (uint32_t*)(smem + LN_OFF_EDGE);   /* same, partial chunks: chunk | lo << 11 | (hi-1) << 15 */
  uint4* lut = (uint4*)(smem + LN_OFF_LUT);           /* [lo] bytes >= lo, [16 + h] bytes <= h */
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tile, s_w1[LN_WARPS], s_w2[LN_WARPS], s_w3[LN_WARPS], s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    __syncthreads(); /* everyone is done with the previous window and lists */
    if (tid == 0) {
      uint32_t t = atomicAdd(P.ticket, 1u);
      s_tile = t;
      if (t < P.ntiles) {
        unsigned long long t0 = (unsigned long long)t * LN_TILE;
        unsigned long long src = t ? t0 - LN_LEFT : 0;
        uint32_t dst_off = t ? 0 : LN_LEFT;
        unsigned long long want = (unsigned long long)LN_WIN - dst_off, have = ((unsigned long long)P.n - src + 15) & ~15ull;
        uint32_t bytes = (uint32_t)(want < have ? want : have);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(win + dst_off)), "l"(P.data + src), "r"(bytes), "r"(bar) : "memory");
      }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= P.ntiles) break;
    const unsigned long long t0 = (unsigned long long)tile * LN_TILE;
    const uint32_t left = (uint32_t)min((unsigned long long)(LN_TILE + LN_MARGIN), (unsigned long long)P.n - t0); /* data bytes from the tile start */
    const uint32_t nv = min(left, (uint32_t)LN_TILE);   /* valid bytes of the tile itself */
    const uint32_t nloc = LN_LEFT + left;               /* window offsets below this hold data */
    const bool full = nv == (uint32_t)LN_TILE;
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) { if (tid == 0) atomicExch(P.out + LN_O_INTERNAL, 1u); break; } }
      parity ^= 1;
    }
    const uint32_t lead = tile == 0 ? P.lead : 0u; /* bytes of the tile's first 16-byte chunk that precede the data */
    if (tile == 0 && tid == 0) { win[LN_LEFT - 1] = '\n'; win[LN_LEFT + lead - 1] = '\n'; }

    /* ---- A: LF flags; lane ↔ adjacent chunks (conflict-free 128-bit shared loads) */
    {
      const uint32_t cbase = warp * (LN_CHUNKS / LN_WARPS) + lane;
#pragma unroll
      for (int i = 0; i < LN_CPT; i++) {
        const uint32_t c = cbase + i * 32;
        uint32_t m = ln_lf_mask16(*(const uint4*)(win + LN_LEFT + 16 * c));
        if (!full) { uint32_t valid = nv > 16 * c ? min(16u, nv - 16 * c) : 0u; m &= (1u << valid) - 1u; }
        if (c == 0) m &= ~((1u << lead) - 1u);
        maskbuf[c] = (uint16_t)m;
      }
    }
    __syncwarp(); /* a thread's 8 consecutive chunks were flagged by its own warp */
    const uint4 mm = *(const uint4*)(maskbuf + LN_CPT * tid);
    const uint32_t tot = __popc(mm.x) + __popc(mm.y) + __popc(mm.z) + __popc(mm.w);

    /* ---- B: prefix of the LF counts inside the tile, look-back across tiles */
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) s_w1[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - tot, cntT = 0;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) { uint32_t x = s_w1[w]; cntT += x; if (w < warp) excl += x; }
    if (warp == 0) {
      unsigned long long acc = 0;
      if (tile > 0) {
        if (lane == 0) st_volatile64(P.tile_state + tile, ST_AGG | cntT);
        int look = (int)tile - 1;
        uint32_t spins = 0;
        for (;;) {
          int idx = look - lane;
          unsigned long long v64 = idx >= 0 ? ld_volatile64(P.tile_state + idx) : ST_INCL;
          while (__any_sync(FULL, (v64 >> 62) == 0)) {
            if ((v64 >> 62) == 0) v64 = ld_volatile64(P.tile_state + idx);
            if (++spins > (1u << 26)) { if (lane == 0) atomicExch(P.out + LN_O_INTERNAL, 2u); v64 = ST_INCL; }
          }
          uint32_t incl_mask = __ballot_sync(FULL, (v64 >> 62) == 2);
          int first = incl_mask ? __ffs(incl_mask) - 1 : 31;
          unsigned long long part = lane <= first ? (v64 & ST_VALUE) : 0ull;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(FULL, part, d);
          acc += part;
          if (incl_mask) break;
          look -= 32;
        }
      }
      if (lane == 0) {
        st_volatile64(P.tile_state + tile, ST_INCL | (acc + cntT));
        s_base = (uint32_t)acc;
        if (tile == P.ntiles - 1) {
          uint32_t cnt = (uint32_t)acc + cntT;
          if (P.virtual_end && P.n > P.lead && win[nloc - 1] != '\n') { if (cnt < P.cap) P.line_end[cnt] = P.n - P.lead; P.out[LN_O_VIRTUAL] = cnt; cnt++; }
          P.out[LN_O_LINES] = cnt; P.out[LN_O_CAPOVF] = cnt > P.cap ? 1u : 0u;
        }
      }
    }
    __syncthreads();
    const uint32_t base_line = s_base;
    const uint32_t cls0 = (base_line + excl - P.j0) & 3u; /* line class at this thread's first byte */
    const uint32_t c0 = LN_CPT * tid;
    const uint32_t mw[4] = {mm.x, mm.y, mm.z, mm.w};

    /* ---- C1: how many work items of each kind does this thread produce? */
    uint32_t n_pure = 0, n_edge = 0; /* sequence count in the low half, quality count in the high half */
    {
      uint32_t cls = cls0;
#pragma unroll
      for (int i = 0; i < LN_CPT; i++) {
        const uint32_t m = (mw[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        const uint32_t nvc = full ? 16u : (nv > 16 * (c0 + i) ? min(16u, nv - 16 * (c0 + i)) : 0u);
        const uint32_t lo0 = (i == 0 && tid == 0) ? lead : 0u;
        if (m == 0 && nvc == 16u && lo0 == 0u) { n_pure += (cls == 1u ? 1u : 0u) + (cls == 3u ? 0x10000u : 0u); }
        else ln_walk_chunk(m, lo0, nvc, cls,
                           [&](uint32_t, uint32_t, uint32_t k) { n_edge += (k == 1u ? 1u : 0u) + (k == 3u ? 0x10000u : 0u); },
                           [&](uint32_t) {});
      }
    }
    uint32_t ip = n_pure, ie = n_edge;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t a = __shfl_up_sync(FULL, ip, d), b = __shfl_up_sync(FULL, ie, d);
      if (lane >= d) { ip += a; ie += b; }
    }
    if (lane == 31) { s_w2[warp] = ip; s_w3[warp] = ie; }
    __syncthreads();
    uint32_t xp = ip - n_pure, xe = ie - n_edge, tp = 0, te = 0;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) { uint32_t a = s_w2[w], b = s_w3[w]; tp += a; te += b; if (w < warp) { xp += a; xe += b; } }
    const uint32_t n_seq_pure = tp & 0xFFFFu, n_qual_pure = tp >> 16, n_seq_edge = te & 0xFFFFu, n_qual_edge = te >> 16;
    const bool too_many = n_seq_edge + n_qual_edge > (uint32_t)LN_EMAX || cntT > (uint32_t)LN_LMAX;
    if (too_many) anomaly |= LN_A_CAPACITY;

    /* ---- C2: emit line ends and work items */
    if (!too_many) {
      uint32_t cls = cls0, rank = excl;
      uint32_t ps = xp & 0xFFFFu, pq = LN_CHUNKS - 1 - (xp >> 16), es = xe & 0xFFFFu, eq = LN_EMAX - 1 - (xe >> 16);
#pragma unroll
      for (int i = 0; i < LN_CPT; i++) {
        const uint32_t m = (mw[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        const uint32_t c = c0 + i;
        const uint32_t nvc = full ? 16u : (nv > 16 * c ? min(16u, nv - 16 * c) : 0u);
        const uint32_t lo0 = (i == 0 && tid == 0) ? lead : 0u;
        if (m == 0 && nvc == 16u && lo0 == 0u) {
          if (cls == 1u) pure[ps++] = (uint16_t)c;
          else if (cls == 3u) pure[pq--] = (uint16_t)c;
        } else
          ln_walk_chunk(m, lo0, nvc, cls,
                        [&](uint32_t lo, uint32_t hi, uint32_t k) {
                          uint32_t it = c | (lo << 11) | ((hi - 1u) << 15);
                          if (k == 1u) edge[es++] = it; else if (k == 3u) edge[eq--] = it;
                        },
                        [&](uint32_t p) { lend[rank++] = (uint16_t)(LN_LEFT + 16 * c + p + 1); });
      }
    }
    __syncthreads();

    if (!too_many) {
      /* ---- D: sequence alphabet, quality range; one predicate per warp instruction */
      for (uint32_t i = tid; i < n_seq_pure; i += LN_THREADS)
        seq_ok &= ln_pred4(*(const uint4*)(win + LN_LEFT + 16 * (uint32_t)pure[i]));
      for (uint32_t i = tid; i < n_seq_edge; i += LN_THREADS) {
        const uint32_t it = edge[i], c = it & 0x7FFu;
        const uint4 v = *(const uint4*)(win + LN_LEFT + 16 * c);
        const uint4 a = lut[(it >> 11) & 15u], b = lut[16u + ((it >> 15) & 15u)];
        /* bytes outside [lo, hi) always pass */
        seq_ok &= (fq_base_pred(v.x) | ~(a.x & b.x)) & (fq_base_pred(v.y) | ~(a.y & b.y)) & (fq_base_pred(v.z) | ~(a.z & b.z)) & (fq_base_pred(v.w) | ~(a.w & b.w));
      }
      for (uint32_t i = tid; i < n_qual_pure; i += LN_THREADS) {
        const uint4 v = *(const uint4*)(win + LN_LEFT + 16 * (uint32_t)pure[LN_CHUNKS - 1 - i]);
        ln_minmax_word(v.x, qmn, qmx); ln_minmax_word(v.y, qmn, qmx); ln_minmax_word(v.z, qmn, qmx); ln_minmax_word(v.w, qmn, qmx);
      }
      for (uint32_t i = tid; i < n_qual_edge; i += LN_THREADS) {
        const uint32_t it = edge[LN_EMAX - 1 - i], c = it & 0x7FFu, lo = (it >> 11) & 15u;
        const uint4 v = *(const uint4*)(win + LN_LEFT + 16 * c);
        const uint4 a = lut[lo], b = lut[16u + ((it >> 15) & 15u)];
        const uint32_t fill = (uint32_t)win[LN_LEFT + 16 * c + lo] * 0x01010101u; /* a byte of the range stands in for the bytes outside it */
        uint32_t m;
        m = a.x & b.x; ln_minmax_word((v.x & m) | (fill & ~m), qmn, qmx);
        m = a.y & b.y; ln_minmax_word((v.y & m) | (fill & ~m), qmn, qmx);
        m = a.z & b.z; ln_minmax_word((v.z & m) | (fill & ~m), qmn, qmx);
        m = a.w & b.w; ln_minmax_word((v.w & m) | (fill & ~m), qmn, qmx);
      }

      /* ---- E1: line ends of the tile → global line index */
      const uint32_t gofs = (uint32_t)(t0 - LN_LEFT) - P.lead; /* window offset → offset inside the chunk */
      for (uint32_t k = tid; k < cntT; k += LN_THREADS) { uint32_t gi = base_line + k; if (gi < P.cap) P.line_end[gi] = gofs + lend[k]; }

      /* ---- E2: header and plus lines that start in this tile.  Line k of the tile (0..cntT) starts at lend[k-1]. */
      {
        const uint32_t kmin = win[LN_LEFT - 1] == '\n' ? 0u : 1u;
        const uint32_t tile_end = LN_LEFT + nv; /* lines starting at or beyond belong to the next tile (or do not exist) */
        uint32_t kh0 = kmin + ((0u - (base_line + kmin - P.j0)) & 3u);
        if (base_line + kh0 < P.j0) kh0 += 4;
        uint32_t kp0 = kmin + ((2u - (base_line + kmin - P.j0)) & 3u);
        if (base_line + kp0 < P.j0) kp0 += 4;
        const uint32_t nH = kh0 <= cntT ? (cntT - kh0) / 4 + 1 : 0, nP = kp0 <= cntT ? (cntT - kp0) / 4 + 1 : 0;
        for (uint32_t u = tid; u < nH + nP; u += LN_THREADS) {
          const bool is_hdr = u < nH;
          const uint32_t k = is_hdr ? kh0 + 4 * u : kp0 + 4 * (u - nH);
          const uint32_t s = k == 0 ? (uint32_t)LN_LEFT + lead : (uint32_t)lend[k - 1];
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
          uint32_t nlen; uint64_t mem_len;
          if (!fq_header_fast(win, s, hl, P.cx.fmt_key, P.cx.pe_key, &nlen, &mem_len)) { anomaly |= LN_A_HEADER; continue; }
          const uint32_t rec = (base_line + k - P.j0) >> 2;
          if (P.names && rec < P.names_cap) {
            FqName nm; nm.off = gofs + s + 1; nm.len = nlen;
            nm.hash = fq_hash_name_words(win, s + 1, nlen, P.cx.seed);
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

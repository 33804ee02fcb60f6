/*
 * fq_cuda.cu — the sm_100a kernels of libfastq_gpu and the FqDevice that launches them on one CUDA stream.
 *
 *   K1  fq_scan_kernel          line index: coalesced 128-bit loads, SWAR newline masks, warp-shuffle prefix sums and a
 *                               single-pass decoupled look-back across tiles (replaces the 4×gzgets splitter, src/fastq.c:245-261)
 *   K1b fq_overlong_kernel      first line that gzgets would split (src/fastq.h:30-37 limits)
 *   K2  fq_records_kernel       reader flags, validation, statistics, event key, name hash (src/fastq.c:300-392, :442-516, :97-110)
 *   K3  fq_index_insert_kernel  open-addressing name index: 64-bit hash, atomicCAS claim, atomicMin of the record index,
 *                               exact byte compare on equal hashes (replaces src/hash.c + src/fastq.c:529-611)
 *   K4  fq_mate_claim_kernel    lookup-then-delete of the mate loop as a claim with atomicMin (src/fastq_info.c:333-350)
 *       fq_pair_compare_kernel  interleaved / sorted pair name equality (src/fastq_info.c:86-91, :133-138)
 *   single-thread helpers: gzgets emulation for over-long lines, first-record sniffers, error details.
 *
 * HBM-bound byte/integer work: no tensor cores.  Event ordering (first error wins) is a 64-bit atomicMin.
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/fastq_gpu.h"
#include "fq_device.h"
#include "fq_record.h"

#define FQ_CUDA_CHECK(call)                                                                            \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      std::string m_ = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
      if (e_ == cudaErrorMemoryAllocation) m_ += " (out of memory)";                                   \
      throw std::runtime_error(m_);                                                                    \
    }                                                                                                  \
  } while (0)

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int kSMs = 148;

/* ------------------------------------------------------------------------------------------------ K1: line index */
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_WARP_BYTES = 2048;                 /* 4 rows of 32 lanes × 16 B */
constexpr int SCAN_TILE = SCAN_WARPS * SCAN_WARP_BYTES; /* 16 KiB per CTA */
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INCL = 2ull << 62, ST_VALUE = (1ull << 62) - 1;

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
/* 16 bytes → 16-bit mask, bit b set iff byte b is LF.  Two words are merged before one multiply gathers their flag
 * bits: flags of word A sit at bits 0,8,16,24 and of word B at 4,12,20,28; × (1+2^7+2^14+2^21) lines them up in bits 21..28. */
__device__ __forceinline__ uint32_t lf_mask16(uint4 v) {
  uint32_t z0 = fq_zero_bytes(v.x ^ 0x0A0A0A0Au), z1 = fq_zero_bytes(v.y ^ 0x0A0A0A0Au);
  uint32_t z2 = fq_zero_bytes(v.z ^ 0x0A0A0A0Au), z3 = fq_zero_bytes(v.w ^ 0x0A0A0A0Au);
  uint32_t lo = (((z0 >> 7) | (z1 >> 3)) * 0x00204081u) >> 21;
  uint32_t hi = (((z2 >> 7) | (z3 >> 3)) * 0x00204081u) >> 21;
  return (lo & 0xFFu) | ((hi & 0xFFu) << 8);
}
/* interleaved flag layout of fq_count_kernel: byte k of word i sits at bit 8k+i; keep the bytes below `valid` (1..15) */
__device__ __forceinline__ uint32_t lf_mask_bits_below(uint32_t valid) {
  uint32_t m = 0;
  for (uint32_t b = 0; b < valid; b++) m |= 1u << (8u * (b & 3u) + (b >> 2));
  return m;
}
__device__ __forceinline__ unsigned long long ld_volatile64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS)
fq_scan_kernel(const uint8_t* __restrict__ data, uint32_t n, int virtual_end, uint32_t* __restrict__ line_end, uint32_t cap,
               unsigned long long* tile_state, uint32_t* ticket, uint32_t ntiles, uint32_t* out2, uint32_t lead) {
  __shared__ uint32_t s_tile, s_warp_tot[SCAN_WARPS], s_warp_base[SCAN_WARPS], s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u); /* tiles are claimed in order: look-back never waits on an unscheduled CTA */
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t wbase = tile * (uint32_t)SCAN_TILE + warp * SCAN_WARP_BYTES + lane * 16;
  uint32_t m[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t off = wbase + i * 512;
    m[i] = 0;
    if (off < n) { /* the allocation is readable 64 bytes past n */
      m[i] = lf_mask16(ld_stream16(data + off));
      if (n - off < 16) m[i] &= (1u << (n - off)) - 1u;
      if (off == 0) m[i] &= ~((1u << lead) - 1u); /* the first `lead` (< 16) bytes are not the chunk's own */
    }
  }
  /* inclusive prefix over the warp's four rows, two 16-bit counters per register */
  uint32_t c01 = __popc(m[0]) | (__popc(m[1]) << 16), c23 = __popc(m[2]) | (__popc(m[3]) << 16);
  uint32_t p01 = c01, p23 = c23;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t a = __shfl_up_sync(FULL, p01, d), b = __shfl_up_sync(FULL, p23, d);
    if (lane >= d) { p01 += a; p23 += b; }
  }
  uint32_t t01 = __shfl_sync(FULL, p01, 31), t23 = __shfl_sync(FULL, p23, 31);
  uint32_t rowbase[4] = {0, t01 & 0xFFFFu, (t01 & 0xFFFFu) + (t01 >> 16), (t01 & 0xFFFFu) + (t01 >> 16) + (t23 & 0xFFFFu)};
  uint32_t wtot = rowbase[3] + (t23 >> 16);
  if (lane == 0) s_warp_tot[warp] = wtot;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = lane < SCAN_WARPS ? s_warp_tot[lane] : 0, incl = v;
#pragma unroll
    for (int d = 1; d < SCAN_WARPS; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
    if (lane < SCAN_WARPS) s_warp_base[lane] = incl - v;
    uint32_t total = __shfl_sync(FULL, incl, SCAN_WARPS - 1);
    /* decoupled look-back: publish this tile's count, then add up the predecessors' */
    unsigned long long acc = 0;
    if (tile > 0) {
      if (lane == 0) st_volatile64(tile_state + tile, ST_AGG | total);
      int look = (int)tile - 1;
      for (;;) {
        int idx = look - lane;
        unsigned long long v64 = idx >= 0 ? ld_volatile64(tile_state + idx) : ST_INCL;
        while (__any_sync(FULL, (v64 >> 62) == 0)) { if ((v64 >> 62) == 0) v64 = ld_volatile64(tile_state + idx); }
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
      st_volatile64(tile_state + tile, ST_INCL | (acc + total));
      s_base = (uint32_t)acc;
      if (tile == ntiles - 1) {
        uint32_t cnt = (uint32_t)acc + total;
        if (virtual_end && n > 0 && data[n - 1] != '\n') { if (cnt < cap) line_end[cnt] = n; cnt++; }
        out2[0] = cnt; out2[1] = cnt > cap ? 1u : 0u;
      }
    }
  }
  __syncthreads();
  const uint32_t base = s_base + s_warp_base[warp];
  const uint32_t incl[4] = {p01 & 0xFFFFu, p01 >> 16, p23 & 0xFFFFu, p23 >> 16};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t mm = m[i];
    if (mm) {
      uint32_t rank = base + rowbase[i] + incl[i] - __popc(mm);
      uint32_t off = wbase + i * 512;
      while (mm) {
        uint32_t b = __ffs(mm) - 1; mm &= mm - 1;
        if (rank < cap) line_end[rank] = off + b + 1;
        rank++;
      }
    }
  }
}

/* LF count only (multi-GPU prescan: the line phase of a byte range must be known before it can be validated) */
__global__ void __launch_bounds__(256)
fq_count_kernel(const uint8_t* __restrict__ data, uint32_t n, unsigned long long* out) {
  unsigned long long c = 0;
  const uint32_t nchunks = (n + 15) / 16;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nchunks; i += gridDim.x * blockDim.x) {
    uint4 v = ld_stream16(data + (size_t)i * 16);
    uint32_t z = (fq_zero_bytes(v.x ^ 0x0A0A0A0Au) >> 7) | (fq_zero_bytes(v.y ^ 0x0A0A0A0Au) >> 6) |
                 (fq_zero_bytes(v.z ^ 0x0A0A0A0Au) >> 5) | (fq_zero_bytes(v.w ^ 0x0A0A0A0Au) >> 4);
    if (n - i * 16 < 16) z &= lf_mask_bits_below(n - i * 16);
    c += __popc(z);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void fq_overlong_kernel(const uint32_t* __restrict__ line_end, uint32_t q, uint32_t j0, uint32_t nlines, uint32_t n,
                                   int tail_from_n, uint32_t* out) {
  uint32_t total = nlines + (tail_from_n ? 1u : 0u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t start = i == 0 ? q : line_end[j0 + i - 1];
    uint32_t end = i < nlines ? line_end[j0 + i] : n;
    uint32_t lim = (i & 1u) == 0 ? FQ_MAX_LABEL_LENGTH : FQ_MAX_READ_LENGTH;
    if (end - start >= lim) atomicMin(out, j0 + i);
  }
}

/* gzgets emulation (zlib: at most max-1 bytes, stops after LF) for the four lines of one record */
__global__ void fq_split_serial_kernel(const uint8_t* data, uint32_t n, uint32_t q, int is_eof, FqLine* lines4, uint32_t* out3) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t p = q, got = 0, lfs = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t maxb = ((i & 1) == 0 ? FQ_MAX_LABEL_LENGTH : FQ_MAX_READ_LENGTH) - 1, k = 0;
    bool lf = false;
    while (k < maxb && p + k < n) { k++; if (data[p + k - 1] == '\n') { lf = true; break; } }
    bool complete = lf || k == maxb || (is_eof && k > 0);
    if (!complete) break;
    lines4[i].off = p; lines4[i].len = k; p += k; got++; lfs += lf ? 1 : 0;
  }
  for (uint32_t i = got; i < 4; i++) { lines4[i].off = p; lines4[i].len = 0; }
  out3[0] = p; out3[1] = got; out3[2] = lfs;
}

__global__ void fq_sniff_kernel(const uint8_t* data, FqLine hdr1, FqLine seq, int32_t* out2) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t cl0 = fq_cstrlen(data, hdr1.off, hdr1.len), cl1 = fq_cstrlen(data, seq.off, seq.len);
  out2[0] = fq_sniff_format(data + hdr1.off + 1, cl0 >= 1 ? cl0 - 1 : 0);
  out2[1] = fq_sniff_colorspace(data + seq.off, cl1);
}

__global__ void fq_explain_kernel(const uint8_t* data, FqLine l0, FqLine l1, FqLine l2, FqLine l3, FqRecCtx cx, FqRecOut* out) {
  if (threadIdx.x || blockIdx.x) return;
  FqLine L[4] = {l0, l1, l2, l3};
  FqRecOut o;
  fq_check_record_careful(data, L, cx, &o);
  *out = o;
}

/* ------------------------------------------------------------------------------------------------ K2: records */
constexpr int REC_THREADS = 128;

__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { unsigned long long o = __shfl_xor_sync(FULL, v, d); v = o < v ? o : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}
/* add `count` to hist[len] for every lane with count > 0, one atomic per distinct length in the warp */
__device__ __forceinline__ void hist_flush(unsigned long long* hist, uint32_t len, uint32_t count) {
  unsigned active = __ballot_sync(FULL, count > 0);
  if (!active) return;
  if (count > 0) {
    unsigned peers = __match_any_sync(active, len);
    unsigned long long sum = 0;
    for (unsigned p = peers; p; p &= p - 1) sum += __shfl_sync(peers, count, __ffs(p) - 1);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(hist + len, sum);
  }
}

struct RecParams {
  const uint8_t* data; const uint32_t* line_end; const FqLine* lines;
  uint32_t q, j0, nrec; unsigned long long g0, step_base;
  FqRecCtx cx;
  FqStats* stats; FqStats* stats_range; unsigned long long* hist; unsigned long long* key; FqName* names;
};

__global__ void __launch_bounds__(REC_THREADS)
fq_records_kernel(const RecParams P) {
  const int lane = threadIdx.x & 31;
  unsigned long long my_key = FQ_KEY_NONE, my_rds = 0, my_names = 0, my_mem = 0;
  uint32_t mn_rl = 0xFFFFFFFFu, mx_rl = 0, mn_q = 255, mx_q = 0;
  uint32_t run_len = 0, run_cnt = 0; /* run-length cache for the histogram: reads of one length are the common case */
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < P.nrec; base += stride) {
    uint32_t k = base + threadIdx.x;
    bool valid = k < P.nrec;
    uint32_t flush_len = 0, flush_cnt = 0;
    if (valid) {
      FqLine L[4];
      if (P.lines) { for (int i = 0; i < 4; i++) L[i] = P.lines[4 * (size_t)k + i]; }
      else {
        uint32_t j = P.j0 + 4 * k;
        uint32_t s = k == 0 ? P.q : P.line_end[j - 1];
#pragma unroll
        for (int i = 0; i < 4; i++) { uint32_t e = P.line_end[j + i]; L[i].off = s; L[i].len = e - s; s = e; }
      }
      FqRecOut o;
      uint64_t hsh;
      fq_check_record(P.data, L, P.cx, &o, &hsh);
      unsigned long long g = P.g0 + k;
      unsigned long long key = fq_record_key(P.cx.loop, g, P.step_base, o);
      if (key < my_key) my_key = key;
      bool named = fq_record_has_name(P.cx.loop, o);
      if (P.names) {
        FqName nm; nm.off = o.name_off; nm.len = o.name_len;
        nm.hash = named ? hsh : FQ_HASH_SKIP;
        P.names[k] = nm;
      }
      if (P.cx.loop == FQ_LOOP_INDEX && named) { my_names++; my_mem += o.mem_len; }
      if (!o.flags && (o.vrank == FQ_V_OK || P.cx.loop == FQ_LOOP_READER)) { /* statistics are only reported when every record is clean */
        my_rds += P.cx.weight;
        mn_rl = min(mn_rl, o.read_len); mx_rl = max(mx_rl, o.read_len);
        if (o.qmin <= o.qmax) { mn_q = min(mn_q, o.qmin); mx_q = max(mx_q, o.qmax); }
        if (run_cnt && o.read_len != run_len) { flush_len = run_len; flush_cnt = run_cnt; run_cnt = 0; }
        run_len = o.read_len; run_cnt += P.cx.weight;
      }
    }
    hist_flush(P.hist, flush_len, flush_cnt);
  }
  hist_flush(P.hist, run_len, run_cnt);
  /* one set of atomics per warp */
  my_key = warp_min64(my_key);
  my_rds = warp_sum64(my_rds); my_names = warp_sum64(my_names); my_mem = warp_sum64(my_mem);
  mn_rl = __reduce_min_sync(FULL, mn_rl); mx_rl = __reduce_max_sync(FULL, mx_rl);
  mn_q = __reduce_min_sync(FULL, mn_q); mx_q = __reduce_max_sync(FULL, mx_q);
  if (lane == 0) {
    if (my_key != FQ_KEY_NONE) atomicMin(P.key, my_key);
    if (my_rds) atomicAdd(&P.stats->num_rds, my_rds);
    if (my_names) { atomicAdd(&P.stats->n_names, my_names); atomicAdd(&P.stats->mem_sum, my_mem); }
    if (mx_rl) { atomicMin(&P.stats_range->min_rl, mn_rl); atomicMax(&P.stats_range->max_rl, mx_rl); }
    if (mn_q <= mx_q) { atomicMin(&P.stats_range->min_q, mn_q); atomicMax(&P.stats_range->max_q, mx_q); }
  }
}

/* ------------------------------------------------------------------------------------------------ K1+K2 fused: one pass over HBM
 * A persistent CTA claims 32 KiB tiles in order.  The tile plus a 4 KiB look-ahead margin (and the 16 bytes before it) is
 * brought into shared memory by ONE bulk async copy (cp.async.bulk → UBLKCP, completion on an mbarrier).  From shared memory:
 * LF masks → block prefix → decoupled look-back gives the tile's global line number → the line ends go out to the line index,
 * and every record that STARTS in the tile is validated in place (one thread per record, fq_check_record on the shared window).
 * Records whose four lines do not fit the window are counted in out[3]; the host then falls back to the two-pass path. */
constexpr int TILE_BYTES = 32768, TILE_MARGIN = 4096, TILE_LEFT = 16;
constexpr int TILE_WIN = TILE_LEFT + TILE_BYTES + TILE_MARGIN;
constexpr int TILE_THREADS = 192;
constexpr int TILE_LMAX = 2048;
constexpr int TILE_CHUNKS_PER_THREAD = (TILE_BYTES + TILE_MARGIN) / 16 / TILE_THREADS; /* 12 */
constexpr int TILE_SMEM = TILE_WIN + 48 + TILE_LMAX * 2;
static_assert((TILE_BYTES + TILE_MARGIN) / 16 % TILE_THREADS == 0, "chunks must divide evenly");

struct TileParams {
  const uint8_t* data; uint32_t n; int virtual_end; uint32_t* line_end; uint32_t cap;
  unsigned long long* tile_state; uint32_t* ticket; uint32_t ntiles;
  uint32_t* out; /* [0] lines, [1] cap overflow, [2] first over-long header line, [3] records that did not fit, [4] internal error,
                    [5] first line whose end was stored by the tail tiles (line ends below 8 are stored too) */
  uint32_t j0, max_rec; unsigned long long g0, step_base; FqRecCtx cx;
  FqStats* stats; FqStats* stats_range; unsigned long long* hist; unsigned long long* key; FqName* names; uint32_t names_cap;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(TILE_THREADS)
fq_tile_kernel(const TileParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* win = smem;
  uint16_t* lend = (uint16_t*)(smem + TILE_WIN + 48);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tile, s_warp_all[TILE_THREADS / 32], s_warp_T[TILE_THREADS / 32], s_base, s_nl, s_cntT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t parity = 0;
  unsigned long long my_key = FQ_KEY_NONE, my_rds = 0, my_names = 0, my_mem = 0;
  uint32_t mn_rl = 0xFFFFFFFFu, mx_rl = 0, mn_q = 255, mx_q = 0, run_len = 0, run_cnt = 0;

  for (;;) {
    __syncthreads(); /* everyone is done with the previous window */
    if (tid == 0) {
      uint32_t t = atomicAdd(P.ticket, 1u);
      s_tile = t;
      if (t < P.ntiles) {
        unsigned long long t0 = (unsigned long long)t * TILE_BYTES;
        unsigned long long src = t ? t0 - TILE_LEFT : 0;
        uint32_t dst_off = t ? 0 : TILE_LEFT;
        unsigned long long want = (unsigned long long)TILE_WIN - dst_off, have = ((unsigned long long)P.n - src + 15) & ~15ull;
        uint32_t bytes = (uint32_t)(want < have ? want : have);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(win + dst_off)), "l"(P.data + src), "r"(bytes), "r"(bar) : "memory");
      }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= P.ntiles) break;
    const unsigned long long t0 = (unsigned long long)tile * TILE_BYTES;
    /* window offset w ↔ global offset t0 - 16 + w; real data below nloc */
    const uint32_t nloc = (uint32_t)min((unsigned long long)TILE_WIN, (unsigned long long)TILE_LEFT + (P.n - t0));
    {
      uint32_t spins = 0;
      while (!mbar_try_wait(bar, parity)) { if (++spins > (1u << 24)) { if (tid == 0) atomicExch(P.out + 4, 1u); break; } }
      parity ^= 1;
    }
    if (tile == 0 && tid == 0) win[TILE_LEFT - 1] = '\n';
    /* LF masks of this thread's 18 consecutive 16-byte chunks (window offsets 16 + 288*tid ...) */
    uint32_t mk[TILE_CHUNKS_PER_THREAD / 2];
    uint32_t c_all = 0, c_T = 0;
    const uint32_t w0 = TILE_LEFT + tid * (TILE_CHUNKS_PER_THREAD * 16);
    if (nloc == (uint32_t)TILE_WIN) { /* every tile but the last few: no bounds to check */
#pragma unroll
      for (int i = 0; i < TILE_CHUNKS_PER_THREAD; i += 2) {
        uint32_t m0 = lf_mask16(*(const uint4*)(win + w0 + i * 16)), m1 = lf_mask16(*(const uint4*)(win + w0 + i * 16 + 16));
        mk[i >> 1] = m0 | (m1 << 16);
      }
      /* a thread's 12 chunks lie on one side of the tile / margin border except for one thread */
      const uint32_t border = TILE_LEFT + TILE_BYTES;
      if (w0 + TILE_CHUNKS_PER_THREAD * 16 <= border) {
#pragma unroll
        for (int j = 0; j < TILE_CHUNKS_PER_THREAD / 2; j++) c_all += __popc(mk[j]);
        c_T = c_all;
      } else {
#pragma unroll
        for (int i = 0; i < TILE_CHUNKS_PER_THREAD; i++) {
          uint32_t c = __popc((i & 1) ? (mk[i >> 1] >> 16) : (mk[i >> 1] & 0xFFFFu));
          c_all += c;
          if (w0 + i * 16 < border) c_T += c;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < TILE_CHUNKS_PER_THREAD; i++) {
        uint32_t w = w0 + i * 16, m = 0;
        if (w < nloc) {
          m = lf_mask16(*(const uint4*)(win + w));
          if (nloc - w < 16) m &= (1u << (nloc - w)) - 1u;
        }
        if (i & 1) mk[i >> 1] |= m << 16; else mk[i >> 1] = m;
        uint32_t c = __popc(m);
        c_all += c;
        if (w < TILE_LEFT + TILE_BYTES) c_T += c;
      }
    }
    /* block-wide exclusive prefix of c_all, total of c_T */
    uint32_t incl = c_all, sumT = c_T;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sumT += __shfl_xor_sync(FULL, sumT, d);
    if (lane == 31) s_warp_all[warp] = incl;
    if (lane == 0) s_warp_T[warp] = sumT;
    __syncthreads();
    uint32_t rank = incl - c_all;
    for (int wgt = 0; wgt < warp; wgt++) rank += s_warp_all[wgt];
    if (warp == 0) {
      uint32_t nl = 0, cntT = 0;
      for (int wgt = 0; wgt < TILE_THREADS / 32; wgt++) { nl += s_warp_all[wgt]; cntT += s_warp_T[wgt]; }
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
            if (++spins > (1u << 26)) { if (lane == 0) atomicExch(P.out + 4, 2u); v64 = ST_INCL; }
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
        s_base = (uint32_t)acc; s_nl = nl; s_cntT = cntT;
        if (tile + 2 == P.ntiles || (P.ntiles == 1 && tile == 0)) P.out[5] = (uint32_t)acc;
        if (tile == P.ntiles - 1) {
          uint32_t cnt = (uint32_t)acc + cntT;
          if (P.virtual_end && P.n > 0 && win[nloc - 1] != '\n') { if (cnt < P.cap) P.line_end[cnt] = P.n; cnt++; }
          P.out[0] = cnt; P.out[1] = cnt > P.cap ? 1u : 0u;
        }
      }
    }
    __syncthreads();
    const uint32_t base_line = s_base, nl_win = s_nl, cntT = s_cntT;
    /* line ends: window-relative list in shared memory, global offsets into the line index */
    {
      const uint32_t gofs = (uint32_t)(t0 - TILE_LEFT);
      /* the host only ever needs the first and the last few line ends of a chunk it validated in this pass: the full line
       * index (4 bytes per line) is not written; the two-pass path rebuilds it on demand */
      const bool keep_all = tile + 2 >= P.ntiles;
#pragma unroll
      for (int j = 0; j < TILE_CHUNKS_PER_THREAD / 2; j++) {
        uint32_t m = mk[j];
        const uint32_t w = w0 + j * 32; /* bit b of the pair of chunks ↔ window offset w + b */
        while (m) {
          uint32_t b = __ffs(m) - 1; m &= m - 1;
          uint32_t e = w + b + 1;
          if (rank < TILE_LMAX) lend[rank] = (uint16_t)e;
          if (rank < cntT) { uint32_t gi = base_line + rank; if ((keep_all || gi < 8u) && gi < P.cap) P.line_end[gi] = gofs + e; }
          rank++;
        }
      }
    }
    __syncthreads();
    /* records that start in this tile */
    {
      const bool starts_here = win[TILE_LEFT - 1] == '\n';
      uint32_t lo = starts_here ? 0u : 1u;
      if (P.j0 > base_line + lo) lo = P.j0 - base_line;
      const uint32_t k0 = lo + ((4u - ((base_line + lo - P.j0) & 3u)) & 3u);
      const uint32_t nl_list = nl_win < (uint32_t)TILE_LMAX ? nl_win : (uint32_t)TILE_LMAX;
      const bool at_data_end = nloc < (uint32_t)TILE_WIN || t0 + TILE_BYTES + TILE_MARGIN >= P.n;
      const uint32_t nrec_tile = k0 <= cntT ? (cntT - k0) / 4 + 1 : 0;
      for (uint32_t rb = 0; rb < nrec_tile; rb += TILE_THREADS) { /* trip count is uniform over the block */
        uint32_t r = rb + tid;
        uint32_t flush_len = 0, flush_cnt = 0;
        if (r < nrec_tile) {
          uint32_t k = k0 + 4 * r;
          uint32_t start = k == 0 ? (uint32_t)TILE_LEFT : (uint32_t)lend[k - 1 < (uint32_t)TILE_LMAX ? k - 1 : 0];
          bool in_list = k == 0 || k - 1 < nl_list;
          if (in_list && start < TILE_LEFT + TILE_BYTES) {
            uint32_t g_local = (base_line + k - P.j0) >> 2;
            if (g_local < P.max_rec) {
              if (k + 3 >= nl_list) { /* the record's last line end is not in the window */
                bool virtual_last = at_data_end && P.virtual_end && k + 3 == nl_list && win[nloc - 1] != '\n' && nl_win <= (uint32_t)TILE_LMAX;
                if (virtual_last) {
                  /* file ends without LF: the fourth line ends at the end of the data */
                } else if (!at_data_end || nl_win > (uint32_t)TILE_LMAX) atomicAdd(P.out + 3, 1u);
                if (!virtual_last) goto next_record;
              }
              {
                FqLine L[4];
                uint32_t s = start;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  uint32_t e = (k + i < nl_list) ? (uint32_t)lend[k + i] : nloc;
                  L[i].off = s; L[i].len = e - s; s = e;
                }
                if (L[0].len >= FQ_MAX_LABEL_LENGTH) atomicMin(P.out + 2, base_line + k);
                if (L[2].len >= FQ_MAX_LABEL_LENGTH) atomicMin(P.out + 2, base_line + k + 2);
                FqRecOut o;
                uint64_t hsh;
                fq_check_record(win, L, P.cx, &o, &hsh);
                unsigned long long g = P.g0 + g_local;
                unsigned long long key = fq_record_key(P.cx.loop, g, P.step_base, o);
                if (key < my_key) my_key = key;
                bool named = fq_record_has_name(P.cx.loop, o);
                if (P.names && g_local < P.names_cap) {
                  FqName nm; nm.off = o.name_off + (uint32_t)(t0 - TILE_LEFT); nm.len = o.name_len;
                  nm.hash = named ? hsh : FQ_HASH_SKIP;
                  P.names[g_local] = nm;
                }
                if (P.cx.loop == FQ_LOOP_INDEX && named) { my_names++; my_mem += o.mem_len; }
                if (!o.flags && (o.vrank == FQ_V_OK || P.cx.loop == FQ_LOOP_READER)) {
                  my_rds += P.cx.weight;
                  mn_rl = min(mn_rl, o.read_len); mx_rl = max(mx_rl, o.read_len);
                  if (o.qmin <= o.qmax) { mn_q = min(mn_q, o.qmin); mx_q = max(mx_q, o.qmax); }
                  if (run_cnt && o.read_len != run_len) { flush_len = run_len; flush_cnt = run_cnt; run_cnt = 0; }
                  run_len = o.read_len; run_cnt += P.cx.weight;
                }
              }
            }
          }
        }
      next_record:
        hist_flush(P.hist, flush_len, flush_cnt);
      }
    }
  }
  hist_flush(P.hist, run_len, run_cnt);
  my_key = warp_min64(my_key);
  my_rds = warp_sum64(my_rds); my_names = warp_sum64(my_names); my_mem = warp_sum64(my_mem);
  mn_rl = __reduce_min_sync(FULL, mn_rl); mx_rl = __reduce_max_sync(FULL, mx_rl);
  mn_q = __reduce_min_sync(FULL, mn_q); mx_q = __reduce_max_sync(FULL, mx_q);
  if (lane == 0) {
    if (my_key != FQ_KEY_NONE) atomicMin(P.key, my_key);
    if (my_rds) atomicAdd(&P.stats->num_rds, my_rds);
    if (my_names) { atomicAdd(&P.stats->n_names, my_names); atomicAdd(&P.stats->mem_sum, my_mem); }
    if (mx_rl) { atomicMin(&P.stats_range->min_rl, mn_rl); atomicMax(&P.stats_range->max_rl, mx_rl); }
    if (mn_q <= mx_q) { atomicMin(&P.stats_range->min_q, mn_q); atomicMax(&P.stats_range->max_q, mx_q); }
  }
}

#include "fq_lanes.cuh"

/* ------------------------------------------------------------------------------------------------ K3 / K4: the index */
/* exact compare of two names, four bytes at a time whatever their alignment (aligned loads + funnel shifts; reads at most 7 bytes
 * past a name: chunks, bridges and blobs are padded) */
__device__ __forceinline__ bool names_equal(const uint8_t* a, const uint8_t* b, uint32_t n) {
  const uint32_t* wa = (const uint32_t*)((uintptr_t)a & ~(uintptr_t)3);
  const uint32_t* wb = (const uint32_t*)((uintptr_t)b & ~(uintptr_t)3);
  const uint32_t sa = ((uint32_t)(uintptr_t)a & 3u) * 8u, sb = ((uint32_t)(uintptr_t)b & 3u) * 8u;
  uint32_t ca = wa[0], cb = wb[0];
  for (uint32_t i = 0; i < n; i += 4) {
    const uint32_t na = wa[(i >> 2) + 1], nb = wb[(i >> 2) + 1];
    uint32_t diff = __funnelshift_r(ca, na, sa) ^ __funnelshift_r(cb, nb, sb);
    ca = na; cb = nb;
    if (n - i < 4) diff &= (1u << (8u * (n - i))) - 1u;
    if (diff) return false;
  }
  return true;
}
/* two names of equal length n in arenas: runs of 16-byte units, zero padded, 16-byte aligned — equal names are equal unit by unit */
__device__ __forceinline__ bool names_equal16(const uint8_t* a, const uint8_t* b, uint32_t n) {
  const uint4* ua = (const uint4*)a; const uint4* ub = (const uint4*)b;
  for (uint32_t u = 0; 16u * u < n; u++) { const uint4 x = ua[u], y = ub[u]; if ((x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w)) return false; }
  return true;
}
__device__ __forceinline__ const uint8_t* dir_name(const FqDirEntry* dir, uint32_t nd, unsigned long long g, uint32_t* len) {
  uint32_t lo = 0, hi = nd;
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (dir[mid].g0 <= g) lo = mid; else hi = mid; }
  const FqName nm = dir[lo].names[g - dir[lo].g0];
  *len = nm.len;
  return dir[lo].data + nm.off;
}

struct TableParams {
  const FqName* names; const uint8_t* data; uint32_t nrec; unsigned long long g0, step_base;
  FqSlot* slots; unsigned long long mask; const FqDirEntry* dir1; uint32_t ndir1;
  unsigned long long* key; unsigned long long* counters;
};

/* {hash, idx1} of a slot claimed in ONE 128-bit compare-and-swap (ATOMG.CAS.128): an insert into a free slot — the common case
 * of a table at most half full — is a single atomic on a single DRAM sector.  false: *old_hash / *old_idx hold what was there. */
__device__ __forceinline__ bool slot_claim128(FqSlot* s, unsigned long long hash, unsigned long long idx, unsigned long long* old_hash, unsigned long long* old_idx) {
  unsigned long long olo, ohi;
  asm volatile("{ .reg .b128 e, n, o;\n"
               "  mov.b128 e, {%3, %4};\n"
               "  mov.b128 n, {%5, %6};\n"
               "  atom.global.cas.b128 o, [%2], e, n;\n"
               "  mov.b128 {%0, %1}, o; }"
               : "=l"(olo), "=l"(ohi) : "l"(s), "l"(FQ_HASH_EMPTY), "l"(FQ_IDX_NONE), "l"(hash), "l"(idx) : "memory");
  *old_hash = olo; *old_idx = ohi;
  return olo == FQ_HASH_EMPTY && ohi == FQ_IDX_NONE;
}

__global__ void __launch_bounds__(256)
fq_index_insert_kernel(const TableParams P) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < P.nrec; k += gridDim.x * blockDim.x) {
    const FqName nm = P.names[k];
    if (nm.hash == FQ_HASH_SKIP) continue;
    const unsigned long long g = P.g0 + k;
    unsigned long long i = nm.hash & P.mask, probes = 0;
    for (;; i = (i + 1) & P.mask) {
      if (++probes > P.mask) { atomicExch(P.counters + 2, 1ull); break; }
      FqSlot* s = P.slots + i;
      unsigned long long cur, cur_idx;
      if (slot_claim128(s, nm.hash, g, &cur, &cur_idx)) break; /* first arrival of this name */
      if (cur != nm.hash) continue;
      /* the slot carries this hash: every record that ever held it has the SAME name (checked on the bytes on arrival), so any of
       * them — the one the failed claim returned — tells whether this is our name.  Another name with the same 64-bit hash
       * (the reference's hashit collision, src/hash.c:38-45 walks on to the next object): so do we, to the next slot. */
      uint32_t ol; const uint8_t* on = dir_name(P.dir1, P.ndir1, cur_idx, &ol);
      if (!(ol == nm.len && names_equal16(on, P.data + nm.off, nm.len))) { atomicAdd(P.counters + 4, 1ull); continue; } /* [4]: diagnostics only */
      /* keep the smallest record index; min over all arrivals of max(old, g) = 2nd smallest of the group = the reference's duplicate */
      unsigned long long old = atomicMin(&s->idx1, g);
      unsigned long long later = old > g ? old : g;
      atomicMin(P.key, FQ_KEY(P.step_base + later, FQ_R_NAME));
      break;
    }
  }
}

__global__ void __launch_bounds__(256)
fq_mate_claim_kernel(const TableParams P) {
  unsigned long long claimed = 0;
  for (uint32_t base = blockIdx.x * blockDim.x; base < P.nrec; base += gridDim.x * blockDim.x) {
    uint32_t k = base + threadIdx.x;
    if (k >= P.nrec) continue;
    const FqName nm = P.names[k];
    if (nm.hash == FQ_HASH_SKIP) continue;
    const unsigned long long g = P.g0 + k;
    unsigned long long i = nm.hash & P.mask, probes = 0, unpaired = FQ_IDX_NONE;
    for (;; i = (i + 1) & P.mask) {
      if (++probes > P.mask + 1) { unpaired = g; break; }
      const FqSlot* s = P.slots + i;
      unsigned long long cur = s->hash;
      if (cur == FQ_HASH_EMPTY) { unpaired = g; break; }
      if (cur != nm.hash) continue;
      uint32_t ol; const uint8_t* on = dir_name(P.dir1, P.ndir1, s->idx1, &ol);
      if (!(ol == nm.len && names_equal16(on, P.data + nm.off, nm.len))) continue; /* another name with this hash: the lookup walks on */
      unsigned long long old = atomicMin(&P.slots[i].claim2, g);
      if (old == FQ_IDX_NONE) claimed++;           /* first claim = the reference's delete */
      else unpaired = old > g ? old : g;           /* the entry was already deleted when the later one arrives */
      break;
    }
    if (unpaired != FQ_IDX_NONE) atomicMin(P.key, FQ_KEY(P.step_base + unpaired, FQ_R_NAME));
  }
  __syncwarp();
  claimed = warp_sum64(claimed);
  if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(P.counters + 1, claimed);
}

struct PairParams {
  const FqName* a; const uint8_t* da; uint32_t stride_a; const FqName* b; const uint8_t* db; uint32_t stride_b;
  uint32_t npairs; unsigned long long p0; uint32_t rank; unsigned long long* key;
};
__global__ void __launch_bounds__(256)
fq_pair_compare_kernel(const PairParams P) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < P.npairs; k += gridDim.x * blockDim.x) {
    const FqName x = P.a[(size_t)k * P.stride_a], y = P.b[(size_t)k * P.stride_b];
    if (x.hash == FQ_HASH_SKIP || y.hash == FQ_HASH_SKIP) continue;
    bool same = x.hash == y.hash && x.len == y.len && names_equal(P.da + x.off, P.db + y.off, x.len);
    if (!same) atomicMin(P.key, FQ_KEY(P.p0 + k, P.rank));
  }
}

/* ------------------------------------------------------------------------------------------------ multi-GPU: names to owners */
/* Per-owner counts and bytes.  A warp votes owner by owner (ballot + one warp reduction each) instead of 64 shared-memory atomics
 * on a handful of addresses; lane 0 keeps the warp's totals, one shared atomic per warp and owner at the end. */
__global__ void __launch_bounds__(256)
fq_names_count_kernel(const FqName* __restrict__ names, uint32_t nrec, uint32_t world, unsigned long long* out) {
  __shared__ unsigned long long sc[2 * FQ_SHARD_MAX_SRC];
  if (threadIdx.x < 2 * world) sc[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned long long cnt_acc = 0, byte_acc = 0; /* lane w (and w + 32) holds owner w's (w + 32's) totals of this warp */
  unsigned long long cnt_acc2 = 0, byte_acc2 = 0;
  for (uint32_t base = blockIdx.x * blockDim.x; base < nrec; base += gridDim.x * blockDim.x) {
    const uint32_t k = base + threadIdx.x;
    unsigned long long h = FQ_HASH_SKIP; uint32_t len4 = 0;
    if (k < nrec) { const FqName nm = names[k]; h = nm.hash; len4 = (nm.len + 3u) & ~3u; }
    const bool valid = h != FQ_HASH_SKIP;
    const uint32_t o = valid ? fq_owner_of(h, world) : 0xFFFFFFFFu;
    for (uint32_t w = 0; w < world; w++) {
      const uint32_t m = __ballot_sync(FULL, o == w);
      if (!m) continue;
      const uint32_t b = __reduce_add_sync(FULL, o == w ? len4 : 0u);
      if ((uint32_t)lane == (w & 31u)) { if (w < 32) { cnt_acc += __popc(m); byte_acc += b; } else { cnt_acc2 += __popc(m); byte_acc2 += b; } }
    }
  }
  if ((uint32_t)lane < world) { if (cnt_acc) { atomicAdd(&sc[2 * lane], cnt_acc); atomicAdd(&sc[2 * lane + 1], byte_acc); } }
  if ((uint32_t)lane + 32 < world) { if (cnt_acc2) { atomicAdd(&sc[2 * (lane + 32)], cnt_acc2); atomicAdd(&sc[2 * (lane + 32) + 1], byte_acc2); } }
  __syncthreads();
  if (threadIdx.x < 2 * world && sc[threadIdx.x]) atomicAdd(out + threadIdx.x, sc[threadIdx.x]);
}

/* Packing: positions inside an owner's stream come from warp votes (rank among the lanes of the same owner) and one warp prefix of
 * the padded lengths per owner; the warp reserves its share with one shared atomic per owner, the block its share with one
 * global atomic per owner. */
__global__ void __launch_bounds__(256)
fq_names_pack_kernel(const FqName* __restrict__ names, const uint8_t* __restrict__ data, uint32_t nrec, unsigned long long g0, uint32_t world,
                     FqPackedName* meta, uint8_t* blob, const unsigned long long* __restrict__ base, unsigned long long* cursor) {
  __shared__ unsigned long long s_cnt[2 * FQ_SHARD_MAX_SRC], s_base[2 * FQ_SHARD_MAX_SRC];
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t b0 = blockIdx.x * blockDim.x; b0 < nrec; b0 += gridDim.x * blockDim.x) {
    if (threadIdx.x < 2 * world) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t k = b0 + threadIdx.x;
    FqName nm; nm.hash = FQ_HASH_SKIP; nm.off = 0; nm.len = 0;
    if (k < nrec) nm = names[k];
    const bool valid = nm.hash != FQ_HASH_SKIP;
    const uint32_t o = valid ? fq_owner_of(nm.hash, world) : 0xFFFFFFFFu;
    const uint32_t len4 = (nm.len + 3u) & ~3u;
    unsigned long long my_m = 0, my_b = 0;
    for (uint32_t w = 0; w < world; w++) {
      const uint32_t m = __ballot_sync(FULL, o == w);
      if (!m) continue;
      uint32_t incl = o == w ? len4 : 0u; /* prefix of the padded lengths over the lanes of this owner */
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
      const uint32_t tot = __shfl_sync(FULL, incl, 31);
      unsigned long long wm = 0, wb = 0;
      if (lane == 0) { wm = atomicAdd(&s_cnt[2 * w], (unsigned long long)__popc(m)); wb = atomicAdd(&s_cnt[2 * w + 1], (unsigned long long)tot); }
      wm = __shfl_sync(FULL, wm, 0); wb = __shfl_sync(FULL, wb, 0);
      if (o == w) { my_m = wm + __popc(m & lt); my_b = wb + incl - len4; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * world) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(cursor + threadIdx.x, s_cnt[threadIdx.x]) : 0ull;
    __syncthreads();
    if (valid) {
      unsigned long long bo = s_base[2 * o + 1] + my_b; /* multiple of 4: names are padded to whole words in the blob */
      FqPackedName pn; pn.hash = nm.hash; pn.record = g0 + k; pn.off = (uint32_t)bo; pn.len = nm.len;
      meta[base[2 * o] + s_base[2 * o] + my_m] = pn;
      if (blob) { /* without a blob only the tuples travel: equal hashes are then verified in a second round (rarely needed) */
        uint32_t* dst = (uint32_t*)(blob + base[2 * o + 1] + bo);
        for (uint32_t i = 0; i < nm.len; i += 4) dst[i >> 2] = fq_low_bytes(fq_ldu32(data, nm.off + i), nm.len - i);
      }
    }
    __syncthreads();
  }
}

/* ------------------------------------------------------------------------------------------------ pipelined routing (tuples only)
 * The names of a finished chunk travel while the next chunk's clean-data pass runs.  Every owner has a region of fixed capacity
 * (local send buffer for an all-to-all, or the owner's own memory mapped over NVLink): no counting pass, no size exchange.  The
 * kernels are capped at 32 registers: one block fits on an SM beside the four resident blocks of the pass. */
struct SlotPackParams {
  const FqName* names; const uint8_t* arena; uint32_t nrec; unsigned long long g0; uint32_t world; unsigned long long cap; uint32_t units;
  unsigned long long* cursors; FqRegionPtrs R;
};
#define FQ_PACK_ILP 2
__global__ void __launch_bounds__(256, 8)
fq_names_pack_slots_kernel(const SlotPackParams P) {
  /* One block per SM is all the room there is beside the pass: the kernel lives on memory-level parallelism instead of occupancy.
   * Every thread has FQ_PACK_ILP names in flight; the rank of a name among the warp's names of the same owner comes from one
   * match.any (no loop over the owners), the warp's share of an owner's region from one shared atomic by the group's first lane.
   * A slot is its 16-byte header and `units` 16-byte units of the name's bytes, copied from the arena (aligned, zero padded). */
  __shared__ unsigned int s_cnt[FQ_SHARD_MAX_SRC];
  __shared__ unsigned long long s_base[FQ_SHARD_MAX_SRC];
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const size_t slot_bytes = fq_route_slot_bytes(P.units);
  for (uint32_t b0 = blockIdx.x * (256u * FQ_PACK_ILP); b0 < P.nrec; b0 += gridDim.x * (256u * FQ_PACK_ILP)) {
    if (threadIdx.x < P.world) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    FqName nm[FQ_PACK_ILP]; uint32_t my[FQ_PACK_ILP];
#pragma unroll
    for (int j = 0; j < FQ_PACK_ILP; j++) {
      const uint32_t k = b0 + j * 256u + threadIdx.x;
      nm[j].hash = FQ_HASH_SKIP; nm[j].len = 0; nm[j].off = 0;
      if (k < P.nrec) nm[j] = P.names[k];
    }
#pragma unroll
    for (int j = 0; j < FQ_PACK_ILP; j++) {
      const bool valid = nm[j].hash != FQ_HASH_SKIP;
      const uint32_t o = valid ? fq_owner_of(nm[j].hash, P.world) : 0xFFFFFFFFu;
      const uint32_t m = __match_any_sync(FULL, o);
      const int leader = __ffs(m) - 1;
      uint32_t wm = 0;
      if (valid && lane == leader) wm = atomicAdd(&s_cnt[o], (unsigned int)__popc(m));
      wm = __shfl_sync(FULL, wm, leader);
      my[j] = wm + __popc(m & lt);
    }
    __syncthreads();
    if (threadIdx.x < P.world) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(P.cursors + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < FQ_PACK_ILP; j++) {
      if (nm[j].hash == FQ_HASH_SKIP) continue;
      const uint32_t o = fq_owner_of(nm[j].hash, P.world);
      const unsigned long long pos = s_base[o] + my[j];
      if (pos >= P.cap) continue; /* counted, not stored: the header's count tells the owner */
      uint4* slot = (uint4*)(P.R.region[o] + 16 + fq_route_counts_bytes(1) + pos * slot_bytes); /* one writer: one stretch of `cap` slots */
      const unsigned long long rec = P.g0 + b0 + j * 256u + threadIdx.x, rl = (rec << 12) | nm[j].len;
      slot[0] = make_uint4((uint32_t)nm[j].hash, (uint32_t)(nm[j].hash >> 32), (uint32_t)rl, (uint32_t)(rl >> 32));
      for (uint32_t u = 0; u < P.units; u++)
        slot[1 + u] = 16u * u < nm[j].len ? *(const uint4*)(P.arena + nm[j].off + 16u * u) : make_uint4(0u, 0u, 0u, 0u);
      if (P.units && nm[j].len > 16u * P.units) atomicOr(P.cursors + FQ_SHARD_MAX_SRC + o, (unsigned long long)FQ_ROUTE_NAME_TOO_LONG);
    }
    __syncthreads();
  }
}
struct SlotHeaderParams { const unsigned long long* cursors; uint32_t world; unsigned long long cap; FqRegionPtrs R; };
__global__ void fq_slots_header_kernel(const SlotHeaderParams P) {
  if (threadIdx.x < P.world) {
    FqRegionHdr h; h.nblocks = 1; h.stride = (uint32_t)P.cap; h.flags = (uint32_t)P.cursors[FQ_SHARD_MAX_SRC + threadIdx.x]; h.pad = 0;
    *(FqRegionHdr*)P.R.region[threadIdx.x] = h;
    const unsigned long long c = P.cursors[threadIdx.x];
    *(uint32_t*)(P.R.region[threadIdx.x] + 16) = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c; /* a count above the stride: the surplus was dropped */
  }
}
/* the name a table slot points at (idx1 = address >> 4 of a route slot in this rank's own memory) against the name of route slot `sl` */
__device__ __forceinline__ bool route_same_name(unsigned long long idx1, const FqRouteSlot* sl) {
  const FqRouteSlot* other = (const FqRouteSlot*)(idx1 << 4);
  const uint32_t len = (uint32_t)(sl->rec_len & 0xFFFull);
  if ((uint32_t)(other->rec_len & 0xFFFull) != len) return false;
  const uint4* a = (const uint4*)(other + 1); const uint4* b = (const uint4*)(sl + 1);
  for (uint32_t u = 0; 16u * u < len; u++) { const uint4 x = a[u], y = b[u]; if ((x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w)) return false; }
  return true;
}
struct SlotInsertParams {
  const uint8_t* regions; uint32_t n_src; size_t region_bytes; uint32_t nblocks; unsigned long long stride; uint32_t units;
  FqSlot* slots; unsigned long long mask; unsigned long long* counters;
  const unsigned long long* flags; unsigned long long expect; /* flags[s] >= expect: source s has delivered its region (NULL: the host knows) */
  long long patience; /* clock cycles a source may stay silent */
};
/* The sources announce their regions with a flag word written behind the region's bytes (same stream of copies, so the bytes are
 * there when the flag is): the owner's kernel waits for the flags on the device — no host takes part in a routing round.  One
 * thread per source polls (system-scope acquire: the writers are the copy engines of other GPUs); a source that stays silent for
 * seconds is given up (counters[2]: the caller repeats the job).  False: do not touch the regions. */
__device__ __forceinline__ bool route_wait_sources(const SlotInsertParams& P) {
  __shared__ int s_late;
  if (!P.flags) return true;
  if (threadIdx.x == 0) s_late = 0;
  __syncthreads();
  if (threadIdx.x < P.n_src) {
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(P.flags + threadIdx.x) : "memory");
      if (v >= P.expect) break;
      if (clock64() - t0 > P.patience) { s_late = 1; atomicExch(P.counters + 2, 1ull); break; } /* about ten seconds */
      __nanosleep(200);
    }
  }
  __syncthreads();
  return s_late == 0;
}
/* The wait is a kernel of its own, one warp in front of the owner's kernel on the same stream: a grid that filled the device while
 * it polled would keep out whatever the flags still depend on (a copy the driver performs with a kernel, a collective's kernel, the
 * pass of a slower peer's host) — measured with four ranks as a deadlock that only the patience above resolved. */
__global__ void __launch_bounds__(32, 1)
fq_route_wait_kernel(const SlotInsertParams P) { route_wait_sources(P); }
/* the owner's kernel itself only looks: the flags are there unless the wait gave up */
__device__ __forceinline__ bool route_sources_there(const SlotInsertParams& P) {
  __shared__ int s_missing;
  if (!P.flags) return true;
  if (threadIdx.x == 0) s_missing = 0;
  __syncthreads();
  if (threadIdx.x < P.n_src) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(P.flags + threadIdx.x) : "memory");
    if (v < P.expect) s_missing = 1;
  }
  __syncthreads();
  return s_missing == 0;
}
/* The slots of the regions, one warp per stretch (source s, writer b): the stretch's count is read once, the lanes take its slots
 * 32 at a time.  f(slot) for every slot; stretches that are not what was planned raise counters[2]. */
template <class F>
__device__ __forceinline__ void route_each_slot(const SlotInsertParams& P, F f) {
  const size_t slot_bytes = fq_route_slot_bytes(P.units);
  const int lane = threadIdx.x & 31;
  const unsigned long long segs = (P.stride + 1023ull) >> 10; /* a stretch is taken in pieces of 1024 slots (a dense region is one long stretch) */
  const unsigned long long nunits = (unsigned long long)P.n_src * P.nblocks * segs;
  const unsigned long long warp0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long un = warp0; un < nunits; un += nwarps) {
    const unsigned long long st = un / segs, sg = un - st * segs;
    const unsigned long long src = st / P.nblocks, b = st - src * P.nblocks;
    const uint8_t* reg = P.regions + src * P.region_bytes;
    const FqRegionHdr h = *(const FqRegionHdr*)reg;
    if (h.nblocks == 0) continue; /* this source had nothing for us in this round */
    if (b == 0 && sg == 0 && lane == 0 && (h.nblocks > P.nblocks || h.stride != P.stride || h.flags)) atomicExch(P.counters + 2, 1ull); /* not what was planned, or a name longer than its slot */
    if (b >= h.nblocks) continue;
    uint32_t cnt = ((const uint32_t*)(reg + 16))[b];
    if (cnt > P.stride) { if (lane == 0 && sg == 0) atomicExch(P.counters + 2, 1ull); cnt = (uint32_t)P.stride; } /* the writer had more names for this owner than its stretch holds */
    const uint8_t* base = reg + 16 + fq_route_counts_bytes(h.nblocks) + b * P.stride * slot_bytes;
    const uint32_t k1 = (uint32_t)min((unsigned long long)cnt, (sg + 1) << 10);
    for (uint32_t k = (uint32_t)(sg << 10) + lane; k < k1; k += 32) f((const FqRouteSlot*)(base + (size_t)k * slot_bytes));
  }
}
__global__ void __launch_bounds__(256, 8)
fq_shard_insert_slots_kernel(const SlotInsertParams P) {
  /* One probe in flight per thread.  Measured on B200 beside the pass: 2 or 4 independent first probes per thread are slower (the
   * atomics, not their latency, are the limit), and so are short blocks on a low-priority stream (they crowd the start of the
   * next pass, whose blocks must all be resident). */
  unsigned long long inserted = 0, equal = 0;
  if (!route_sources_there(P)) return;
  route_each_slot(P, [&](const FqRouteSlot* sl) {
    const unsigned long long hash = sl->hash, me = (unsigned long long)(uintptr_t)sl >> 4;
    unsigned long long i = hash & P.mask, probes = 0;
    for (;; i = (i + 1) & P.mask) {
      if (++probes > P.mask) { atomicExch(P.counters + 2, 1ull); break; }
      unsigned long long cur, cur_idx;
      if (slot_claim128(P.slots + i, hash, me, &cur, &cur_idx)) { inserted++; break; }
      if (cur != hash) continue;
      if (P.units && !route_same_name(cur_idx, sl)) continue; /* another name with this hash: walk on */
      equal++; /* the name is there already (units = 0: or its hash is, which tuples alone cannot judge): the exact path reports it */
      break;
    }
  });
  __syncwarp();
  inserted = warp_sum64(inserted); equal = warp_sum64(equal);
  if ((threadIdx.x & 31) == 0) { if (inserted) atomicAdd(P.counters + 1, inserted); if (equal) atomicAdd(P.counters + 0, equal); }
}
__global__ void __launch_bounds__(256, 8)
fq_shard_claim_slots_kernel(const SlotInsertParams P) {
  unsigned long long claimed = 0, unpaired = 0;
  if (!route_sources_there(P)) return;
  route_each_slot(P, [&](const FqRouteSlot* sl) {
    const unsigned long long hash = sl->hash, rec = sl->rec_len >> 12;
    unsigned long long i = hash & P.mask, probes = 0;
    for (;; i = (i + 1) & P.mask) {
      if (++probes > P.mask + 1) { unpaired++; break; }
      const unsigned long long cur = P.slots[i].hash;
      if (cur == FQ_HASH_EMPTY) { unpaired++; break; }
      if (cur != hash) continue;
      if (!route_same_name(P.slots[i].idx1, sl)) continue;
      const unsigned long long old = atomicMin(&P.slots[i].claim2, rec);
      if (old == FQ_IDX_NONE) claimed++; else unpaired++; /* the entry was deleted by an earlier mate: the reference reports the later one */
      break;
    }
  });
  __syncwarp();
  claimed = warp_sum64(claimed); unpaired = warp_sum64(unpaired);
  if ((threadIdx.x & 31) == 0) { if (claimed) atomicAdd(P.counters + 8, claimed); if (unpaired) atomicAdd(P.counters + 9, unpaired); }
}

struct ShardParams { FqShardArgs a; };
__device__ __forceinline__ const uint8_t* shard_name(const FqShardArgs& a, unsigned long long pos, uint32_t* len) {
  uint32_t src = 0;
  while (src + 1 < a.n_src && pos >= a.meta_start[src + 1]) src++;
  const FqPackedName pn = a.meta[pos];
  *len = pn.len;
  return a.blob + a.blob_start[src] + pn.off;
}
__global__ void __launch_bounds__(256)
fq_shard_insert_kernel(const ShardParams P) {
  const FqShardArgs& a = P.a;
  const unsigned long long posmask = (1ull << FQ_SHARD_POS_BITS) - 1;
  for (unsigned long long m = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; m < a.n; m += (unsigned long long)gridDim.x * blockDim.x) {
    const FqPackedName pn = a.meta[m];
    const unsigned long long mine = (pn.record << FQ_SHARD_POS_BITS) | m; /* orders by record index first */
    unsigned long long i = pn.hash & a.mask, probes = 0;
    for (;; i = (i + 1) & a.mask) {
      if (++probes > a.mask) { atomicExch(a.counters + 2, 1ull); break; }
      FqSlot* s = a.slots + i;
      unsigned long long cur, cur_idx;
      if (slot_claim128(s, pn.hash, mine, &cur, &cur_idx)) break; /* first arrival of this name: one atomic on one sector */
      if (cur != pn.hash) continue;
      if (!a.blob) { atomicAdd(a.counters + 0, 1ull); break; } /* tuples only: an equal hash cannot be judged here */
      uint32_t ol, ml; const uint8_t* on = shard_name(a, cur_idx & posmask, &ol); const uint8_t* mn = shard_name(a, m, &ml);
      if (!(ol == ml && names_equal(on, mn, ml))) { atomicAdd(a.counters + 4, 1ull); continue; } /* another name with this hash: next slot */
      unsigned long long old = atomicMin(&s->idx1, mine);
      unsigned long long og = old >> FQ_SHARD_POS_BITS, later = og > pn.record ? og : pn.record;
      atomicMin(a.dup_key, FQ_KEY(later, FQ_R_NAME));
      break;
    }
  }
}
struct ClaimParams { FqShardArgs a; FqShardArgs ins; unsigned long long sb; };
__global__ void __launch_bounds__(256)
fq_shard_claim_kernel(const ClaimParams P) {
  const FqShardArgs& a = P.a;
  const unsigned long long posmask = (1ull << FQ_SHARD_POS_BITS) - 1;
  unsigned long long claimed = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long b0 = (unsigned long long)blockIdx.x * blockDim.x; b0 < a.n; b0 += stride) {
    unsigned long long m = b0 + threadIdx.x;
    if (m >= a.n) continue;
    const FqPackedName pn = a.meta[m];
    unsigned long long i = pn.hash & a.mask, probes = 0, unpaired = FQ_IDX_NONE;
    for (;; i = (i + 1) & a.mask) {
      if (++probes > a.mask + 1) { unpaired = pn.record; break; }
      const FqSlot* s = a.slots + i;
      unsigned long long cur = s->hash;
      if (cur == FQ_HASH_EMPTY) { unpaired = pn.record; break; }
      if (cur != pn.hash) continue;
      uint32_t ol, ml; const uint8_t* on = shard_name(P.ins, s->idx1 & posmask, &ol); const uint8_t* mn = shard_name(a, m, &ml);
      if (!(ol == ml && names_equal(on, mn, ml))) continue; /* another name with this hash: the lookup walks on */
      unsigned long long old = atomicMin(&a.slots[i].claim2, pn.record);
      if (old == FQ_IDX_NONE) claimed++;
      else unpaired = old > pn.record ? old : pn.record;
      break;
    }
    if (unpaired != FQ_IDX_NONE) atomicMin(a.dup_key, FQ_KEY(P.sb + unpaired, FQ_R_NAME));
  }
  __syncwarp();
  claimed = warp_sum64(claimed);
  if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(a.counters + 1, claimed);
}
__global__ void fq_shard_find_kernel(const FqPackedName* __restrict__ meta, unsigned long long n, unsigned long long record, unsigned long long* out_pos) {
  for (unsigned long long m = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (unsigned long long)gridDim.x * blockDim.x)
    if (meta[m].record == record) atomicMin(out_pos, m);
}

/* ------------------------------------------------------------------------------------------------ the name arena
 * What the reference keeps of a record is a copy of its name (new_indexentry, src/fastq.c:590-611).  Here the names of a segment
 * are copied out of their chunk into one block of 16-byte units (zero padded: equal names are equal unit by unit), after which
 * nothing refers to the chunk's bytes any more and the engine may release them.  This kernel serves the chunks of the per-record
 * kernels; the clean-data pass writes its names into the arena itself, from shared memory. */
__global__ void __launch_bounds__(256)
fq_names_measure_kernel(const FqName* __restrict__ names, uint32_t nrec, unsigned long long* out_units) {
  unsigned long long u = 0;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nrec; k += gridDim.x * blockDim.x) {
    const FqName nm = names[k];
    if (nm.hash != FQ_HASH_SKIP) u += (nm.len + 15u) >> 4;
  }
  u = warp_sum64(u);
  if ((threadIdx.x & 31) == 0 && u) atomicAdd(out_units, u);
}
__global__ void __launch_bounds__(256)
fq_names_gather_kernel(FqName* names, const uint8_t* __restrict__ data, uint32_t nrec, uint8_t* arena, unsigned long long* cursor_units) {
  __shared__ uint32_t s_w[8];
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t b0 = blockIdx.x * 256u; b0 < nrec; b0 += gridDim.x * 256u) {
    const uint32_t k = b0 + threadIdx.x;
    FqName nm; nm.hash = FQ_HASH_SKIP; nm.off = 0; nm.len = 0;
    if (k < nrec) nm = names[k];
    const uint32_t units = nm.hash != FQ_HASH_SKIP ? (nm.len + 15u) >> 4 : 0u;
    uint32_t incl = units;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += a; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t excl = incl - units, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const uint32_t x = s_w[w]; total += x; if (w < warp) excl += x; }
    if (threadIdx.x == 0) s_base = total ? atomicAdd(cursor_units, (unsigned long long)total) : 0ull; /* the block's names lie together, in any order among the blocks */
    __syncthreads();
    if (units) {
      const unsigned long long at = (s_base + excl) * 16ull;
      uint4* dst = (uint4*)(arena + at);
      for (uint32_t u = 0; u < units; u++) {
        const uint32_t o = nm.off + 16u * u, left = nm.len - 16u * u; /* left >= 1 */
        uint4 v;
        v.x = fq_low_bytes(fq_ldu32(data, o), left);
        v.y = left > 4 ? fq_low_bytes(fq_ldu32(data, o + 4), left - 4) : 0u;
        v.z = left > 8 ? fq_low_bytes(fq_ldu32(data, o + 8), left - 8) : 0u;
        v.w = left > 12 ? fq_low_bytes(fq_ldu32(data, o + 12), left - 12) : 0u;
        dst[u] = v;
      }
      names[k].off = (uint32_t)at;
    }
    __syncthreads(); /* s_w / s_base are rewritten by the next round */
  }
}
/* open statistics -> main statistics (fq_device.h: stats_fold); two launches: the bins first (they read the open ranges), then the scalars */
struct FoldParams { FqStats* main_[2]; FqStats* open_[2]; unsigned long long* hist[2]; unsigned long long* hopen[2]; };
__global__ void fq_stats_fold_hist_kernel(const FoldParams P) {
  uint32_t lo = min(P.open_[0]->min_rl, P.open_[1]->min_rl), hi = max(P.open_[0]->max_rl, P.open_[1]->max_rl);
  if (hi >= FQ_MAX_READ_LENGTH) hi = FQ_MAX_READ_LENGTH - 1;
  if (lo > hi) return;
  for (uint32_t l = lo + blockIdx.x * blockDim.x + threadIdx.x; l <= hi; l += gridDim.x * blockDim.x)
    for (int f = 0; f < 2; f++) { const unsigned long long v = P.hopen[f][l]; if (v) { P.hist[f][l] += v; P.hopen[f][l] = 0; } }
}
__global__ void fq_stats_fold_scalars_kernel(const FoldParams P) {
  if (threadIdx.x || blockIdx.x) return;
  for (int f = 0; f < 2; f++) {
    FqStats* m = P.main_[f]; FqStats* o = P.open_[f];
    m->num_rds += o->num_rds; m->mem_sum += o->mem_sum; m->n_names += o->n_names;
    m->min_rl = min(m->min_rl, o->min_rl); m->max_rl = max(m->max_rl, o->max_rl);
    m->min_q = min(m->min_q, o->min_q); m->max_q = max(m->max_q, o->max_q);
    o->num_rds = 0; o->mem_sum = 0; o->n_names = 0; o->min_rl = 0xFFFFFFFFu; o->max_rl = 0; o->min_q = 255u; o->max_q = 0;
  }
}

/* fastq_filter_n (src/fastq_filter_n.c:77-86): one warp per sequence line, the lanes stride over its bytes */
__global__ void __launch_bounds__(256)
fq_count_n_kernel(const uint8_t* __restrict__ data, const FqLine* __restrict__ lines, uint32_t n, uint32_t* out2) {
  const int lane = threadIdx.x & 31;
  for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += (gridDim.x * blockDim.x) >> 5) {
    const FqLine L = lines[k];
    uint32_t cnt = 0, stop = L.len, nul = L.len; /* first LF or NUL / first NUL */
    for (uint32_t base = 0; base < L.len; base += 32) {
      const uint32_t i = base + lane;
      const uint32_t c = i < L.len ? data[L.off + i] : 1u;
      const uint32_t mstop = __ballot_sync(FULL, c == '\n' || c == 0), mnul = __ballot_sync(FULL, c == 0);
      if (mnul && nul == L.len) nul = base + (uint32_t)__ffs(mnul) - 1;
      const uint32_t upto = mstop ? (uint32_t)__ffs(mstop) - 1 : 32u;
      cnt += __popc(__ballot_sync(FULL, (c == 'N' || c == 'n') && (uint32_t)lane < upto));
      if (mstop) { stop = base + upto; break; }
    }
    if (nul == L.len && stop < L.len) { /* the line goes on behind its LF only when gzgets cut it: look for a NUL there too (strlen) */
      for (uint32_t base = stop & ~31u; base < L.len && nul == L.len; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t mnul = __ballot_sync(FULL, i < L.len && i >= stop && data[L.off + i] == 0);
        if (mnul) nul = base + (uint32_t)__ffs(mnul) - 1;
      }
    }
    if (lane == 0) { out2[2 * k] = cnt; out2[2 * k + 1] = nul; }
  }
}
/* fastq_filterpair (src/fastq_filterpair.c): names of header lines, and where the index holds them */
__global__ void __launch_bounds__(256)
fq_header_names_kernel(const uint8_t* __restrict__ data, const FqLine* __restrict__ lines, uint32_t n, int fmt, int is_pe, uint32_t seed, FqName* out) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = fq_header_name(data, lines[k].off, lines[k].len, fmt, is_pe, seed);
}
__global__ void __launch_bounds__(256)
fq_names_lookup_kernel(const TableParams P, unsigned long long* out_idx) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < P.nrec; k += gridDim.x * blockDim.x) {
    const FqName nm = P.names[k];
    unsigned long long found = FQ_IDX_NONE;
    if (nm.hash != FQ_HASH_SKIP) {
      unsigned long long i = nm.hash & P.mask, probes = 0;
      for (;; i = (i + 1) & P.mask) {
        if (++probes > P.mask + 1) break;
        const FqSlot* s = P.slots + i;
        const unsigned long long cur = s->hash;
        if (cur == FQ_HASH_EMPTY) break;
        if (cur != nm.hash) continue;
        uint32_t ol; const uint8_t* on = dir_name(P.dir1, P.ndir1, s->idx1, &ol);
        if (!(ol == nm.len && fq_bytes_equal(on, P.data + nm.off, nm.len))) continue; /* another name with this hash: the lookup walks on (src/hash.c:38-45) */
        found = s->idx1;
        break;
      }
    }
    out_idx[k] = found;
  }
}
/* fastq_trim_poly_at (src/fastq_trim_poly_at.c:77-115): one thread per sequence line (the scans stop at the first other base) */
__global__ void __launch_bounds__(256)
fq_poly_at_kernel(const uint8_t* __restrict__ data, const FqLine* __restrict__ lines, uint32_t n, uint32_t* out3) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const FqLine L = lines[k];
    uint32_t rl, a, b;
    fq_poly_at(data + L.off, L.len, &rl, &a, &b);
    out3[3 * k] = rl; out3[3 * k + 1] = a; out3[3 * k + 2] = b;
  }
}
struct WordsParams { uint32_t* dst; uint32_t n; uint32_t w[32]; };
__global__ void fq_set_words_kernel(const WordsParams P) { if (threadIdx.x < P.n) P.dst[threadIdx.x] = P.w[threadIdx.x]; }

/* ------------------------------------------------------------------------------------------------ the device */
class FqCudaDevice : public FqDevice {
 public:
  explicit FqCudaDevice(int ordinal) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw std::runtime_error("no CUDA device available: libfastq_gpu has no CPU fallback");
    if (ordinal < 0 || ordinal >= count) throw std::runtime_error("fqg_create: CUDA ordinal out of range");
    dev_ = ordinal;
    FQ_CUDA_CHECK(cudaSetDevice(dev_));
    cudaDeviceProp prop; FQ_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev_));
    sms_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : kSMs;
    FQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    FQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st2_, cudaStreamNonBlocking));
    FQ_CUDA_CHECK(cudaEventCreateWithFlags(&evx_, cudaEventDisableTiming));
    FQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev_pre_, cudaEventDisableTiming));
    FQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev_side_, cudaEventDisableTiming));
    FQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev_order_, cudaEventDisableTiming));
    FQ_CUDA_CHECK(cudaEventCreate(&ev0_)); FQ_CUDA_CHECK(cudaEventCreate(&ev1_));
    cudaMemPool_t pool; FQ_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev_));
    unsigned long long thr = ~0ull; /* keep freed blocks in the pool: allocations repeat every chunk */
    FQ_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    max_tiles_ = (uint32_t)((1ull << 31) / SCAN_TILE) + 2;
    FQ_CUDA_CHECK(cudaMalloc(&tile_state_, (size_t)max_tiles_ * sizeof(unsigned long long) + 64));
    ticket_ = (uint32_t*)(tile_state_ + max_tiles_);
  }
  ~FqCudaDevice() override {
    cudaSetDevice(dev_);
    cudaStreamSynchronize(st_); cudaStreamSynchronize(st2_);
    for (auto& p : pending_) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto& d : deferred_) cudaEventDestroy(d.ready);
    for (auto e : free_ev_) cudaEventDestroy(e);
    cudaFree(tile_state_); if (stage_) cudaFree(stage_); if (small_pinned_) cudaFreeHost(small_pinned_);
    cudaEventDestroy(ev0_); cudaEventDestroy(ev1_); cudaEventDestroy(evx_); cudaEventDestroy(ev_pre_); cudaEventDestroy(ev_side_); cudaEventDestroy(ev_order_);
    for (int k = 0; k < 7; k++) if (st_lane_[k]) { cudaStreamDestroy(st_lane_[k]); cudaEventDestroy(ev_lane_[k]); }
    cudaStreamDestroy(st_); cudaStreamDestroy(st2_);
  }
  const char* name() const override { return "cuda"; }
  void* alloc(size_t n) override {
    void* p = nullptr;
    FQ_CUDA_CHECK(cudaSetDevice(dev_));
    FQ_CUDA_CHECK(cudaMallocAsync(&p, n ? n : 1, st_));
    return p;
  }
  void release(void* p) override { if (p) cudaFreeAsync(p, st_); }
  void* host_alloc(size_t n) override { void* p = nullptr; FQ_CUDA_CHECK(cudaSetDevice(dev_)); FQ_CUDA_CHECK(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault)); return p; }
  void host_release(void* p) override { if (p) cudaFreeHost(p); }
  void upload(void* d, const void* s, size_t n) override { if (n) FQ_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st_)); }
  void download(void* d, const void* s, size_t n) override {
    if (n && n <= sizeof small_) { /* result words: through page-locked memory (no staging inside the driver) */
      if (!small_pinned_) FQ_CUDA_CHECK(cudaHostAlloc(&small_pinned_, sizeof small_, cudaHostAllocDefault));
      FQ_CUDA_CHECK(cudaMemcpyAsync(small_pinned_, s, n, cudaMemcpyDeviceToHost, st_));
      FQ_CUDA_CHECK(cudaStreamSynchronize(st_));
      memcpy(d, small_pinned_, n);
      return;
    }
    if (n) FQ_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st_));
    FQ_CUDA_CHECK(cudaStreamSynchronize(st_));
  }
  void set_words(uint32_t* d, const uint32_t* words, uint32_t nwords) override { /* (a launch instead of a copy out of pageable memory) */
    WordsParams P; P.dst = d; P.n = nwords > 32 ? 32 : nwords;
    for (uint32_t i = 0; i < P.n; i++) P.w[i] = words[i];
    fq_set_words_kernel<<<1, 32, 0, st_>>>(P);
    launched();
  }
  void copy(void* d, const void* s, size_t n) override { if (n) FQ_CUDA_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st_)); }
  void fill(void* d, int b, size_t n) override { if (n) FQ_CUDA_CHECK(cudaMemsetAsync(d, b, n, st_)); }
  void fill_index(void* d, int b, size_t n) override { if (n) FQ_CUDA_CHECK(cudaMemsetAsync(d, b, n, st2_)); }
  void sync() override { flush_deferred(); FQ_CUDA_CHECK(cudaStreamSynchronize(st_)); FQ_CUDA_CHECK(cudaStreamSynchronize(st2_)); }
  void sync_main() override { FQ_CUDA_CHECK(cudaStreamSynchronize(st_)); }
  void timer_start() override { FQ_CUDA_CHECK(cudaEventRecord(ev0_, st_)); }
  double timer_stop_ms() override {
    FQ_CUDA_CHECK(cudaEventRecord(ev1_, st_)); FQ_CUDA_CHECK(cudaEventSynchronize(ev1_));
    float ms = 0; FQ_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    return ms;
  }
  unsigned long long launches() const override { return n_launch_; }
  bool kernel_stat(int which, double* ms, uint64_t* launches, uint64_t* bytes, uint64_t* items) override {
    if (which < 0 || which >= FQG_K_COUNT) return false;
    collect();
    *ms = kst_[which].ms; *launches = kst_[which].launches; *bytes = kst_[which].bytes; *items = kst_[which].items;
    return true;
  }
  void kernel_stats_reset() override { collect(); for (auto& k : kst_) k = KStat(); }

  void scan_lines(const uint8_t* data, uint32_t n, int virtual_end, uint32_t* line_end, uint32_t cap, uint32_t* out2, uint32_t lead = 0) override {
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 0) { FQ_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(uint32_t), st_)); return; }
    if (ntiles > max_tiles_) throw std::runtime_error("scan_lines: chunk larger than 2 GiB (n=" + std::to_string(n) + ")");
    FQ_CUDA_CHECK(cudaMemsetAsync(tile_state_, 0, (size_t)ntiles * sizeof(unsigned long long), st_));
    FQ_CUDA_CHECK(cudaMemsetAsync(ticket_, 0, sizeof(uint32_t), st_));
    tic(FQG_K_SCAN, n, ntiles);
    fq_scan_kernel<<<ntiles, SCAN_THREADS, 0, st_>>>(data, n, virtual_end, line_end, cap, tile_state_, ticket_, ntiles, out2, lead);
    toc();
    launched();
  }
  void count_lines(const uint8_t* data, uint32_t n, unsigned long long* out) override {
    if (!n) return;
    tic(FQG_K_SCAN, n, 1);
    fq_count_kernel<<<sms_ * 8, 256, 0, st_>>>(data, n, out);
    toc(); launched();
  }
  void find_overlong(const uint32_t* line_end, uint32_t q, uint32_t j0, uint32_t nlines, uint32_t n, int tail_from_n, uint32_t* out) override {
    uint32_t total = nlines + (tail_from_n ? 1u : 0u);
    if (!total) return;
    int grid = (int)std::min<uint32_t>((total + 255) / 256, (uint32_t)sms_ * 8);
    fq_overlong_kernel<<<grid, 256, 0, st_>>>(line_end, q, j0, nlines, n, tail_from_n, out);
    launched();
  }
  void split_serial(const uint8_t* data, uint32_t n, uint32_t q, int is_eof, FqLine* lines4, uint32_t* out3) override {
    fq_split_serial_kernel<<<1, 32, 0, st_>>>(data, n, q, is_eof, lines4, out3);
    launched();
  }
  void sniff(const uint8_t* data, FqLine hdr1, FqLine seq, int32_t* out2) override {
    fq_sniff_kernel<<<1, 32, 0, st_>>>(data, hdr1, seq, out2);
    launched();
  }
  void records(const FqRecordsArgs& a) override {
    if (!a.nrec) return;
    RecParams P;
    P.data = a.data; P.line_end = a.line_end; P.lines = a.lines; P.q = a.q; P.j0 = a.j0; P.nrec = a.nrec; P.g0 = a.g0;
    P.step_base = a.step_base; P.cx = a.cx; P.stats = a.stats; P.stats_range = a.stats_range; P.hist = a.hist; P.key = a.key; P.names = a.names;
    int grid = (int)std::min<uint32_t>((a.nrec + REC_THREADS - 1) / REC_THREADS, (uint32_t)sms_ * 16);
    tic(FQG_K_RECORDS, a.span_bytes, a.nrec);
    fq_records_kernel<<<grid, REC_THREADS, 0, st_>>>(P);
    toc();
    launched();
    flush_deferred();
  }
  bool tile_pass(const FqTileArgs& a) override {
    if (!a.n || ((uintptr_t)a.data & 15u)) return false; /* its bulk copies need a 16-byte aligned chunk */
    uint32_t ntiles = (a.n + TILE_BYTES - 1) / TILE_BYTES;
    if (ntiles > max_tiles_) return false;
    if (tile_blocks_ == 0) {
      FQ_CUDA_CHECK(cudaFuncSetAttribute(fq_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
      int per_sm = 0;
      FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fq_tile_kernel, TILE_THREADS, TILE_SMEM));
      if (per_sm < 1) return false;
      tile_blocks_ = per_sm * sms_; /* every CTA resident: the look-back may wait on any earlier tile */
    }
    TileParams P;
    P.data = a.data; P.n = a.n; P.virtual_end = a.virtual_end; P.line_end = a.line_end; P.cap = a.cap;
    P.tile_state = tile_state_; P.ticket = ticket_; P.ntiles = ntiles; P.out = a.out5;
    P.j0 = a.j0; P.max_rec = a.max_rec; P.g0 = a.g0; P.step_base = a.step_base; P.cx = a.cx;
    P.stats = a.stats; P.stats_range = a.stats_range; P.hist = a.hist; P.key = a.key; P.names = a.names; P.names_cap = a.names_cap;
    FQ_CUDA_CHECK(cudaMemsetAsync(tile_state_, 0, (size_t)ntiles * sizeof(unsigned long long), st_));
    FQ_CUDA_CHECK(cudaMemsetAsync(ticket_, 0, sizeof(uint32_t), st_));
    int grid = (int)std::min<uint32_t>(ntiles, (uint32_t)tile_blocks_);
    tic(FQG_K_TILE, a.n, ntiles);
    fq_tile_kernel<<<grid, TILE_THREADS, TILE_SMEM, st_>>>(P);
    toc(); launched();
    flush_deferred();
    return true;
  }
  bool lanes_pass(const FqTileArgs& a, bool* self_judged) override {
    *self_judged = false;
    if (!a.n) return false;
    if ((uintptr_t)a.data & 15u) return false; /* bulk copies need a 16-byte aligned source (the engine aligns what it is fed) */
    /* short lines (they must end within the 1 KiB margin): one thread per line; anything longer: chunk-parallel */
    const bool lines_mode = a.hint_line_len > 0 && a.hint_line_len <= 512 && !getenv("FQG_NO_LINES");
    const uint32_t tile_bytes = lines_mode ? LS_TILE : LN_TILE;
    uint32_t ntiles = (a.n + tile_bytes - 1) / tile_bytes;
    if (ntiles > max_tiles_) return false;
    ensure_lanes_blocks();
    if (lanes_blocks_ < 1) return false;
    LanesParams P;
    P.data = a.data; P.lead = a.lead; P.n = a.n; P.virtual_end = a.virtual_end; P.line_end = a.line_end; P.cap = a.cap;
    P.tile_state = tile_state_; P.ticket = ticket_; P.ntiles = ntiles; P.out = a.out5;
    P.j0 = a.j0; P.cx = a.cx; P.names = a.names; P.names_cap = a.names_cap;
    if (!stage_) FQ_CUDA_CHECK(cudaMalloc(&stage_, sizeof(LanesStage)));
    P.stage = stage_; P.arena = lines_mode ? a.arena : nullptr; P.arena_units = a.arena_units;
    P.route_world = lines_mode ? a.route_world : 0; P.route_stride = a.route_stride; P.route_units = a.route_units;
    for (int o = 0; o < FQ_ROUTE_MAX_WORLD; o++) P.route_region[o] = a.route_region[o];
    if (a.route_world && !lines_mode) return false; /* only the per-line mode routes names itself */
    if (side_marked_) { FQ_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_side_, 0)); side_marked_ = false; } /* copies out of the regions this pass overwrites */
    for (int k = 0; k < 7; k++) if (lane_marked_ & (1u << k)) FQ_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_lane_[k], 0));
    lane_marked_ = 0;
    for (uint32_t o = 0; o < P.route_world; o++) FQ_CUDA_CHECK(cudaMemsetAsync(P.route_region[o], 0, sizeof(FqRegionHdr), st_)); /* (flags are or-ed in; nblocks = 0 until the pass is through) */
    { const char* e = getenv("FQG_LANES_TUNE"); P.tune = e ? (uint32_t)atoi(e) : 0u; }
    FQ_CUDA_CHECK(cudaEventRecord(ev_pre_, st_));
    if (lines_mode) FQ_CUDA_CHECK(cudaMemsetAsync(stage_, 0, sizeof(LanesStage), st_));
    FQ_CUDA_CHECK(cudaMemsetAsync(tile_state_, 0, (size_t)ntiles * sizeof(unsigned long long), st_));
    FQ_CUDA_CHECK(cudaMemsetAsync(ticket_, 0, sizeof(uint32_t), st_));
    int grid = (int)std::min<uint32_t>(ntiles, (uint32_t)lanes_blocks_);
    tic(FQG_K_LANES, a.n, ntiles);
    if (lines_mode) fq_lanes_kernel<true><<<grid, LN_THREADS, LN_SMEM, st_>>>(P);
    else fq_lanes_kernel<false><<<grid, LN_THREADS, LN_SMEM, st_>>>(P);
    toc(); launched();
    if (lines_mode) { /* the pass judged the records itself: one block turns what it staged into statistics if the chunk is clean */
      LanesPostParams R;
      R.out = a.out5; R.line_end = a.line_end; R.stage = stage_; R.j0 = a.j0; R.cx = a.cx; R.stats = a.stats; R.stats_range = a.stats_range; R.hist = a.hist; R.names_cap = a.names ? a.names_cap : 0;
      fq_lanes_post_kernel<<<1, 1024, 0, st_>>>(R);
      launched();
      *self_judged = true;
    } else lanes_records(a, false);
    flush_deferred(); /* the previous chunk's inserts run beside this pass */
    return true;
  }
  void lanes_records(const FqTileArgs& a, bool undo) {
    LanesRecParams R;
    R.line_end = a.line_end; R.out = a.out5; R.j0 = a.j0; R.lead = a.lead; R.names = a.names; R.cx = a.cx; R.stats = a.stats; R.hist = a.hist; R.undo = undo ? 1 : 0;
    uint32_t max_rec = a.cap / 4 + 1;
    int grid = (int)std::min<uint32_t>((max_rec + 255) / 256, (uint32_t)sms_ * 8);
    tic(FQG_K_RECORDS, 0, max_rec);
    fq_lanes_records_kernel<<<grid, 256, 0, st_>>>(R);
    toc(); launched();
  }
  void lanes_commit(const FqTileArgs& a, bool undo) override {
    if (undo) { lanes_records(a, true); return; }
    fq_lanes_commit_kernel<<<1, 32, 0, st_>>>(a.out5, a.stats_range);
    launched();
  }
  void ensure_lanes_blocks() {
    if (lanes_blocks_ != 0) return;
    FQ_CUDA_CHECK(cudaFuncSetAttribute(fq_lanes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    FQ_CUDA_CHECK(cudaFuncSetAttribute(fq_lanes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    int per_sm = 0, per_sm2 = 0;
    FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fq_lanes_kernel<false>, LN_THREADS, LN_SMEM));
    FQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, fq_lanes_kernel<true>, LN_THREADS, LN_SMEM));
    per_sm = std::min(per_sm, per_sm2);
    if (per_sm < 1) { lanes_blocks_ = -1; return; }
    if (const char* e = getenv("FQG_LANES_CTAS")) { int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; } /* tuning hook: leave room for the index kernel */
    lanes_blocks_ = per_sm * sms_; /* every CTA resident: the look-back may wait on any earlier tile */
  }
  static TableParams table_params(const FqTableArgs& a) {
    TableParams P;
    P.names = a.names; P.data = a.data; P.nrec = a.nrec; P.g0 = a.g0; P.step_base = a.step_base; P.slots = a.slots; P.mask = a.mask;
    P.dir1 = a.dir1; P.ndir1 = a.ndir1; P.key = a.key; P.counters = a.counters;
    return P;
  }
  /* The insert kernel of a chunk is not launched at once: a kernel that starts on an idle GPU takes every register file, and the
   * clean-data pass of the NEXT chunk (persistent CTAs, 56 registers) could not start beside it.  The launch waits until that pass
   * has been launched (it leaves room for exactly one insert block per SM) or until somebody needs the results. */
  void index_insert(const FqTableArgs& a) override { defer(a, false); }
  void mate_claim(const FqTableArgs& a) override { /* not deferred: a claim compares name bytes, one block per SM beside the pass is too slow for it (measured 40 ms against 9) */
    if (!a.nrec) return;
    flush_deferred();
    int grid = (int)std::min<uint32_t>((a.nrec + 255) / 256, (uint32_t)sms_ * 8);
    after_main();
    tic(FQG_K_MATE, 0, a.nrec, st2_);
    fq_mate_claim_kernel<<<grid, 256, 0, st2_>>>(table_params(a));
    toc(st2_);
    launched();
  }
  void defer(const FqTableArgs& a, bool claim) {
    if (!a.nrec) return;
    Deferred d; d.tp = table_params(a); d.claim = claim;
    if (free_ev_.empty()) { FQ_CUDA_CHECK(cudaEventCreate(&d.ready)); } else { d.ready = free_ev_.back(); free_ev_.pop_back(); }
    FQ_CUDA_CHECK(cudaEventRecord(d.ready, st_)); /* names, table fills and directories queued so far */
    deferred_.push_back(d);
  }
  void flush_deferred() { /* in order: the claims of the mate loop come after every insert of the index loop */
    for (auto& d : deferred_) {
      int grid = (int)std::min<uint32_t>((d.tp.nrec + 255) / 256, (uint32_t)sms_ * 8);
      FQ_CUDA_CHECK(cudaStreamWaitEvent(st2_, d.ready, 0));
      tic(d.claim ? FQG_K_MATE : FQG_K_INDEX, 0, d.tp.nrec, st2_);
      if (d.claim) fq_mate_claim_kernel<<<grid, 256, 0, st2_>>>(d.tp);
      else fq_index_insert_kernel<<<grid, 256, 0, st2_>>>(d.tp);
      toc(st2_);
      launched();
      free_ev_.push_back(d.ready); /* reused only after a later record: the wait above has been queued already */
    }
    deferred_.clear();
  }
  void pair_compare(const FqPairArgs& a) override {
    if (!a.npairs) return;
    PairParams P;
    P.a = a.a; P.da = a.da; P.stride_a = a.stride_a; P.b = a.b; P.db = a.db; P.stride_b = a.stride_b;
    P.npairs = a.npairs; P.p0 = a.p0; P.rank = a.rank; P.key = a.key;
    int grid = (int)std::min<uint32_t>((a.npairs + 255) / 256, (uint32_t)sms_ * 8);
    tic(FQG_K_PAIR, 0, a.npairs);
    fq_pair_compare_kernel<<<grid, 256, 0, st_>>>(P);
    toc();
    launched();
  }
  void names_count(const FqName* names, uint32_t nrec, uint32_t world, unsigned long long* out) override {
    if (!nrec) return;
    int grid = (int)std::min<uint32_t>((nrec + 255) / 256, (uint32_t)sms_ * 8);
    tic(FQG_K_OTHER, 0, nrec);
    fq_names_count_kernel<<<grid, 256, 0, st_>>>(names, nrec, world, out);
    toc(); launched();
  }
  void names_pack(const FqName* names, const uint8_t* data, uint32_t nrec, uint64_t g0, uint32_t world, FqPackedName* meta, uint8_t* blob,
                  const unsigned long long* base, unsigned long long* cursor) override {
    if (!nrec) return;
    int grid = (int)std::min<uint32_t>((nrec + 255) / 256, (uint32_t)sms_ * 8);
    tic(FQG_K_OTHER, 0, nrec);
    fq_names_pack_kernel<<<grid, 256, 0, st_>>>(names, data, nrec, g0, world, meta, blob, base, cursor);
    toc(); launched();
  }
  void shard_insert(const FqShardArgs& a) override {
    if (!a.n) return;
    ShardParams P; P.a = a;
    int grid = (int)std::min<unsigned long long>((a.n + 255) / 256, (unsigned long long)sms_ * 8);
    after_main();
    tic(FQG_K_INDEX, 0, a.n, st2_);
    fq_shard_insert_kernel<<<grid, 256, 0, st2_>>>(P);
    toc(st2_); launched();
  }
  void shard_claim(const FqShardArgs& a, const FqShardArgs& ins, unsigned long long sb) override {
    if (!a.n) return;
    ClaimParams P; P.a = a; P.ins = ins; P.sb = sb;
    int grid = (int)std::min<unsigned long long>((a.n + 255) / 256, (unsigned long long)sms_ * 8);
    after_main();
    tic(FQG_K_MATE, 0, a.n, st2_);
    fq_shard_claim_kernel<<<grid, 256, 0, st2_>>>(P);
    toc(st2_); launched();
  }
  void route_begin(unsigned long long* cursors, uint32_t world, bool beside) override {
    /* beside: the main stream is busy with a clean-data pass launched after the names were complete; wait only for what was
     * queued before that pass (ev_pre_).  Otherwise: everything queued on the main stream so far. */
    (void)world;
    if (beside) { FQ_CUDA_CHECK(cudaStreamWaitEvent(st2_, ev_pre_, 0)); } else after_main();
    FQ_CUDA_CHECK(cudaMemsetAsync(cursors, 0, 2 * FQ_SHARD_MAX_SRC * sizeof(unsigned long long), st2_));
  }
  void names_pack_slots(const FqName* names, const uint8_t* arena, uint32_t nrec, uint64_t g0, uint32_t world, const FqRegionPtrs& R, uint64_t cap,
                        uint32_t units, unsigned long long* cursors) override {
    if (!nrec) return;
    SlotPackParams P; P.names = names; P.arena = arena; P.nrec = nrec; P.g0 = g0; P.world = world; P.cap = cap; P.units = units; P.cursors = cursors; P.R = R;
    int grid = (int)std::min<uint32_t>((nrec + 256u * FQ_PACK_ILP - 1) / (256u * FQ_PACK_ILP), (uint32_t)sms_ * 8);
    tic(FQG_K_OTHER, 0, nrec, st2_);
    fq_names_pack_slots_kernel<<<grid, 256, 0, st2_>>>(P);
    toc(st2_); launched();
  }
  void route_end(const unsigned long long* cursors, uint32_t world, const FqRegionPtrs& R, uint64_t cap) override {
    SlotHeaderParams P; P.cursors = cursors; P.world = world; P.cap = cap; P.R = R;
    fq_slots_header_kernel<<<1, FQ_SHARD_MAX_SRC, 0, st2_>>>(P);
    launched();
    FQ_CUDA_CHECK(cudaStreamSynchronize(st2_));
  }
  void shard_insert_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units, FqSlot* slots,
                          unsigned long long mask, unsigned long long* counters, bool beside, const unsigned long long* flags, unsigned long long expect) override {
    slots_kernel(regions, n_src, region_bytes, nblocks, stride, units, slots, mask, counters, beside, false, flags, expect);
  }
  void shard_claim_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units, FqSlot* slots,
                         unsigned long long mask, unsigned long long* counters, bool beside, const unsigned long long* flags, unsigned long long expect) override {
    slots_kernel(regions, n_src, region_bytes, nblocks, stride, units, slots, mask, counters, beside, true, flags, expect);
  }
  void slots_kernel(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units, FqSlot* slots,
                    unsigned long long mask, unsigned long long* counters, bool beside, bool claim, const unsigned long long* flags, unsigned long long expect) {
    if (!n_src || !nblocks || !stride) return;
    SlotInsertParams P; P.regions = regions; P.n_src = n_src; P.region_bytes = region_bytes; P.nblocks = nblocks; P.stride = stride; P.units = units;
    P.slots = slots; P.mask = mask; P.counters = counters; P.flags = flags; P.expect = expect;
    static const long long patience_ms = getenv("FQG_FLAG_PATIENCE_MS") ? atoll(getenv("FQG_FLAG_PATIENCE_MS")) : 10000; /* (tests, debugging) */
    P.patience = patience_ms * 2000000ll;
    unsigned long long total = (unsigned long long)n_src * nblocks * stride;
    /* one warp per stretch when there are many stretches; a dense region (one writer) is one stretch: its warp would work alone */
    const unsigned long long pieces = (unsigned long long)n_src * nblocks * ((stride + 1023) >> 10);
    int grid = (int)std::min<unsigned long long>(std::max<unsigned long long>((pieces + 7) / 8, 1), (unsigned long long)sms_ * (beside ? 1 : 8));
    after_main();
    if (flags) { fq_route_wait_kernel<<<1, 32, 0, st2_>>>(P); launched(); }
    tic(claim ? FQG_K_MATE : FQG_K_INDEX, 0, total, st2_);
    if (claim) fq_shard_claim_slots_kernel<<<grid, 256, 0, st2_>>>(P); else fq_shard_insert_slots_kernel<<<grid, 256, 0, st2_>>>(P);
    toc(st2_); launched();
  }
  uint32_t lanes_max_blocks() override { ensure_lanes_blocks(); return (uint32_t)lanes_blocks_; }
  void order_after(bool my_side, FqDevice& earlier, bool their_side) override {
    auto* e = dynamic_cast<FqCudaDevice*>(&earlier);
    if (!e || e->dev_ != dev_) throw std::runtime_error("fqg_order_after: two contexts of one CUDA device");
    FQ_CUDA_CHECK(cudaEventRecord(ev_order_, their_side ? e->st2_ : e->st_));
    FQ_CUDA_CHECK(cudaStreamWaitEvent(my_side ? st2_ : st_, ev_order_, 0));
  }
  void side_mark() override {
    FQ_CUDA_CHECK(cudaEventRecord(ev_side_, st2_)); side_marked_ = true;
    /* (the side stream joins the lanes as well: what it does next — the pack kernel of a file's last round writes into the staging
     * regions — comes behind every copy out of them, without the host waiting for anything) */
    for (int k = 0; k < 7; k++) if (lane_dirty_ & (1u << k)) {
      FQ_CUDA_CHECK(cudaEventRecord(ev_lane_[k], st_lane_[k])); lane_marked_ |= 1u << k;
      FQ_CUDA_CHECK(cudaStreamWaitEvent(st2_, ev_lane_[k], 0));
    }
    lane_dirty_ = 0;
  }
  void side_copy_lane(int lane, void* dst, const void* src, size_t n) override {
    if (lane <= 0) return side_copy(dst, src, n);
    const int k = lane - 1;
    if (!st_lane_[k]) {
      FQ_CUDA_CHECK(cudaStreamCreateWithFlags(&st_lane_[k], cudaStreamNonBlocking));
      FQ_CUDA_CHECK(cudaEventCreateWithFlags(&ev_lane_[k], cudaEventDisableTiming));
    }
    if (n) FQ_CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, st_lane_[k]));
    lane_dirty_ |= 1u << k; lane_unsynced_ |= 1u << k;
  }
  void side_copy(void* dst, const void* src, size_t n) override { if (n) FQ_CUDA_CHECK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, st2_)); }
  void side_sync() override {
    FQ_CUDA_CHECK(cudaStreamSynchronize(st2_));
    for (int k = 0; k < 7; k++) if (lane_unsynced_ & (1u << k)) FQ_CUDA_CHECK(cudaStreamSynchronize(st_lane_[k]));
    lane_unsynced_ = 0;
  }
  void* ipc_alloc(size_t n, uint8_t handle[64]) override {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr; FQ_CUDA_CHECK(cudaSetDevice(dev_)); FQ_CUDA_CHECK(cudaMalloc(&p, n ? n : 1));
    cudaIpcMemHandle_t h; FQ_CUDA_CHECK(cudaIpcGetMemHandle(&h, p)); memcpy(handle, &h, 64);
    return p;
  }
  void* ipc_open(const uint8_t handle[64]) override {
    cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
    void* p = nullptr; FQ_CUDA_CHECK(cudaSetDevice(dev_)); FQ_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return p;
  }
  void ipc_close(void* p) override { if (p) FQ_CUDA_CHECK(cudaIpcCloseMemHandle(p)); }
  void ipc_free(void* p) override { if (p) { FQ_CUDA_CHECK(cudaDeviceSynchronize()); FQ_CUDA_CHECK(cudaFree(p)); } }
  void shard_find(const FqPackedName* meta, unsigned long long n, unsigned long long record, unsigned long long* out_pos) override {
    if (!n) return;
    int grid = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)sms_ * 8);
    fq_shard_find_kernel<<<grid, 256, 0, st_>>>(meta, n, record, out_pos);
    launched();
  }
  void names_measure(const FqName* names, uint32_t nrec, unsigned long long* out_units) override {
    if (!nrec) return;
    int grid = (int)std::min<uint32_t>((nrec + 255) / 256, (uint32_t)sms_ * 8);
    fq_names_measure_kernel<<<grid, 256, 0, st_>>>(names, nrec, out_units);
    launched();
  }
  void names_gather(FqName* names, const uint8_t* data, uint32_t nrec, uint8_t* arena, unsigned long long* cursor_units) override {
    if (!nrec) return;
    int grid = (int)std::min<uint32_t>((nrec + 255) / 256, (uint32_t)sms_ * 8);
    tic(FQG_K_OTHER, 0, nrec);
    fq_names_gather_kernel<<<grid, 256, 0, st_>>>(names, data, nrec, arena, cursor_units);
    toc(); launched();
  }
  void stats_fold(FqStats* const main2[2], FqStats* const open2[2], unsigned long long* const hist2[2], unsigned long long* const hist_open2[2]) override {
    FoldParams P;
    for (int f = 0; f < 2; f++) { P.main_[f] = main2[f]; P.open_[f] = open2[f]; P.hist[f] = hist2[f]; P.hopen[f] = hist_open2[f]; }
    fq_stats_fold_hist_kernel<<<sms_, 256, 0, st_>>>(P); launched();
    fq_stats_fold_scalars_kernel<<<1, 32, 0, st_>>>(P); launched();
  }
  void count_n(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out2) override {
    if (!n) return;
    int grid = (int)std::min<uint32_t>((n + 7) / 8, (uint32_t)sms_ * 8);
    fq_count_n_kernel<<<grid, 256, 0, st_>>>(data, seq_lines, n, out2);
    launched();
  }
  void header_names(const uint8_t* data, const FqLine* hdr_lines, uint32_t n, int fmt, int is_pe, uint32_t seed, FqName* out) override {
    if (!n) return;
    int grid = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)sms_ * 8);
    fq_header_names_kernel<<<grid, 256, 0, st_>>>(data, hdr_lines, n, fmt, is_pe, seed, out);
    launched();
  }
  void names_lookup(const FqTableArgs& a, unsigned long long* out_idx) override {
    if (!a.nrec) return;
    TableParams P; P.names = a.names; P.data = a.data; P.nrec = a.nrec; P.g0 = a.g0; P.step_base = a.step_base;
    P.slots = a.slots; P.mask = a.mask; P.dir1 = a.dir1; P.ndir1 = a.ndir1; P.key = a.key; P.counters = a.counters;
    FQ_CUDA_CHECK(cudaStreamSynchronize(st2_)); /* the inserts run on the index stream */
    int grid = (int)std::min<uint32_t>((a.nrec + 255) / 256, (uint32_t)sms_ * 8);
    fq_names_lookup_kernel<<<grid, 256, 0, st_>>>(P, out_idx);
    launched();
  }
  void poly_at(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out3) override {
    if (!n) return;
    int grid = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)sms_ * 8);
    fq_poly_at_kernel<<<grid, 256, 0, st_>>>(data, seq_lines, n, out3);
    launched();
  }
  void explain(const uint8_t* data, const FqLine* L, const FqRecCtx& cx, FqRecOut* out_dev) override {
    fq_explain_kernel<<<1, 32, 0, st_>>>(data, L[0], L[1], L[2], L[3], cx, out_dev);
    launched();
  }

 private:
  void launched() { n_launch_++; FQ_CUDA_CHECK(cudaGetLastError()); }
  /* CUDA-event stopwatch around one launch; elapsed times are read back lazily (collect) */
  struct KStat { double ms = 0; uint64_t launches = 0, bytes = 0, items = 0; };
  struct Pending { int cls; cudaEvent_t a, b; bool side; };
  /* The index kernels (random-access, latency-bound) run on a second stream so that they overlap the next chunk's
   * streaming pass; they start after everything queued on the main stream so far (names, table fills, directories). */
  void after_main() {
    FQ_CUDA_CHECK(cudaEventRecord(evx_, st_));
    FQ_CUDA_CHECK(cudaStreamWaitEvent(st2_, evx_, 0));
  }
  void tic(int cls, uint64_t bytes, uint64_t items, cudaStream_t st = nullptr) {
    if (!st) st = st_;
    Pending p; p.cls = cls; p.side = st == st2_;
    if (free_ev_.size() >= 2) { p.a = free_ev_.back(); free_ev_.pop_back(); p.b = free_ev_.back(); free_ev_.pop_back(); }
    else { FQ_CUDA_CHECK(cudaEventCreate(&p.a)); FQ_CUDA_CHECK(cudaEventCreate(&p.b)); }
    kst_[cls].launches++; kst_[cls].bytes += bytes; kst_[cls].items += items;
    if (!g_origin && getenv("FQG_TIMELINE")) { FQ_CUDA_CHECK(cudaEventCreate(&g_origin)); FQ_CUDA_CHECK(cudaEventRecord(g_origin, st)); }
    FQ_CUDA_CHECK(cudaEventRecord(p.a, st));
    pending_.push_back(p);
  }
  void toc(cudaStream_t st = nullptr) { FQ_CUDA_CHECK(cudaEventRecord(pending_.back().b, st ? st : st_)); }
  void collect() {
    flush_deferred();
    if (pending_.empty()) return;
    FQ_CUDA_CHECK(cudaStreamSynchronize(st_)); FQ_CUDA_CHECK(cudaStreamSynchronize(st2_));
    /* FQG_TIMELINE=<file>: where every timed launch of this batch sat, in ms after the batch's first one (a development aid) */
    const char* tl = getenv("FQG_TIMELINE");
    FILE* tf = tl && *tl && g_origin ? fopen((std::string(tl) + "." + std::to_string(dev_)).c_str(), "a") : nullptr; /* one file per device */
    if (tf) fprintf(tf, "# device %d object %p, %zu launches\n", dev_, (void*)this, pending_.size());
    for (auto& p : pending_) {
      float ms = 0; FQ_CUDA_CHECK(cudaEventElapsedTime(&ms, p.a, p.b));
      kst_[p.cls].ms += ms;
      if (tf) {
        float t0 = 0; FQ_CUDA_CHECK(cudaEventElapsedTime(&t0, g_origin, p.a));
        fprintf(tf, "%d %s %.3f %.3f\n", p.cls, p.side ? "side" : "main", t0, t0 + ms);
      }
      free_ev_.push_back(p.a); free_ev_.push_back(p.b);
    }
    pending_.clear();
    if (tf) fclose(tf);
  }
  static cudaEvent_t g_origin; /* FQG_TIMELINE: the first timed launch of the process */
  struct Deferred { TableParams tp; cudaEvent_t ready; bool claim; };
  std::vector<Deferred> deferred_;
  KStat kst_[FQG_K_COUNT];
  std::vector<Pending> pending_;
  std::vector<cudaEvent_t> free_ev_;
  int dev_ = 0, sms_ = kSMs, tile_blocks_ = 0, lanes_blocks_ = 0;
  cudaStream_t st_ = nullptr, st2_ = nullptr;
  cudaEvent_t evx_ = nullptr;
  cudaEvent_t ev_side_ = nullptr; bool side_marked_ = false;
  cudaEvent_t ev_order_ = nullptr;
  cudaStream_t st_lane_[7] = {}; cudaEvent_t ev_lane_[7] = {}; /* more copy streams (side_copy_lane) */
  uint32_t lane_dirty_ = 0, lane_marked_ = 0, lane_unsynced_ = 0;
  cudaEvent_t ev_pre_ = nullptr; /* main stream just before the latest clean-data pass: what the side stream waits for when it works beside that pass */
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  unsigned long long* tile_state_ = nullptr; uint32_t* ticket_ = nullptr; uint32_t max_tiles_ = 0;
  LanesStage* stage_ = nullptr;
  uint8_t small_[4096]; void* small_pinned_ = nullptr;
  unsigned long long n_launch_ = 0;
};

cudaEvent_t FqCudaDevice::g_origin = nullptr;

}  // namespace

FqDevice* fq_make_cuda_device(int ordinal) { return new FqCudaDevice(ordinal); }
FqDevice* fq_default_device(int ordinal) { return fq_make_cuda_device(ordinal); }

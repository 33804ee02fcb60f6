/*
 * tests/sim/fq_sim.cpp — TEST INFRASTRUCTURE ONLY.  A sequential stand-in for FqCudaDevice so that the host
 * engine (chunking, bridging, event ordering, report, rendering) and the shared per-record semantics
 * (fq_record.h) can be exercised without a GPU.  Built into tests/sim/libfastq_sim.so by tests/sim/Makefile;
 * never linked into libfastq_gpu.so and never used by bench.py or the -m gpu tests.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <map>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include "../../fastq_utils_b200/csrc/fq_device.h"
#include "../../fastq_utils_b200/csrc/fq_record.h"

namespace {

class FqSimDevice : public FqDevice {
 public:
  const char* name() const override { return "sim"; }
  void* alloc(size_t n) override { void* p = calloc(1, n + 64); if (!p) throw std::bad_alloc(); return p; }
  void release(void* p) override { free(p); }
  void upload(void* d, const void* s, size_t n) override { memcpy(d, s, n); }
  void download(void* d, const void* s, size_t n) override { memcpy(d, s, n); }
  void copy(void* d, const void* s, size_t n) override { memmove(d, s, n); }
  void fill(void* d, int b, size_t n) override { memset(d, b, n); }
  void sync() override {}
  void timer_start() override {}
  double timer_stop_ms() override { return 0.0; }
  unsigned long long launches() const override { return n_launch_; }
  bool kernel_stat(int, double* ms, uint64_t* l, uint64_t* b, uint64_t* i) override { *ms = 0; *l = *b = *i = 0; return true; }
  void kernel_stats_reset() override {}

  void scan_lines(const uint8_t* data, uint32_t n, int virtual_end, uint32_t* line_end, uint32_t cap, uint32_t* out2, uint32_t lead = 0) override {
    n_launch_++;
    uint32_t k = 0;
    for (uint32_t i = lead; i < n; i++)
      if (data[i] == '\n') { if (k < cap) line_end[k] = i + 1; k++; }
    if (virtual_end && n > 0 && data[n - 1] != '\n') { if (k < cap) line_end[k] = n; k++; }
    out2[0] = k; out2[1] = k > cap;
  }
  void count_lines(const uint8_t* data, uint32_t n, unsigned long long* out) override {
    n_launch_++;
    for (uint32_t i = 0; i < n; i++) *out += data[i] == '\n';
  }
  void find_overlong(const uint32_t* line_end, uint32_t q, uint32_t j0, uint32_t nlines, uint32_t n, int tail_from_n, uint32_t* out) override {
    n_launch_++;
    for (uint32_t i = 0; i <= nlines; i++) {
      uint32_t start = i == 0 ? q : line_end[j0 + i - 1];
      uint32_t end;
      if (i < nlines) end = line_end[j0 + i];
      else { if (!tail_from_n) break; end = n; }
      uint32_t lim = (i & 1) == 0 ? FQ_MAX_LABEL_LENGTH : FQ_MAX_READ_LENGTH;
      if (end - start >= lim) { if (j0 + i < *out) *out = j0 + i; break; }
    }
  }
  void split_serial(const uint8_t* data, uint32_t n, uint32_t q, int is_eof, FqLine* lines4, uint32_t* out3) override {
    n_launch_++;
    uint32_t p = q, got = 0, lfs = 0;
    for (int i = 0; i < 4; i++) {
      uint32_t maxb = ((i & 1) == 0 ? FQ_MAX_LABEL_LENGTH : FQ_MAX_READ_LENGTH) - 1, k = 0;
      bool lf = false;
      while (k < maxb && p + k < n) { k++; if (data[p + k - 1] == '\n') { lf = true; break; } }
      bool complete = lf || k == maxb || (is_eof && k > 0);
      if (!complete) break;
      lines4[i].off = p; lines4[i].len = k; p += k; got++; lfs += lf;
    }
    for (uint32_t i = got; i < 4; i++) { lines4[i].off = p; lines4[i].len = 0; }
    out3[0] = p; out3[1] = got; out3[2] = lfs;
  }
  void sniff(const uint8_t* data, FqLine hdr1, FqLine seq, int32_t* out2) override {
    n_launch_++;
    uint32_t cl0 = fq_cstrlen(data, hdr1.off, hdr1.len), cl1 = fq_cstrlen(data, seq.off, seq.len);
    out2[0] = fq_sniff_format(data + hdr1.off + 1, cl0 >= 1 ? cl0 - 1 : 0);
    out2[1] = fq_sniff_colorspace(data + seq.off, cl1);
  }
  static void lines_of(const FqRecordsArgs& a, uint32_t k, FqLine L[4]) {
    if (a.lines) { memcpy(L, a.lines + 4 * k, 4 * sizeof(FqLine)); return; }
    uint32_t j = a.j0 + 4 * k;
    uint32_t s = k == 0 ? a.q : a.line_end[j - 1];
    for (int i = 0; i < 4; i++) { uint32_t e = a.line_end[j + i]; L[i].off = s; L[i].len = e - s; s = e; }
  }
  void records(const FqRecordsArgs& a) override {
    n_launch_++;
    for (uint32_t k = 0; k < a.nrec; k++) {
      FqLine L[4]; lines_of(a, k, L);
      FqRecOut o; uint64_t hsh; fq_check_record(a.data, L, a.cx, &o, &hsh);
      uint64_t g = a.g0 + k;
      uint64_t key = fq_record_key(a.cx.loop, g, a.step_base, o);
      if (key < *a.key) *a.key = key;
      bool named = fq_record_has_name(a.cx.loop, o);
      if (a.names) {
        a.names[k].off = o.name_off; a.names[k].len = o.name_len;
        a.names[k].hash = named ? hsh : FQ_HASH_SKIP;
      }
      if (a.cx.loop == FQ_LOOP_INDEX && named) { a.stats->n_names++; a.stats->mem_sum += o.mem_len; }
      /* statistics are only ever reported when every record was clean, so only clean records are counted */
      if (o.flags || (o.vrank != FQ_V_OK && a.cx.loop != FQ_LOOP_READER)) continue; /* the reader loop counts what it read */
      uint32_t w = a.cx.weight;
      a.stats->num_rds += w;
      if (o.read_len < a.stats_range->min_rl) a.stats_range->min_rl = o.read_len;
      if (o.read_len > a.stats_range->max_rl) a.stats_range->max_rl = o.read_len;
      a.hist[o.read_len] += w;
      if (o.qmin <= o.qmax) {
        if (o.qmin < a.stats_range->min_q) a.stats_range->min_q = o.qmin;
        if (o.qmax > a.stats_range->max_q) a.stats_range->max_q = o.qmax;
      }
    }
  }
  bool tile_pass(const FqTileArgs&) override { return false; }
  static const uint8_t* name_of(const FqDirEntry* dir, uint32_t nd, uint64_t g, uint32_t* len) {
    uint32_t lo = 0, hi = nd;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (dir[mid].g0 <= g) lo = mid; else hi = mid; }
    const FqName& nm = dir[lo].names[g - dir[lo].g0];
    *len = nm.len;
    return dir[lo].data + nm.off;
  }
  void index_insert(const FqTableArgs& a) override {
    n_launch_++;
    for (uint32_t k = 0; k < a.nrec; k++) {
      const FqName& nm = a.names[k];
      if (nm.hash == FQ_HASH_SKIP) continue;
      uint64_t g = a.g0 + k, i = nm.hash & a.mask, probes = 0;
      for (;; i = (i + 1) & a.mask) {
        if (++probes > a.mask) { a.counters[2] = 1; break; }
        FqSlot& s = a.slots[i];
        if (s.hash == FQ_HASH_EMPTY) { s.hash = nm.hash; s.idx1 = g; break; }
        if (s.hash != nm.hash) continue;
        uint64_t old = s.idx1;
        uint32_t ol; const uint8_t* on = name_of(a.dir1, a.ndir1, old, &ol);
        if (!(ol == nm.len && fq_bytes_equal(on, a.data + nm.off, nm.len))) { a.counters[4]++; continue; } /* another name with this hash: next slot */
        if (g < old) s.idx1 = g;
        uint64_t later = old > g ? old : g;
        uint64_t key = FQ_KEY(a.step_base + later, FQ_R_NAME);
        if (key < *a.key) *a.key = key;
        break;
      }
    }
  }
  void mate_claim(const FqTableArgs& a) override {
    n_launch_++;
    for (uint32_t k = 0; k < a.nrec; k++) {
      const FqName& nm = a.names[k];
      if (nm.hash == FQ_HASH_SKIP) continue;
      uint64_t g = a.g0 + k, i = nm.hash & a.mask, probes = 0;
      for (;; i = (i + 1) & a.mask) {
        uint64_t unpaired = FQ_KEY_NONE;
        if (++probes > a.mask + 1) { unpaired = g; }
        else {
          FqSlot& s = a.slots[i];
          if (s.hash == FQ_HASH_EMPTY) unpaired = g;
          else if (s.hash != nm.hash) continue;
          else {
            uint32_t ol; const uint8_t* on = name_of(a.dir1, a.ndir1, s.idx1, &ol);
            if (!(ol == nm.len && fq_bytes_equal(on, a.data + nm.off, nm.len))) continue; /* another name with this hash: the lookup walks on */
            uint64_t old = s.claim2; if (g < old) s.claim2 = g;
            if (old == FQ_IDX_NONE) a.counters[1]++;
            else unpaired = old > g ? old : g;
          }
        }
        if (unpaired != FQ_KEY_NONE) { uint64_t key = FQ_KEY(a.step_base + unpaired, FQ_R_NAME); if (key < *a.key) *a.key = key; }
        break;
      }
    }
  }
  void pair_compare(const FqPairArgs& a) override {
    n_launch_++;
    for (uint32_t k = 0; k < a.npairs; k++) {
      const FqName& x = a.a[(size_t)k * a.stride_a]; const FqName& y = a.b[(size_t)k * a.stride_b];
      if (x.hash == FQ_HASH_SKIP || y.hash == FQ_HASH_SKIP) continue;
      bool same = x.hash == y.hash && x.len == y.len && fq_bytes_equal(a.da + x.off, a.db + y.off, x.len);
      if (!same) { uint64_t key = FQ_KEY(a.p0 + k, a.rank); if (key < *a.key) *a.key = key; }
    }
  }
  void names_count(const FqName* names, uint32_t nrec, uint32_t world, unsigned long long* out) override {
    n_launch_++;
    for (uint32_t k = 0; k < nrec; k++) {
      if (names[k].hash == FQ_HASH_SKIP) continue;
      uint32_t o = fq_owner_of(names[k].hash, world);
      out[2 * o]++; out[2 * o + 1] += (names[k].len + 3u) & ~3u;
    }
  }
  void names_pack(const FqName* names, const uint8_t* data, uint32_t nrec, uint64_t g0, uint32_t world, FqPackedName* meta, uint8_t* blob,
                  const unsigned long long* base, unsigned long long* cursor) override {
    n_launch_++;
    for (uint32_t k = 0; k < nrec; k++) {
      const FqName& nm = names[k];
      if (nm.hash == FQ_HASH_SKIP) continue;
      uint32_t o = fq_owner_of(nm.hash, world);
      unsigned long long mi = cursor[2 * o]++, bo = cursor[2 * o + 1]; cursor[2 * o + 1] += (nm.len + 3u) & ~3u;
      FqPackedName& pn = meta[base[2 * o] + mi];
      pn.hash = nm.hash; pn.record = g0 + k; pn.off = (uint32_t)bo; pn.len = nm.len;
      if (blob) memcpy(blob + base[2 * o + 1] + bo, data + nm.off, nm.len);
    }
  }
  static const uint8_t* shard_name(const FqShardArgs& a, unsigned long long pos, uint32_t* len) {
    uint32_t src = 0; while (src + 1 < a.n_src && pos >= a.meta_start[src + 1]) src++;
    *len = a.meta[pos].len;
    return a.blob + a.blob_start[src] + a.meta[pos].off;
  }
  void shard_insert(const FqShardArgs& a) override {
    n_launch_++;
    const unsigned long long posmask = (1ull << FQ_SHARD_POS_BITS) - 1;
    for (unsigned long long m = 0; m < a.n; m++) {
      const FqPackedName& pn = a.meta[m];
      unsigned long long i = pn.hash & a.mask, probes = 0, mine = (pn.record << FQ_SHARD_POS_BITS) | m;
      for (;; i = (i + 1) & a.mask) {
        if (++probes > a.mask) { a.counters[2] = 1; break; }
        FqSlot& s = a.slots[i];
        if (s.hash == FQ_HASH_EMPTY) { s.hash = pn.hash; s.idx1 = mine; break; }
        if (s.hash != pn.hash) continue;
        unsigned long long old = s.idx1;
        if (!a.blob) { a.counters[0]++; break; } /* tuples only: an equal hash cannot be judged here */
        uint32_t ol, ml; const uint8_t* on = shard_name(a, old & posmask, &ol); const uint8_t* mn = shard_name(a, m, &ml);
        if (!(ol == ml && fq_bytes_equal(on, mn, ml))) { a.counters[4]++; continue; } /* another name with this hash: next slot */
        if (mine < old) s.idx1 = mine;
        unsigned long long later = std::max(old >> FQ_SHARD_POS_BITS, (unsigned long long)pn.record);
        unsigned long long key = FQ_KEY(later, FQ_R_NAME);
        if (key < *a.dup_key) *a.dup_key = key;
        break;
      }
    }
  }
  void shard_claim(const FqShardArgs& a, const FqShardArgs& ins, unsigned long long sb) override {
    n_launch_++;
    const unsigned long long posmask = (1ull << FQ_SHARD_POS_BITS) - 1;
    for (unsigned long long m = 0; m < a.n; m++) {
      const FqPackedName& pn = a.meta[m];
      unsigned long long i = pn.hash & a.mask, probes = 0, unpaired = FQ_IDX_NONE;
      for (;; i = (i + 1) & a.mask) {
        if (++probes > a.mask + 1) { unpaired = pn.record; break; }
        FqSlot& s = a.slots[i];
        if (s.hash == FQ_HASH_EMPTY) { unpaired = pn.record; break; }
        if (s.hash != pn.hash) continue;
        uint32_t ol, ml; const uint8_t* on = shard_name(ins, s.idx1 & posmask, &ol); const uint8_t* mn = shard_name(a, m, &ml);
        if (!(ol == ml && fq_bytes_equal(on, mn, ml))) continue; /* another name with this hash: the lookup walks on */
        unsigned long long old = s.claim2; if (pn.record < old) s.claim2 = pn.record;
        if (old == FQ_IDX_NONE) a.counters[1]++;
        else unpaired = old > pn.record ? old : pn.record;
        break;
      }
      if (unpaired != FQ_IDX_NONE) { unsigned long long key = FQ_KEY(sb + unpaired, FQ_R_NAME); if (key < *a.dup_key) *a.dup_key = key; }
    }
  }
  void route_begin(unsigned long long* cursors, uint32_t, bool) override { memset(cursors, 0, 2 * FQ_SHARD_MAX_SRC * sizeof(unsigned long long)); }
  void names_pack_slots(const FqName* names, const uint8_t* arena, uint32_t nrec, uint64_t g0, uint32_t world, const FqRegionPtrs& R, uint64_t cap,
                        uint32_t units, unsigned long long* cursors) override {
    n_launch_++;
    const size_t sb = fq_route_slot_bytes(units);
    for (uint32_t k = 0; k < nrec; k++) {
      const FqName& nm = names[k];
      if (nm.hash == FQ_HASH_SKIP) continue;
      uint32_t o = fq_owner_of(nm.hash, world);
      unsigned long long pos = cursors[o]++;
      if (pos >= cap) continue;
      uint8_t* slot = R.region[o] + 16 + fq_route_counts_bytes(1) + pos * sb;
      FqRouteSlot h; h.hash = nm.hash; h.rec_len = ((unsigned long long)(g0 + k) << 12) | nm.len;
      memcpy(slot, &h, 16);
      if (units) { memset(slot + 16, 0, (size_t)units * 16); memcpy(slot + 16, arena + nm.off, std::min<size_t>(nm.len, (size_t)units * 16)); }
      if (units && nm.len > 16u * units) cursors[FQ_SHARD_MAX_SRC + o] |= FQ_ROUTE_NAME_TOO_LONG;
    }
  }
  void route_end(const unsigned long long* cursors, uint32_t world, const FqRegionPtrs& R, uint64_t cap) override {
    for (uint32_t o = 0; o < world; o++) {
      FqRegionHdr h; h.nblocks = 1; h.stride = (uint32_t)cap; h.flags = (uint32_t)cursors[FQ_SHARD_MAX_SRC + o]; h.pad = 0;
      memcpy(R.region[o], &h, sizeof h);
      uint32_t c = cursors[o] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cursors[o];
      memcpy(R.region[o] + 16, &c, 4);
    }
  }
  static bool same_name(unsigned long long idx1, const FqRouteSlot* sl) {
    const FqRouteSlot* other = (const FqRouteSlot*)(uintptr_t)(idx1 << 4);
    uint32_t len = (uint32_t)(sl->rec_len & 0xFFF);
    return (uint32_t)(other->rec_len & 0xFFF) == len && memcmp(other + 1, sl + 1, len) == 0;
  }
  /* calls f(slot) for every slot the n_src regions hold; flags what contradicts the plan in counters[2] */
  template <class F> static void each_slot(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units,
                                           unsigned long long* counters, F f) {
    const size_t sb = fq_route_slot_bytes(units);
    for (uint32_t src = 0; src < n_src; src++) {
      const uint8_t* reg = regions + (size_t)src * region_bytes;
      FqRegionHdr h; memcpy(&h, reg, sizeof h);
      if (h.nblocks == 0) continue;
      if (h.nblocks > nblocks || h.stride != stride || h.flags) counters[2] = 1;
      for (uint32_t b = 0; b < std::min(h.nblocks, nblocks); b++) {
        uint32_t cnt; memcpy(&cnt, reg + 16 + 4 * (size_t)b, 4);
        if (cnt > stride) { counters[2] = 1; cnt = (uint32_t)stride; }
        for (uint32_t k = 0; k < cnt; k++) f((const FqRouteSlot*)(reg + 16 + fq_route_counts_bytes(h.nblocks) + ((size_t)b * stride + k) * sb));
      }
    }
  }
  /* the sources' flag words (fq_device.h): written by other processes through the shared-memory "peer" mapping */
  static bool wait_sources(const unsigned long long* flags, unsigned long long expect, uint32_t n_src, unsigned long long* counters) {
    if (!flags) return true;
    for (uint32_t s2 = 0; s2 < n_src; s2++) {
      const volatile unsigned long long* f = flags + s2;
      for (long spins = 0; __atomic_load_n(f, __ATOMIC_ACQUIRE) < expect; spins++) {
        if (spins > 200000000L) { counters[2] = 1; return false; }
        if ((spins & 1023) == 1023) usleep(50);
      }
    }
    return true;
  }
  void shard_insert_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units, FqSlot* slots,
                          unsigned long long mask, unsigned long long* counters, bool, const unsigned long long* flags, unsigned long long expect) override {
    n_launch_++;
    if (!wait_sources(flags, expect, n_src, counters)) return;
    each_slot(regions, n_src, region_bytes, nblocks, stride, units, counters, [&](const FqRouteSlot* sl) {
      unsigned long long i = sl->hash & mask, probes = 0;
      for (;; i = (i + 1) & mask) {
        if (++probes > mask) { counters[2] = 1; break; }
        FqSlot& s = slots[i];
        if (s.hash == FQ_HASH_EMPTY) { s.hash = sl->hash; s.idx1 = (unsigned long long)(uintptr_t)sl >> 4; counters[1]++; break; }
        if (s.hash != sl->hash) continue;
        if (units && !same_name(s.idx1, sl)) continue;
        counters[0]++;
        break;
      }
    });
  }
  void shard_claim_slots(const uint8_t* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t units, FqSlot* slots,
                         unsigned long long mask, unsigned long long* counters, bool, const unsigned long long* flags, unsigned long long expect) override {
    n_launch_++;
    if (!wait_sources(flags, expect, n_src, counters)) return;
    each_slot(regions, n_src, region_bytes, nblocks, stride, units, counters, [&](const FqRouteSlot* sl) {
      unsigned long long i = sl->hash & mask, probes = 0;
      for (;; i = (i + 1) & mask) {
        if (++probes > mask + 1) { counters[9]++; break; }
        FqSlot& s = slots[i];
        if (s.hash == FQ_HASH_EMPTY) { counters[9]++; break; }
        if (s.hash != sl->hash || !same_name(s.idx1, sl)) continue;
        if (s.claim2 == FQ_IDX_NONE) { s.claim2 = sl->rec_len >> 12; counters[8]++; } else counters[9]++;
        break;
      }
    });
  }
  /* "peer memory" of the stand-in: a POSIX shared-memory segment that the other ranks of a gloo test map by name (the handle),
   * so that the peer-memory routing rounds of dist.py run on CPU exactly as they do over CUDA IPC */
  void* ipc_alloc(size_t n, uint8_t handle[64]) override {
    static int counter = 0;
    char name[64]; snprintf(name, sizeof name, "/fqgsim_%d_%d", (int)getpid(), counter++);
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) throw std::runtime_error("stand-in ipc_alloc: shm_open failed");
    if (ftruncate(fd, (off_t)(n ? n : 1)) != 0) { close(fd); shm_unlink(name); throw std::runtime_error("stand-in ipc_alloc: ftruncate failed"); }
    void* p = mmap(nullptr, n ? n : 1, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { shm_unlink(name); throw std::runtime_error("stand-in ipc_alloc: mmap failed"); }
    memset(handle, 0, 64); memcpy(handle, name, strlen(name));
    shm_[p] = Seg{std::string(name), n ? n : 1, true};
    return p;
  }
  void* ipc_open(const uint8_t handle[64]) override {
    char name[65]; memcpy(name, handle, 64); name[64] = 0;
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) throw std::runtime_error("stand-in ipc_open: shm_open failed");
    off_t n = lseek(fd, 0, SEEK_END);
    void* p = mmap(nullptr, (size_t)n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) throw std::runtime_error("stand-in ipc_open: mmap failed");
    shm_[p] = Seg{std::string(name), (size_t)n, false};
    return p;
  }
  void ipc_close(void* p) override { auto it = shm_.find(p); if (it != shm_.end()) { munmap(p, it->second.n); shm_.erase(it); } }
  void ipc_free(void* p) override { auto it = shm_.find(p); if (it != shm_.end()) { munmap(p, it->second.n); if (it->second.owner) shm_unlink(it->second.name.c_str()); shm_.erase(it); } }
  ~FqSimDevice() override { for (auto& kv : shm_) { munmap(kv.first, kv.second.n); if (kv.second.owner) shm_unlink(kv.second.name.c_str()); } }
  void side_copy(void* dst, const void* src, size_t n) override { memmove(dst, src, n); }
  void side_sync() override {}
  void shard_find(const FqPackedName* meta, unsigned long long n, unsigned long long record, unsigned long long* out_pos) override {
    n_launch_++;
    for (unsigned long long m = 0; m < n; m++) if (meta[m].record == record && m < *out_pos) *out_pos = m;
  }
  void names_measure(const FqName* names, uint32_t nrec, unsigned long long* out_units) override {
    n_launch_++;
    for (uint32_t k = 0; k < nrec; k++) if (names[k].hash != FQ_HASH_SKIP) *out_units += (names[k].len + 15u) >> 4;
  }
  void names_gather(FqName* names, const uint8_t* data, uint32_t nrec, uint8_t* arena, unsigned long long* cursor_units) override {
    n_launch_++;
    for (uint32_t k = nrec; k-- > 0;) { /* (backwards: the order of the copies is free, and tests must not rely on it) */
      FqName& nm = names[k];
      if (nm.hash == FQ_HASH_SKIP || nm.len == 0) continue;
      uint32_t units = (nm.len + 15u) >> 4;
      unsigned long long at = *cursor_units * 16ull; *cursor_units += units;
      memset(arena + at, 0, (size_t)units * 16); memcpy(arena + at, data + nm.off, nm.len);
      nm.off = (uint32_t)at;
    }
  }
  void stats_fold(FqStats* const m2[2], FqStats* const o2[2], unsigned long long* const h2[2], unsigned long long* const ho2[2]) override {
    n_launch_++;
    uint32_t lo = std::min(o2[0]->min_rl, o2[1]->min_rl), hi = std::min<uint32_t>(std::max(o2[0]->max_rl, o2[1]->max_rl), FQ_MAX_READ_LENGTH - 1);
    for (int f = 0; f < 2; f++) {
      if (lo <= hi) for (uint32_t l = lo; l <= hi; l++) { h2[f][l] += ho2[f][l]; ho2[f][l] = 0; }
      FqStats* m = m2[f]; FqStats* o = o2[f];
      m->num_rds += o->num_rds; m->mem_sum += o->mem_sum; m->n_names += o->n_names;
      m->min_rl = std::min(m->min_rl, o->min_rl); m->max_rl = std::max(m->max_rl, o->max_rl);
      m->min_q = std::min(m->min_q, o->min_q); m->max_q = std::max(m->max_q, o->max_q);
      o->num_rds = 0; o->mem_sum = 0; o->n_names = 0; o->min_rl = 0xFFFFFFFFu; o->max_rl = 0; o->min_q = 255u; o->max_q = 0;
    }
  }
  void count_n(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out2) override {
    n_launch_++;
    for (uint32_t k = 0; k < n; k++) {
      const FqLine& L = seq_lines[k];
      uint32_t cnt = 0, i = 0;
      for (; i < L.len; i++) { uint8_t c = data[L.off + i]; if (c == '\n' || c == 0) break; if (c == 'N' || c == 'n') cnt++; }
      uint32_t nul = 0; while (nul < L.len && data[L.off + nul] != 0) nul++;
      out2[2 * k] = cnt; out2[2 * k + 1] = nul;
    }
  }
  void header_names(const uint8_t* data, const FqLine* hdr_lines, uint32_t n, int fmt, int is_pe, uint32_t seed, FqName* out) override {
    n_launch_++;
    for (uint32_t k = 0; k < n; k++) out[k] = fq_header_name(data, hdr_lines[k].off, hdr_lines[k].len, fmt, is_pe, seed);
  }
  void names_lookup(const FqTableArgs& a, unsigned long long* out_idx) override {
    n_launch_++;
    for (uint32_t k = 0; k < a.nrec; k++) {
      const FqName& nm = a.names[k];
      uint64_t found = FQ_IDX_NONE;
      if (nm.hash != FQ_HASH_SKIP) {
        uint64_t i = nm.hash & a.mask, probes = 0;
        for (;; i = (i + 1) & a.mask) {
          if (++probes > a.mask + 1) break;
          const FqSlot& s = a.slots[i];
          if (s.hash == FQ_HASH_EMPTY) break;
          if (s.hash != nm.hash) continue;
          uint32_t ol; const uint8_t* on = name_of(a.dir1, a.ndir1, s.idx1, &ol);
          if (!(ol == nm.len && fq_bytes_equal(on, a.data + nm.off, nm.len))) continue;
          found = s.idx1;
          break;
        }
      }
      out_idx[k] = found;
    }
  }
  void poly_at(const uint8_t* data, const FqLine* seq_lines, uint32_t n, uint32_t* out3) override {
    n_launch_++;
    for (uint32_t k = 0; k < n; k++) fq_poly_at(data + seq_lines[k].off, seq_lines[k].len, out3 + 3 * k, out3 + 3 * k + 1, out3 + 3 * k + 2);
  }
  void explain(const uint8_t* data, const FqLine* lines4, const FqRecCtx& cx, FqRecOut* out) override {
    n_launch_++;
    fq_check_record_careful(data, lines4, cx, out);
  }

 private:
  unsigned long long n_launch_ = 0;
  struct Seg { std::string name; size_t n; bool owner; };
  std::map<void*, Seg> shm_;
};

}  // namespace

FqDevice* fq_default_device(int) { return new FqSimDevice(); }

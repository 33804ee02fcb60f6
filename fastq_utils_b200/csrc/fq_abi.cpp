/*
 * fq_abi.cpp — the extern "C" surface declared in include/fastq_gpu.h.  Exceptions stop here.
 */
#include <cstdio>
#include <cstring>
#include <new>
#include "fq_engine.h"

FqDevice* fq_default_device(int ordinal);

struct fqg_ctx {
  FqDevice* dev = nullptr;
  FqEngine* eng = nullptr;
  std::string err;
  double ms = 0.0;
  bool timing = false;
};

static int classify(fqg_ctx* c, const std::exception& ex) {
  if (c) c->err = ex.what();
  if (strstr(ex.what(), "no CUDA")) return FQG_ERR_NO_DEVICE;
  if (strstr(ex.what(), "out of memory")) return FQG_ERR_OOM;
  if (strstr(ex.what(), "CUDA")) return FQG_ERR_CUDA;
  if (strstr(ex.what(), "fqg_")) return FQG_ERR_USAGE;
  return FQG_ERR_INTERNAL;
}
#define FQG_GUARD(c, body)                                   \
  try { body; return 0; }                                    \
  catch (const std::bad_alloc&) { if (c) (c)->err = "host out of memory"; return FQG_ERR_OOM; } \
  catch (const std::exception& ex) { return classify(c, ex); }

extern "C" int fqg_create(const fqg_config* cfg, fqg_ctx** out) {
  if (!cfg || !out) return FQG_ERR_USAGE;
  if (cfg->mode < FQG_MODE_SINGLE || cfg->mode > FQG_MODE_READER) return FQG_ERR_USAGE;
  *out = nullptr;
  fqg_ctx* c = new (std::nothrow) fqg_ctx();
  if (!c) return FQG_ERR_OOM;
  try {
    c->dev = fq_default_device(cfg->device);
    c->eng = new FqEngine(*cfg, c->dev);
  } catch (const std::exception& ex) {
    int rc = classify(c, ex);
    fprintf(stderr, "libfastq_gpu: %s\n", ex.what());
    delete c->eng; delete c->dev; delete c;
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" void fqg_destroy(fqg_ctx* c) {
  if (!c) return;
  try { delete c->eng; delete c->dev; } catch (...) {}
  delete c;
}

extern "C" int fqg_feed(fqg_ctx* c, int file, const void* bytes, size_t n, int last) {
  if (!c || file < 0 || file > 1 || (!bytes && n)) return FQG_ERR_USAGE;
  FQG_GUARD(c, { if (!c->timing) { c->dev->timer_start(); c->timing = true; } c->eng->feed_host(file, bytes, n, last != 0); })
}
extern "C" int fqg_feed_device(fqg_ctx* c, int file, const void* dptr, size_t n, int last) {
  if (!c || file < 0 || file > 1 || (!dptr && n)) return FQG_ERR_USAGE;
  FQG_GUARD(c, { if (!c->timing) { c->dev->timer_start(); c->timing = true; } c->eng->feed_device(file, dptr, n, last != 0); })
}
extern "C" int fqg_finish(fqg_ctx* c, fqg_report* out) {
  if (!c || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, { c->eng->finish(out); if (c->timing) { c->ms = c->dev->timer_stop_ms(); c->timing = false; } })
}
extern "C" int fqg_reset(fqg_ctx* c) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, { c->eng->reset(); c->timing = false; })
}
extern "C" const char* fqg_last_error(const fqg_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" uint64_t fqg_launch_count(const fqg_ctx* c) { return c ? c->dev->launches() : 0; }
extern "C" double fqg_device_ms(const fqg_ctx* c) { return c ? c->ms : 0.0; }

extern "C" int fqg_index_records(fqg_ctx* c, const void* host_bytes, size_t n, uint64_t* starts, size_t cap, uint64_t* n_records) {
  if (!c || (!host_bytes && n) || !n_records || (cap && !starts)) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->index_records(host_bytes, n, starts, cap, n_records))
}

extern "C" int fqg_kernel_stats(fqg_ctx* c, int which, fqg_kernel_stat* out) {
  if (!c || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, { if (!c->dev->kernel_stat(which, &out->ms, &out->launches, &out->bytes, &out->items)) throw std::runtime_error("fqg_kernel_stats: unknown kernel class"); })
}
extern "C" int fqg_path_counts(fqg_ctx* c, uint64_t out[4]) {
  if (!c || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, { for (int i = 0; i < 4; i++) out[i] = c->eng->path_counts[i]; })
}
extern "C" int fqg_memory_stats(fqg_ctx* c, uint64_t out[5]) {
  if (!c || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, { for (int i = 0; i < 4; i++) out[i] = c->eng->mem_stats[i]; out[4] = c->eng->collisions_walked(); })
}
extern "C" int fqg_kernel_stats_reset(fqg_ctx* c) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->kernel_stats_reset())
}

extern "C" int fqg_prescan_device(fqg_ctx* c, int file, const void* dptr, size_t n, int at_eof, uint64_t* n_lines, int32_t* ends_lf, uint64_t first_ends[4]) {
  if (!c || file < 0 || file > 1 || !dptr || !n || !n_lines || !ends_lf || !first_ends) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->prescan_device(file, dptr, n, at_eof != 0, n_lines, ends_lf, first_ends))
}
extern "C" int fqg_set_stream_start(fqg_ctx* c, int file, uint32_t skip_lines, uint64_t first_record) {
  if (!c || file < 0 || file > 1) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_stream_start(file, skip_lines, first_record))
}
extern "C" int fqg_names_count(fqg_ctx* c, int file, uint32_t world, uint64_t* counts, uint64_t* bytes) {
  if (!c || file < 0 || file > 1 || !counts || !bytes) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->names_count(file, world, counts, bytes))
}
extern "C" int fqg_names_pack(fqg_ctx* c, int file, uint32_t world, void* meta, void* blob, const uint64_t* meta_base, const uint64_t* blob_base) {
  if (!c || file < 0 || file > 1 || !meta_base || !blob_base) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->names_pack(file, world, meta, blob, meta_base, blob_base))
}
extern "C" int fqg_shard_insert(fqg_ctx* c, const void* meta, uint64_t n, const void* blob, uint32_t n_src, const uint64_t* meta_start, const uint64_t* blob_start) {
  if (!c || !meta_start || !blob_start) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_insert(meta, n, blob, n_src, meta_start, blob_start))
}
extern "C" int fqg_shard_result(fqg_ctx* c, uint64_t* key, uint64_t* record, char name[1024], uint32_t* name_len, uint64_t* collisions) {
  if (!c || !key || !record || !name || !name_len || !collisions) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_result(key, record, name, name_len, collisions))
}
extern "C" int fqg_hist_range(fqg_ctx* c, int file, uint64_t lo, uint64_t hi, uint64_t* out) {
  if (!c || file < 0 || file > 1 || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->hist_range(file, lo, hi, out))
}

extern "C" int fqg_shard_claim(fqg_ctx* c, const void* meta, uint64_t n, const void* blob, uint32_t n_src, const uint64_t* meta_start, const uint64_t* blob_start, uint64_t step_base) {
  if (!c || !meta_start || !blob_start) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_claim(meta, n, blob, n_src, meta_start, blob_start, step_base))
}
extern "C" int fqg_shard_claim_result(fqg_ctx* c, uint64_t* key, uint64_t* record, char name[1024], uint32_t* name_len, uint64_t* claimed, uint64_t* collisions) {
  if (!c || !key || !record || !name || !name_len || !claimed || !collisions) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_claim_result(key, record, name, name_len, claimed, collisions))
}
extern "C" int fqg_sniff_device(fqg_ctx* c, int file, const void* dptr, size_t n, uint32_t skip, int32_t* fmt, int32_t* color) {
  if (!c || file < 0 || file > 1 || !dptr || !n || !fmt || !color) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->sniff_device(file, dptr, n, skip, fmt, color))
}
extern "C" int fqg_set_sniff(fqg_ctx* c, int file, int32_t fmt, int32_t color) {
  if (!c || file < 0 || file > 1) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_sniff(file, fmt, color))
}
extern "C" int fqg_set_file_total(fqg_ctx* c, int file, uint64_t total) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_file_total(file, total))
}

/* ---- pipelined routing ---- */
extern "C" int fqg_set_chunk_hook(fqg_ctx* c, fqg_chunk_hook hook, void* user) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_chunk_hook(hook, user))
}
extern "C" int fqg_names_new(fqg_ctx* c, int file, uint64_t* n_new) {
  if (!c || file < 0 || file > 1 || !n_new) return FQG_ERR_USAGE;
  FQG_GUARD(c, *n_new = c->eng->names_new(file))
}
extern "C" int fqg_names_pack_slots(fqg_ctx* c, int file, uint32_t world, void* const* region_ptrs, uint64_t region_cap, uint32_t name_units) {
  if (!c || file < 0 || file > 1 || !region_ptrs || !region_cap) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->names_pack_slots(file, world, region_ptrs, region_cap, name_units))
}
extern "C" int fqg_shard_reserve(fqg_ctx* c, uint64_t n_names) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_reserve(n_names))
}
extern "C" int fqg_shard_claim_slots(fqg_ctx* c, const void* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t name_units, int beside,
                                     const void* device_flags, uint64_t expect) {
  if (!c || !regions) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_claim_slots(regions, n_src, region_bytes, nblocks, stride, name_units, beside != 0, device_flags, expect))
}
extern "C" int fqg_shard_insert_slots(fqg_ctx* c, const void* regions, uint32_t n_src, size_t region_bytes, uint32_t nblocks, uint64_t stride, uint32_t name_units, int beside,
                                      const void* device_flags, uint64_t expect) {
  if (!c || !regions) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_insert_slots(regions, n_src, region_bytes, nblocks, stride, name_units, beside != 0, device_flags, expect))
}
extern "C" int fqg_set_route(fqg_ctx* c, int file, uint32_t world, void* const* region_ptrs, size_t region_bytes, uint32_t depth, uint32_t stride, uint32_t name_units) {
  if (!c || file < 0 || file > 1 || (world && !region_ptrs)) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_route(file, world, region_ptrs, region_bytes, depth, stride, name_units))
}
extern "C" int fqg_route_chunks(fqg_ctx* c, int file, uint64_t* n_chunks, int32_t* broken) {
  if (!c || file < 0 || file > 1 || !n_chunks || !broken) return FQG_ERR_USAGE;
  FQG_GUARD(c, { int b = 0; *n_chunks = c->eng->route_chunks(file, &b); *broken = b; })
}
extern "C" int fqg_route_blocks(fqg_ctx* c, uint32_t* nblocks) {
  if (!c || !nblocks) return FQG_ERR_USAGE;
  FQG_GUARD(c, *nblocks = c->dev->lanes_max_blocks())
}
extern "C" size_t fqg_route_region_bytes(uint32_t nblocks, uint64_t stride, uint32_t name_units) { return fq_route_region_bytes(nblocks, stride, name_units); }
extern "C" int fqg_order_after(fqg_ctx* later, int later_side, fqg_ctx* earlier, int earlier_side) {
  if (!later || !earlier) return FQG_ERR_USAGE;
  FQG_GUARD(later, later->dev->order_after(later_side != 0, *earlier->dev, earlier_side != 0))
}
extern "C" int fqg_side_mark(fqg_ctx* c) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->side_mark())
}
extern "C" int fqg_shard_slots_result(fqg_ctx* c, uint64_t* inserted, uint64_t* equal_hashes, int32_t* overflow, uint64_t* claimed, uint64_t* unpaired) {
  if (!c || !inserted || !equal_hashes || !overflow) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->shard_slots_result(inserted, equal_hashes, overflow, claimed, unpaired))
}
extern "C" int fqg_ipc_alloc(fqg_ctx* c, size_t bytes, void** dptr, uint8_t handle[64]) {
  if (!c || !dptr || !handle) return FQG_ERR_USAGE;
  FQG_GUARD(c, *dptr = c->dev->ipc_alloc(bytes, handle))
}
extern "C" int fqg_ipc_open(fqg_ctx* c, const uint8_t handle[64], void** dptr) {
  if (!c || !dptr || !handle) return FQG_ERR_USAGE;
  FQG_GUARD(c, *dptr = c->dev->ipc_open(handle))
}
extern "C" int fqg_ipc_close(fqg_ctx* c, void* dptr) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->ipc_close(dptr))
}
extern "C" int fqg_ipc_free(fqg_ctx* c, void* dptr) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->ipc_free(dptr))
}
extern "C" int fqg_side_copy(fqg_ctx* c, void* dst, const void* src, size_t n) {
  if (!c || ((!dst || !src) && n)) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->side_copy(dst, src, n))
}
extern "C" int fqg_side_copy_lane(fqg_ctx* c, int lane, void* dst, const void* src, size_t n) {
  if (!c || lane < 0 || lane > 7) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->side_copy_lane(lane, dst, src, n))
}
extern "C" int fqg_side_sync(fqg_ctx* c) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->side_sync())
}
extern "C" int fqg_set_line_hint(fqg_ctx* c, int file, uint32_t len) {
  if (!c || file < 0 || file > 1) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_line_hint(file, len))
}
extern "C" int fqg_records_fed(fqg_ctx* c, int file, uint64_t* n) {
  if (!c || file < 0 || file > 1 || !n) return FQG_ERR_USAGE;
  FQG_GUARD(c, *n = c->eng->records_fed(file))
}
extern "C" int fqg_set_hash_seed(fqg_ctx* c, uint32_t seed) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->set_hash_seed(seed))
}

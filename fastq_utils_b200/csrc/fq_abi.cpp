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
  if (cfg->mode < FQG_MODE_SINGLE || cfg->mode > FQG_MODE_SORTED_PAIR) return FQG_ERR_USAGE;
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
  if (!c || (!host_bytes && n) || !n_records) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->eng->index_records(host_bytes, n, starts, cap, n_records))
}

extern "C" int fqg_kernel_stats(fqg_ctx* c, int which, fqg_kernel_stat* out) {
  if (!c || !out) return FQG_ERR_USAGE;
  FQG_GUARD(c, { if (!c->dev->kernel_stat(which, &out->ms, &out->launches, &out->bytes, &out->items)) throw std::runtime_error("fqg_kernel_stats: unknown kernel class"); })
}
extern "C" int fqg_kernel_stats_reset(fqg_ctx* c) {
  if (!c) return FQG_ERR_USAGE;
  FQG_GUARD(c, c->dev->kernel_stats_reset())
}

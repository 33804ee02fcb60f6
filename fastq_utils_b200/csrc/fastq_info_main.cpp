/*
 * fastq_info_gpu — the reference's fastq_info command line (src/fastq_info.c:190-396) on top of libfastq_gpu.
 * Same options (-r -s -e -q -h, "pe"), same stdout / stderr text, same exit status.  The host only inflates
 * (zlib, like the reference's gzopen/gzgets, src/fastq.c:631-661; gzip members are concatenated, plain text passes through,
 * `-` is standard input) and prints; every FASTQ byte is checked on the GPU.  The files are streamed: the library asks for the
 * operands it would open (fqg_fastq_info_stream) and reads them piece by piece through the callbacks below, inflating one piece
 * on a helper thread while the previous one is copied to the device and validated.
 */
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/fastq_gpu.h"

static void* gz_open(void*, const char* name) {
  gzFile fd = strcmp(name, "-") == 0 ? gzdopen(fileno(stdin), "r") : gzopen(name, "r");
  if (fd) gzbuffer(fd, 1 << 20);
  return fd;
}
static long gz_read(void*, void* h, void* buf, size_t cap) {
  const unsigned want = cap > (1u << 30) ? (1u << 30) : (unsigned)cap;
  return gzread((gzFile)h, buf, want);
}
static void gz_close(void*, void* h) { gzclose((gzFile)h); }

int main(int argc, char** argv) {
  fqg_stream_io io; io.user = nullptr; io.open = gz_open; io.read = gz_read; io.close = gz_close;
  fqg_transcript t; memset(&t, 0, sizeof t);
  const char* dev = getenv("FQG_DEVICE");
  const char* piece = getenv("FQG_PIECE_BYTES");
  int st = fqg_fastq_info_stream(argc, (const char**)argv, &io, dev ? atoi(dev) : 0, piece ? (size_t)strtoull(piece, nullptr, 10) : 0, &t);
  if (st != 0) { fprintf(stderr, "\nERROR: libfastq_gpu failed (%d)\n", st); return 2; } /* SYS_INT_ERROR_EXIT_STATUS */
  fwrite(t.out, 1, t.out_len, stdout);
  fwrite(t.err, 1, t.err_len, stderr);
  int rc = t.rc;
  fqg_transcript_free(&t);
  return rc;
}

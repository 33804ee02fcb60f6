/*
 * fastq_info_gpu — the reference's fastq_info command line (src/fastq_info.c:190-396) on top of libfastq_gpu.
 * Same options (-r -s -e -q -h, "pe"), same stdout / stderr text, same exit status.  The host only inflates
 * (zlib, like the reference's gzopen/gzgets, src/fastq.c:631-661) and prints; every FASTQ byte is checked on the GPU.
 */
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/fastq_gpu.h"

static bool slurp(const char* path, std::vector<unsigned char>& out) {
  gzFile fd = strcmp(path, "-") == 0 ? gzdopen(fileno(stdin), "r") : gzopen(path, "r");
  if (!fd) return false;
  gzbuffer(fd, 1 << 20);
  std::vector<unsigned char> buf(8u << 20);
  for (;;) {
    int k = gzread(fd, buf.data(), (unsigned)buf.size());
    if (k <= 0) break;
    out.insert(out.end(), buf.begin(), buf.begin() + k);
  }
  gzclose(fd);
  return true;
}

int main(int argc, char** argv) {
  /* the file operands are the words getopt leaves behind */
  std::vector<const char*> pos;
  bool stop = false;
  for (int i = 1; i < argc; i++) {
    if (!stop && !strcmp(argv[i], "--")) { stop = true; continue; }
    if (!stop && argv[i][0] == '-' && argv[i][1]) continue;
    pos.push_back(argv[i]);
  }
  std::vector<unsigned char> d1, d2;
  size_t n1 = (size_t)-1, n2 = (size_t)-1;
  if (pos.size() >= 1 && slurp(pos[0], d1)) n1 = d1.size();
  if (pos.size() >= 2 && strncmp(pos[1], "pe", 2) != 0 && slurp(pos[1], d2)) n2 = d2.size();
  fqg_transcript t; memset(&t, 0, sizeof t);
  const char* dev = getenv("FQG_DEVICE");
  int st = fqg_fastq_info_mem(argc, (const char**)argv, d1.data(), n1, d2.data(), n2, dev ? atoi(dev) : 0, (size_t)1 << 30, &t);
  if (st != 0) { fprintf(stderr, "\nERROR: libfastq_gpu failed (%d)\n", st); return 2; } /* SYS_INT_ERROR_EXIT_STATUS */
  fwrite(t.out, 1, t.out_len, stdout);
  fwrite(t.err, 1, t.err_len, stderr);
  int rc = t.rc;
  fqg_transcript_free(&t);
  return rc;
}

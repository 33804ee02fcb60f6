/*
 * fq_synth.cu — synthetic FASTQ generators writing straight into device memory (bench.py and the large parity tests).
 * Counter-based (splitmix64 keyed by seed and record index): any shard can be produced independently on any GPU.
 *
 * Illumina record i (fixed width, 359 bytes; SURVEY.md §8d config 3/4 with zero-free coordinate ranges so the width is constant):
 *   @A00123:45:HXXXXXXXX:1:TTTT:XXXXX:YYYYY M:N:0:ACGTACGT\n   TTTT = 1101 + i / 4e8, XXXXX = 10000 + i % 20000,
 *   <150 bases, ACGT uniform, N with p = 1/1024>\n+\n           YYYYY = 10000 + (i / 20000) % 20000, M = mate (1 or 2)
 *   <150 qualities in [35,73]>\n                                record 0 holds both 35 and 73
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fastq_gpu.h"

namespace {
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
constexpr int ILL_HDR = 55, ILL_LEN = 150, ILL_REC = ILL_HDR + ILL_LEN + 1 + 2 + ILL_LEN + 1;

__device__ void put_dec(uint8_t* p, uint32_t v, int digits) { for (int i = digits - 1; i >= 0; i--) { p[i] = '0' + v % 10; v /= 10; } }

/* one warp per record: lane-strided byte writes (coalesced) */
__global__ void fq_synth_illumina_kernel(uint8_t* out, uint64_t first, uint64_t nrec, uint64_t seed, int mate, const uint64_t* perm_window) {
  const int lane = threadIdx.x & 31;
  uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < nrec; r += nwarps) {
    uint64_t slot = first + r, i = slot;
    if (perm_window) { /* mate file: records permuted inside windows of *perm_window (keeps pairs, breaks the order) */
      uint64_t w = *perm_window, base = slot / w * w, k = slot - base;
      i = base + (k * 7 + 3) % w; /* w is a power of two: an odd multiplier permutes the window */
    }
    uint8_t* p = out + r * (uint64_t)ILL_REC;
    if (lane == 0) {
      const char* pre = "@A00123:45:HXXXXXXXX:1:";
      for (int k = 0; k < 23; k++) p[k] = pre[k];
      put_dec(p + 23, 1101 + (uint32_t)(i / 400000000ull), 4); p[27] = ':';
      put_dec(p + 28, 10000 + (uint32_t)(i % 20000), 5); p[33] = ':';
      put_dec(p + 34, 10000 + (uint32_t)((i / 20000) % 20000), 5);
      const char* suf = " 1:N:0:ACGTACGT\n";
      for (int k = 0; k < 16; k++) p[39 + k] = suf[k];
      p[40] = mate == 2 ? '2' : '1';
      p[ILL_HDR + ILL_LEN] = '\n'; p[ILL_HDR + ILL_LEN + 1] = '+'; p[ILL_HDR + ILL_LEN + 2] = '\n';
      p[ILL_REC - 1] = '\n';
    }
    uint8_t* sq = p + ILL_HDR; uint8_t* ql = p + ILL_HDR + ILL_LEN + 3;
    for (int k = lane; k < ILL_LEN; k += 32) {
      uint64_t h = splitmix64(seed ^ (i * 0x100000001B3ull + (uint64_t)k * 2 + (uint64_t)mate * 0x51ED27ull));
      uint32_t b = (uint32_t)h & 3u;
      sq[k] = ((h >> 8) & 1023u) == 0 ? 'N' : (b == 0 ? 'A' : b == 1 ? 'C' : b == 2 ? 'G' : 'T');
      uint8_t q = 35 + (uint8_t)((h >> 32) % 39u);
      if (i == 0 && k == 0) q = 35;
      if (i == 0 && k == 1) q = 73;
      ql[k] = q;
    }
  }
}

/* long reads: record r occupies [offsets[r], offsets[r+1]); sequence length = (record bytes - header - 4) / 2 */
__global__ void fq_synth_long_kernel(uint8_t* out, const uint64_t* offsets, uint64_t first, uint64_t nrec, uint64_t seed, uint32_t hdr_len) {
  for (uint64_t r = blockIdx.x; r < nrec; r += gridDim.x) {
    uint64_t i = first + r, o = offsets[r] - offsets[0], len = (offsets[r + 1] - offsets[r] - hdr_len - 4) / 2;
    uint8_t* p = out + o;
    if (threadIdx.x == 0) { /* @<16 hex> runid=<8 hex> read=<10 dec> ch=<3 dec> start_time=2024-01-01T00:00:00Z\n */
      uint64_t h = splitmix64(seed ^ (i * 0x9E3779B1ull));
      int k = 0; p[k++] = '@';
      for (int d = 0; d < 16; d++) { uint32_t x = (h >> (4 * d)) & 15; p[k++] = x < 10 ? '0' + x : 'a' + x - 10; }
      const char* a = " runid="; for (int d = 0; a[d]; d++) p[k++] = a[d];
      for (int d = 0; d < 8; d++) { uint32_t x = (seed >> (4 * d)) & 15; p[k++] = x < 10 ? '0' + x : 'a' + x - 10; }
      const char* b = " read="; for (int d = 0; b[d]; d++) p[k++] = b[d];
      put_dec(p + k, (uint32_t)(i % 4000000000ull), 10); k += 10;
      const char* c = " ch="; for (int d = 0; c[d]; d++) p[k++] = c[d];
      put_dec(p + k, 1 + (uint32_t)(i % 512), 3); k += 3;
      const char* e = " start_time=2024-01-01T00:00:00Z\n"; for (int d = 0; e[d]; d++) p[k++] = e[d];
      p[hdr_len + len] = '\n'; p[hdr_len + len + 1] = '+'; p[hdr_len + len + 2] = '\n'; p[hdr_len + 2 * len + 3] = '\n';
    }
    uint8_t* sq = p + hdr_len; uint8_t* ql = p + hdr_len + len + 3;
    for (uint64_t k = threadIdx.x; k < len; k += blockDim.x) {
      uint64_t h = splitmix64(seed ^ (i * 0x100000001B3ull + k));
      uint32_t b = (uint32_t)h & 3u;
      sq[k] = b == 0 ? 'A' : b == 1 ? 'C' : b == 2 ? 'G' : 'T';
      ql[k] = 35 + (uint8_t)((h >> 32) % 59u);
    }
  }
}
}  // namespace

extern "C" int fqg_synth_illumina_record_bytes(void) { return ILL_REC; }
extern "C" int fqg_synth_long_header_bytes(void) { return 1 + 16 + 7 + 8 + 6 + 10 + 4 + 3 + 33; }

extern "C" int fqg_synth_illumina(void* device_out, uint64_t first_record, uint64_t n_records, uint64_t seed, int mate,
                                  uint64_t perm_window, void* cuda_stream) {
  if (!device_out || !n_records) return FQG_ERR_USAGE;
  uint64_t* dw = nullptr;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (perm_window > 1) {
    if (perm_window & (perm_window - 1)) return FQG_ERR_USAGE;
    if (cudaMallocAsync(&dw, sizeof(uint64_t), st) != cudaSuccess) return FQG_ERR_OOM;
    cudaMemcpyAsync(dw, &perm_window, sizeof perm_window, cudaMemcpyHostToDevice, st);
  }
  fq_synth_illumina_kernel<<<148 * 16, 256, 0, st>>>((uint8_t*)device_out, first_record, n_records, seed, mate, dw);
  cudaError_t e = cudaGetLastError();
  if (dw) { cudaStreamSynchronize(st); cudaFreeAsync(dw, st); }
  return e == cudaSuccess ? 0 : FQG_ERR_CUDA;
}
extern "C" int fqg_synth_longreads(void* device_out, const uint64_t* device_offsets, uint64_t first_record, uint64_t n_records,
                                   uint64_t seed, void* cuda_stream) {
  if (!device_out || !device_offsets || !n_records) return FQG_ERR_USAGE;
  fq_synth_long_kernel<<<148 * 8, 256, 0, (cudaStream_t)cuda_stream>>>((uint8_t*)device_out, device_offsets, first_record, n_records, seed,
                                                                       (uint32_t)fqg_synth_long_header_bytes());
  return cudaGetLastError() == cudaSuccess ? 0 : FQG_ERR_CUDA;
}

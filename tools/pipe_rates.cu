// Micro-benchmark: issue rate (warp-instructions / clock / SM) of the integer instructions the FASTQ kernels are made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define DEF(name, BODY)                                                                                   \
  __global__ void __launch_bounds__(512) k_##name(uint32_t* out, long long* cyc, uint32_t seed) {         \
    uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, \
             a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;                                                          \
    uint32_t c = seed | 0x0A0A0A0Au, d = seed * 77u;                                                      \
    __syncthreads();                                                                                      \
    long long t0 = clock64();                                                                             \
    _Pragma("unroll 4") for (int i = 0; i < ITERS; i++) { BODY }                                          \
    long long t1 = clock64();                                                                             \
    __syncthreads();                                                                                      \
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;                   \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                      \
  }
#define R8(OP) OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#define LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(c), "r"(d));
#define IADD(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(c));
#define IMADSHL(x) asm volatile("mul.lo.u32 %0, %0, 128;" : "+r"(x));
#define IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define SHR(x) asm volatile("shr.u32 %0, %0, 3;" : "+r"(x));
#define SHL(x) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));
#define PRMT(x) asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(x) : "r"(c));
#define POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x));
#define DP4A(x) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define VMIN2(x) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(c));
#define VMIN3(x) asm volatile("{ .reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2; }" : "+r"(x) : "r"(c), "r"(d));
#define FFS(x) asm volatile("{ .reg .u32 t; brev.b32 t, %0; bfind.shiftamt.u32 %0, t; }" : "+r"(x));
#define BFIND(x) asm volatile("bfind.u32 %0, %0;" : "+r"(x));
#define VABS4(x) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define VSETEQ4(x) asm volatile("vset4.u32.u32.eq %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define SHF(x) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define MIXLI(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mul.lo.u32 %0, %0, 128;" : "+r"(x) : "r"(c), "r"(d));
#define MIXLA(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.lo.u32 %0, %0, 1, %1;" : "+r"(x) : "r"(c), "r"(d));
#define MIX21(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; lop3.b32 %0, %0, %2, %1, 0xe8; mad.lo.u32 %0, %0, 1, %1;" : "+r"(x) : "r"(c), "r"(d));
#define MIXLP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; popc.b32 %0, %0;" : "+r"(x) : "r"(c), "r"(d));
#define MIXLD(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define MIXID(x) asm volatile("mul.lo.u32 %0, %0, 128; dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(c), "r"(d));
#define SHFL(x) x = __shfl_xor_sync(0xFFFFFFFFu, x, 1);
#define BALLOT(x) x = __ballot_sync(0xFFFFFFFFu, x & 1);
DEF(lop3, R8(LOP3)) DEF(iadd, R8(IADD)) DEF(imadshl, R8(IMADSHL)) DEF(imad, R8(IMAD)) DEF(shr, R8(SHR)) DEF(shl, R8(SHL))
DEF(prmt, R8(PRMT)) DEF(popc, R8(POPC)) DEF(dp4a, R8(DP4A)) DEF(vmin2, R8(VMIN2)) DEF(min3, R8(VMIN3)) DEF(ffs, R8(FFS)) DEF(bfind, R8(BFIND))
DEF(vabs4, R8(VABS4)) DEF(vseteq4, R8(VSETEQ4)) DEF(shf, R8(SHF)) DEF(mix_lop_imadshl, R8(MIXLI)) DEF(mix_lop_imadadd, R8(MIXLA))
DEF(mix_2lop_1imad, R8(MIX21)) DEF(mix_lop_popc, R8(MIXLP)) DEF(mix_lop_dp4a, R8(MIXLD)) DEF(mix_imad_dp4a, R8(MIXID))
DEF(shfl, R8(SHFL)) DEF(ballot, R8(BALLOT))

__global__ void __launch_bounds__(512) k_lds128(uint32_t* out, long long* cyc, uint32_t seed) {
  __shared__ uint4 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += 512) buf[i] = make_uint4(i, seed, i * 3, 7);
  __syncthreads();
  uint32_t acc = 0; uint32_t idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < ITERS; i++) { uint4 v = buf[(idx + i * 32) & 2047]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K> void run(const char* name, K k, int per_iter, uint32_t* out, long long* cyc) {
  k<<<148, 512>>>(out, cyc, 12345u);
  cudaDeviceSynchronize();
  k<<<148, 512>>>(out, cyc, 12345u);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  printf("%-22s %7.3f warp-inst/clk/SM  (%d inst/iter, %.0f cycles)%s\n", name, 16.0 * ITERS * per_iter / avg, per_iter, avg, e ? " ERROR" : "");
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
#define RUN(n, per) run(#n, k_##n, per, out, cyc)
  RUN(lop3, 8); RUN(iadd, 8); RUN(imadshl, 8); RUN(imad, 8); RUN(shr, 8); RUN(shl, 8); RUN(shf, 8); RUN(prmt, 8); RUN(popc, 8); RUN(dp4a, 8);
  RUN(vmin2, 8); RUN(min3, 16); RUN(ffs, 16); RUN(bfind, 8); RUN(vabs4, 8); RUN(vseteq4, 8);
  RUN(mix_lop_imadshl, 16); RUN(mix_lop_imadadd, 16); RUN(mix_2lop_1imad, 24); RUN(mix_lop_popc, 16); RUN(mix_lop_dp4a, 16); RUN(mix_imad_dp4a, 16);
  RUN(shfl, 8); RUN(ballot, 8); RUN(lds128, 1);
  return 0;
}

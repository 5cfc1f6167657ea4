// Second pipe microbenchmark (B200, sm_100a): exact SASS forms used by the Goldilocks kernels.
// Every kernel is ITER x UNR groups on 4-8 independent chains per thread; the SASS of each body is
// checked with cuobjdump (tools/microbench/README.md).  Reports warp-instructions/clk/SMSP per SASS op.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 1000
typedef uint32_t u32; typedef uint64_t u64;
#define K(name) __global__ void __launch_bounds__(256) name(u32* out, u32 seed, u32 m1, u32 m2)
#define INIT32(n) u32 a[n]; for (int i = 0; i < n; i++) a[i] = seed * (i + 3) + threadIdx.x;
#define FIN32(n) { u32 r = 0; for (int i = 0; i < n; i++) r ^= a[i]; out[blockIdx.x * blockDim.x + threadIdx.x] = r; }

// 4 chains x (lo,hi): pure accumulating IMAD.WIDE.U32
K(k_wide_acc) { INIT32(8)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
  } FIN32(8) }
// WIDE with carry-out captured by an IADD3.X into a third word
K(k_wide_cout) { INIT32(12)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 12; j += 3) asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(a[j]), "+r"(a[j+1]), "+r"(a[j+2]) : "r"(m1), "r"(m2));
  } FIN32(12) }
// mul.wide with fresh operands (no addend): a = lo(a)*m1 as 64-bit
K(k_wide_mul) { u64 a[4]; for (int i = 0; i < 4; i++) a[i] = seed * (i + 3) + threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 4; j++) asm volatile("{.reg .u32 t, w; mov.b64 {t, w}, %0; xor.b32 t, t, w; mul.wide.u32 %0, t, %1;}" : "+l"(a[j]) : "r"(m1));
  } out[blockIdx.x * blockDim.x + threadIdx.x] = (u32)(a[0] ^ a[1] ^ a[2] ^ a[3]); }
// carry chain of IADD3.X: one add.cc then 7 addc.cc
K(k_iadd3x) { INIT32(8)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
      asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %8; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, %8; addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %8; addc.cc.u32 %6, %6, %8; addc.u32 %7, %7, %8;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]) : "r"(m1));
  } FIN32(8) }
// independent 2-instruction carry pairs on 8 chains: add.cc + addc
K(k_addpair) { INIT32(16)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int j = 0; j < 16; j += 2) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
  } FIN32(16) }
K(k_sel) { INIT32(8)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j++) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a[j]) : "r"(m1), "r"(m2));
  } FIN32(8) }
K(k_i2f) { INIT32(8) double d[8]; for (int i = 0; i < 8; i++) d[i] = 0;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j++) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[j]) : "r"(a[j])); a[j] += (u32)__double2loint(d[j]); }
  } FIN32(8) }
// mixes (per group): W = accumulating WIDE pair, X = add.cc+addc pair (IADD3 + IADD3.X), F = DFMA
K(k_mix_w_x) { INIT32(16)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[8+j]), "+r"(a[9+j]) : "r"(m1), "r"(m2));
      }
  } FIN32(16) }
K(k_mix_w_2x) { INIT32(24)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[8+j]), "+r"(a[9+j]) : "r"(m1), "r"(m2));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[16+j]), "+r"(a[17+j]) : "r"(m1), "r"(m2));
      }
  } FIN32(24) }
K(k_mix_w_f) { INIT32(8) double d[4]; for (int i = 0; i < 4; i++) d[i] = seed + i; double fb = seed * 1e-3, fc = threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j/2]) : "d"(fb), "d"(fc));
      }
  } a[0] ^= (u32)(d[0] + d[1] + d[2] + d[3]); FIN32(8) }
K(k_mix_w_2f) { INIT32(8) double d[8]; for (int i = 0; i < 8; i++) d[i] = seed + i; double fb = seed * 1e-3, fc = threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(fb), "d"(fc));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j+1]) : "d"(fb), "d"(fc));
      }
  } a[0] ^= (u32)(d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6] + d[7]); FIN32(8) }
K(k_mix_w_x_f) { INIT32(16) double d[4]; for (int i = 0; i < 4; i++) d[i] = seed + i; double fb = seed * 1e-3, fc = threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2));
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[8+j]), "+r"(a[9+j]) : "r"(m1), "r"(m2));
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j/2]) : "d"(fb), "d"(fc));
      }
  } a[0] ^= (u32)(d[0] + d[1] + d[2] + d[3]); FIN32(16) }


#define WACC(j) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(m1), "r"(m2))
#define LOP(j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(m1), "r"(m2))
#define ADD(j) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(a[(j) ^ 1]))
#define SHF(j) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[j]) : "r"(m2))
#define MIXK(name, n, BODY) K(name) { INIT32(n) for (int it = 0; it < ITER; it++) { _Pragma("unroll") for (int u = 0; u < 4; u++) _Pragma("unroll") for (int j = 0; j < 8; j += 2) { BODY } } FIN32(n) }
MIXK(k_w_1lop, 16, WACC(j); LOP(8 + j);)
MIXK(k_w_2lop, 16, WACC(j); LOP(8 + j); LOP(9 + j);)
MIXK(k_w_4lop, 24, WACC(j); LOP(8 + j); LOP(9 + j); LOP(16 + j); LOP(17 + j);)
MIXK(k_w_1add, 16, WACC(j); ADD(8 + j);)
MIXK(k_w_2add, 16, WACC(j); ADD(8 + j); ADD(9 + j);)
MIXK(k_w_4add, 24, WACC(j); ADD(8 + j); ADD(9 + j); ADD(16 + j); ADD(17 + j);)
MIXK(k_w_2shf, 16, WACC(j); SHF(8 + j); SHF(9 + j);)
MIXK(k_w_2add_2lop, 24, WACC(j); ADD(8 + j); ADD(9 + j); LOP(16 + j); LOP(17 + j);)
MIXK(k_2add_2lop, 24, ADD(8 + j); ADD(9 + j); LOP(16 + j); LOP(17 + j);)
MIXK(k_4add_2lop, 24, ADD(8 + j); ADD(9 + j); ADD(j); ADD(j + 1); LOP(16 + j); LOP(17 + j);)
// one IADD3 + three IADD3.X (a 128-bit add) beside two WIDE
K(k_w_x4) { INIT32(16)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 4) { WACC(j); WACC(j + 2);
        asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %6; addc.u32 %3, %3, %7;"
          : "+r"(a[8 + j]), "+r"(a[9 + j]), "+r"(a[10 + j]), "+r"(a[11 + j]) : "r"(a[11 + j]), "r"(a[10 + j]), "r"(a[9 + j]), "r"(a[8 + j])); }
  } FIN32(16) }
K(k_x4) { INIT32(16)
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %6; addc.u32 %3, %3, %7;"
          : "+r"(a[8 + j]), "+r"(a[9 + j]), "+r"(a[10 + j]), "+r"(a[11 + j]) : "r"(a[11 + j]), "r"(a[10 + j]), "r"(a[9 + j]), "r"(a[8 + j])); }
  } FIN32(16) }

template <typename Kn> void run(const char* name, Kn k, double groups_per_iter, const char* what, u32* d_out, int sms, double mhz, int tpb, int bps) {
  int blocks = sms * bps;
  k<<<blocks, tpb>>>(d_out, 12345, 0x9E3779B1u, 0x85EBCA77u); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<<<blocks, tpb>>>(d_out, 12345, 0x9E3779B1u, 0x85EBCA77u); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps_per_smsp = (double)bps * tpb / 32 / 4;
  double cycles = ms * 1e-3 * mhz * 1e6;
  printf("%-12s %7.3f ms  %6.2f cycles per group per warp-slot  (%s; %.0f warps/SMSP)\n", name, ms, cycles / (ITER * groups_per_iter * warps_per_smsp), what, warps_per_smsp);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1e3; printf("%s SMs=%d clock=%.0f MHz (cycle figures assume the GPU runs at this clock)\n", p.name, sms, mhz);
  u32* d; cudaMalloc(&d, sms * 8 * 256 * 4);
  for (int bps = 4; bps >= 2; bps /= 2) {
    printf("-- %d blocks x 256 threads per SM\n", bps);
    run("wide_acc", k_wide_acc, 16, "group = 1 IMAD.WIDE acc", d, sms, mhz, 256, bps);
    run("wide_cout", k_wide_cout, 16, "group = 1 IMAD.WIDE carry-out + 1 IADD3.X", d, sms, mhz, 256, bps);
    run("wide_mul", k_wide_mul, 16, "group = LOP3 + 1 IMAD.WIDE no addend", d, sms, mhz, 256, bps);
    run("iadd3x", k_iadd3x, 32, "group = 1 IADD3(.X) in an 8-long carry chain", d, sms, mhz, 256, bps);
    run("addpair", k_addpair, 16, "group = IADD3 + IADD3.X", d, sms, mhz, 256, bps);
    run("sel", k_sel, 32, "group = ISETP + SEL", d, sms, mhz, 256, bps);
    run("i2f", k_i2f, 32, "group = I2F.F64.U32 + IADD", d, sms, mhz, 256, bps);
    run("w+x", k_mix_w_x, 16, "group = WIDE + IADD3 + IADD3.X", d, sms, mhz, 256, bps);
    run("w+2x", k_mix_w_2x, 16, "group = WIDE + 2 IADD3 + 2 IADD3.X", d, sms, mhz, 256, bps);
    run("w+f", k_mix_w_f, 16, "group = WIDE + DFMA", d, sms, mhz, 256, bps);
    run("w+2f", k_mix_w_2f, 16, "group = WIDE + 2 DFMA", d, sms, mhz, 256, bps);
    run("w+x+f", k_mix_w_x_f, 16, "group = WIDE + IADD3 + IADD3.X + DFMA", d, sms, mhz, 256, bps);
    run("w+1lop", k_w_1lop, 16, "group = WIDE + LOP3", d, sms, mhz, 256, bps);
    run("w+2lop", k_w_2lop, 16, "group = WIDE + 2 LOP3", d, sms, mhz, 256, bps);
    run("w+4lop", k_w_4lop, 16, "group = WIDE + 4 LOP3", d, sms, mhz, 256, bps);
    run("w+1add", k_w_1add, 16, "group = WIDE + IADD3", d, sms, mhz, 256, bps);
    run("w+2add", k_w_2add, 16, "group = WIDE + 2 IADD3", d, sms, mhz, 256, bps);
    run("w+4add", k_w_4add, 16, "group = WIDE + 4 IADD3", d, sms, mhz, 256, bps);
    run("w+2shf", k_w_2shf, 16, "group = WIDE + 2 SHF", d, sms, mhz, 256, bps);
    run("w+2add+2lop", k_w_2add_2lop, 16, "group = WIDE + 2 IADD3 + 2 LOP3", d, sms, mhz, 256, bps);
    run("2add+2lop", k_2add_2lop, 16, "group = 2 IADD3 + 2 LOP3", d, sms, mhz, 256, bps);
    run("4add+2lop", k_4add_2lop, 16, "group = 4 IADD3 + 2 LOP3", d, sms, mhz, 256, bps);
    run("2w+add128", k_w_x4, 8, "group = 2 WIDE + IADD3 + 3 IADD3.X", d, sms, mhz, 256, bps);
    run("add128", k_x4, 8, "group = IADD3 + 3 IADD3.X", d, sms, mhz, 256, bps);
  }
  return 0;
}

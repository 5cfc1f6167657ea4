// Throughput microbenchmark of the integer instructions the Goldilocks/Poseidon kernels are made of
// (B200, sm_100a).  Each kernel runs ITER x 64 instructions of one kind on 8 independent chains per
// thread, 1024 threads x (4 blocks per SM); reports warp-instructions per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2000
typedef uint32_t u32; typedef uint64_t u64;

#define KERNEL(name, DECL, BODY, OUT) \
__global__ void __launch_bounds__(256) name(u32* out, u32 seed) { \
  DECL; \
  for (int it = 0; it < ITER; it++) { _Pragma("unroll") for (int u = 0; u < 8; u++) { BODY } } \
  out[blockIdx.x * blockDim.x + threadIdx.x] = OUT; }

#define DECL8_32 u32 a[8], b = seed | 1, c = threadIdx.x; for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x
#define DECL8_64 u64 a[8]; u32 b = seed | 1, c = threadIdx.x | 3; for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x
#define OUT32 (a[0]^a[1]^a[2]^a[3]^a[4]^a[5]^a[6]^a[7])
#define OUT64 (u32)((a[0]^a[1]^a[2]^a[3]^a[4]^a[5]^a[6]^a[7]) >> 7)

// 8 chains x 8 ops per iteration => 64 ops
KERNEL(k_imad_lo, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));, OUT32)
KERNEL(k_imad_hi, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));, OUT32)
// the multiplier is the low word of the chain itself, so the product cannot be hoisted out of the loop
#define WIDE_DEP(j) asm volatile("{.reg .u32 t, u; mov.b64 {t, u}, %0; mad.wide.u32 %0, t, %1, %0;}" : "+l"(a[j]) : "r"(b))
KERNEL(k_imad_wide, DECL8_64, for (int j = 0; j < 8; j++) WIDE_DEP(j);, OUT64)
KERNEL(k_iadd3, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(b));, OUT32)
KERNEL(k_lop3, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));, OUT32)
KERNEL(k_add64, DECL8_64, for (int j = 0; j < 8; j++) asm volatile("add.u64 %0, %0, %1;" : "+l"(a[j]) : "l"((u64)b << 20 | c));, OUT64)
KERNEL(k_shf, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(a[j]) : "r"(b));, OUT32)
// mixes: per j one IMAD.WIDE + n IADD3 (independent registers)
#define DECLMIX u64 a[8]; u32 d[8]; u32 b = seed | 1, c = threadIdx.x | 3; for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; d[i] = i + seed; }
#define OUTMIX (u32)((a[0]^a[1]^a[2]^a[3]^a[4]^a[5]^a[6]^a[7]) >> 7) ^ d[0]^d[1]^d[2]^d[3]^d[4]^d[5]^d[6]^d[7]
KERNEL(k_mix_w1a1, DECLMIX, for (int j = 0; j < 4; j++) { WIDE_DEP(j); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j]) : "r"(b), "r"(c)); }, OUTMIX)
KERNEL(k_mix_w1a3, DECLMIX, for (int j = 0; j < 2; j++) { WIDE_DEP(j); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j+2]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j+4]) : "r"(c), "r"(b)); }, OUTMIX)
// three pipes at once: 1 WIDE (fmaheavy) + 2 LOP3 (alu) + 1 DFMA (fp64) per group, two groups per iteration
#define DECLMIX3 DECLMIX double f[8]; for (int i = 0; i < 8; i++) f[i] = seed + i; double fb = seed * 1e-3, fc = threadIdx.x;
#define OUTMIX3 (OUTMIX) ^ (u32)(f[0] + f[1])
KERNEL(k_mix_w1a2d1, DECLMIX3, for (int j = 0; j < 2; j++) { WIDE_DEP(j); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[j+2]) : "r"(b), "r"(c)); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j]) : "d"(fb), "d"(fc)); }, OUTMIX3)
// IADD3 with three live operands (cannot become IMAD.IADD) and the 64-bit add pair
KERNEL(k_iadd3_3op, DECL8_32, for (int j = 0; j < 8; j++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[j]) : "r"(b), "r"(a[(j + 1) & 7]));, OUT32)
KERNEL(k_mix_l1a1, DECLMIX, for (int j = 0; j < 4; j++) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d[j]) : "r"(b), "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(d[j+4]) : "r"(b)); }, OUTMIX)
// carry chains: add.cc / addc pairs (64-bit add as two 32-bit ops)
KERNEL(k_addcc, DECL8_32, for (int j = 0; j < 8; j += 2) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(b), "r"(c));, OUT32)
// mad.lo.cc + madc.hi (what ptxas fuses into IMAD.WIDE with carry)
KERNEL(k_madcc, DECL8_32, for (int j = 0; j < 8; j += 2) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[j]), "+r"(a[j+1]) : "r"(b), "r"(c));, OUT32)
KERNEL(k_dfma, double a[8]; double b = seed * 1e-3; double c = threadIdx.x; for (int i = 0; i < 8; i++) a[i] = seed + i, for (int j = 0; j < 8; j++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[j]) : "d"(b), "d"(c));, (u32)(a[0]+a[1]+a[2]+a[3]+a[4]+a[5]+a[6]+a[7]))
KERNEL(k_ffma, float a[8]; float b = seed * 1e-3f; float c = threadIdx.x; for (int i = 0; i < 8; i++) a[i] = seed + i, for (int j = 0; j < 8; j++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c));, (u32)(a[0]+a[1]+a[2]+a[3]+a[4]+a[5]+a[6]+a[7]))

template <typename K> void run(const char* name, K k, double ops_per_iter, u32* d_out, int sms, double mhz) {
  int blocks = sms * 4, threads = 256;
  k<<<blocks, threads>>>(d_out, 12345); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<<<blocks, threads>>>(d_out, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)blocks * threads / 32 * ITER * ops_per_iter;
  double cycles = ms * 1e-3 * mhz * 1e6;
  printf("%-14s %8.3f ms  %6.2f warp-instr/clk/SM  (%.2f per SMSP)\n", name, ms, warp_instr / cycles / sms, warp_instr / cycles / sms / 4);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1e3; printf("%s SMs=%d clock=%.0f MHz (nominal max; rates assume the GPU runs at it)\n", p.name, sms, mhz);
  u32* d; cudaMalloc(&d, sms * 4 * 256 * 4);
  run("imad.lo", k_imad_lo, 64, d, sms, mhz); run("imad.hi", k_imad_hi, 64, d, sms, mhz); run("imad.wide", k_imad_wide, 64, d, sms, mhz);
  run("iadd", k_iadd3, 64, d, sms, mhz); run("lop3", k_lop3, 64, d, sms, mhz); run("add.u64", k_add64, 64, d, sms, mhz); run("shf", k_shf, 64, d, sms, mhz);
  run("wide+1lop", k_mix_w1a1, 64, d, sms, mhz); run("wide+3lop", k_mix_w1a3, 64, d, sms, mhz); run("wide+2lop+dfma", k_mix_w1a2d1, 64, d, sms, mhz); run("iadd3 3-op", k_iadd3_3op, 64, d, sms, mhz); run("lo+1add", k_mix_l1a1, 64, d, sms, mhz);
  run("add.cc/addc", k_addcc, 64, d, sms, mhz); run("mad.cc/madc", k_madcc, 64, d, sms, mhz);
  run("dfma", k_dfma, 64, d, sms, mhz); run("ffma", k_ffma, 64, d, sms, mhz);
  return 0;
}

// Dependent-issue latencies on a lone warp (clock64 around unrolled chains): the cost model behind the latency form of the
// lane-cooperative permutation (poseidon_g_coop2.cuh).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 lat.cu
#include "../../stark-verifier_b200/csrc/fri_kernels.cuh"
#include <cstdio>
using namespace svb;

#define N 256
template <int K>
__global__ void lat_kernel(u64* out, u64 seed, u32* smem_init) {
    __shared__ u32 chase[64];
    if (threadIdx.x < 64) chase[threadIdx.x] = (threadIdx.x + 1) & 63;
    __syncthreads();
    u64 x = seed + threadIdx.x, y = seed * 3 + 1;
    u32 a = (u32)x, b = (u32)y, c = (u32)(x >> 32);
    long long t0 = clock64();
    if (K == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) x = mul(x, x);
    } else if (K == 1) {
#pragma unroll
        for (int i = 0; i < N; i++) x = (u64)(u32)x * b + x;           // IMAD.WIDE.U32, dependent through the addend and a factor
    } else if (K == 2) {
#pragma unroll
        for (int i = 0; i < N; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
        x = a;
    } else if (K == 3) {
#pragma unroll
        for (int i = 0; i < N; i++) asm volatile("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, %0;" : "+r"(a), "+r"(c) : "r"(b));
        x = ((u64)c << 32) | a;
    } else if (K == 4) {
#pragma unroll
        for (int i = 0; i < N; i++) a = __shfl_sync(0xFFFFFFFFu, a, (a & 1) ^ 1, 16);
        x = a;
    } else if (K == 5) {
#pragma unroll
        for (int i = 0; i < N; i++) a = chase[a & 63];
        x = a;
    } else if (K == 6) {
#pragma unroll 8
        for (int i = 0; i < N / 4; i++) x = sbox7(x);
    } else if (K == 7) {
#pragma unroll
        for (int i = 0; i < N; i++) x = mul_add(x, y, x);
    } else if (K == 8) {
#pragma unroll
        for (int i = 0; i < N; i++) x = (u64)(u32)x * b;                // IMAD.WIDE.U32 without addend
    } else if (K == 9) {
#pragma unroll
        for (int i = 0; i < N; i++) { u32 r0, r1, r2, r3; mulw4(x, y, r0, r1, r2, r3); x = ((u64)(r3 ^ r1) << 32) | (r2 ^ r0); }   // product only
    } else if (K == 10) {
#pragma unroll
        for (int i = 0; i < N; i++) x = red4((u32)x, (u32)(x >> 32), b, c);   // reduction only
    } else if (K == 11) {
        double d = __longlong_as_double((long long)(x & 0xFFFFFFFF)), e = 17.0;
#pragma unroll
        for (int i = 0; i < N; i++) d = __fma_rn(d, e, d);
        x = (u64)__double_as_longlong(d);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[64 + K] = (u64)(t1 - t0);
}

int main() {
    u64* d;
    cudaMalloc(&d, 1024 * 8);
    const char* names[] = {"mul (modular multiplication, LOOSE)", "IMAD.WIDE.U32 with addend", "IADD (32-bit)", "add.cc + addc pair", "SHFL.IDX (width 16)",
                           "LDS.32 pointer chase", "sbox7 / 4 (per multiplication level ~ /0.75)", "mul_add", "IMAD.WIDE.U32 no addend", "mulw4 only",
                           "red4 only", "DFMA (subnormal operand)"};
    u64 h[128];
#define RUN(K) lat_kernel<K><<<1, 32>>>(d, 0x123456789ABCDEFull, nullptr); lat_kernel<K><<<1, 32>>>(d, 0x123456789ABCDEFull, nullptr); cudaDeviceSynchronize(); \
    cudaMemcpy(h, d, 128 * 8, cudaMemcpyDeviceToHost); printf("%-50s %7.2f cycles per link\n", names[K], (double)h[64 + K] / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    return 0;
}

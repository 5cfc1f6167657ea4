// Cost of many small strided H2D copies against one large one (is a per-proof split of the last chunk's copy affordable?)
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t proof = 156812, qb = 5251, rounds = 28, a = 3452, n = 416, front = 9000;
    unsigned char *h, *d;
    cudaHostAlloc(&h, proof * n, cudaHostAllocDefault);
    cudaMalloc(&d, proof * n);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, s);
        cudaMemcpy2DAsync(d, qb * rounds + 8, h + front, proof, qb * rounds, n, cudaMemcpyHostToDevice, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("one 2D copy, %zu rows of %zu B: %.3f ms (%.1f GB/s)\n", n, qb * rounds, ms, qb * rounds * n / ms / 1e6);
        cudaEventRecord(e0, s);
        for (size_t p = 0; p < n; p++) cudaMemcpy2DAsync(d + p * (qb * rounds + 8), qb, h + front + p * proof, qb, a, rounds, cudaMemcpyHostToDevice, s);
        for (size_t p = 0; p < n; p++) cudaMemcpy2DAsync(d + p * (qb * rounds + 8) + a, qb, h + front + p * proof + a, qb, qb - a, rounds, cudaMemcpyHostToDevice, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("2 x %zu per-proof 2D copies (28 rows of %zu / %zu B): %.3f ms\n", n, a, qb - a, ms);
        cudaEventRecord(e0, s);
        for (size_t p = 0; p < n; p++) cudaMemcpyAsync(d + p * (qb * rounds + 8), h + front + p * proof, qb * rounds, cudaMemcpyHostToDevice, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("%zu per-proof 1D copies of %zu B: %.3f ms\n", n, qb * rounds, ms);
    }
    return 0;
}

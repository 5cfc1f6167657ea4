// Can SMs pull pinned host memory over PCIe as fast as the copy engine pushes it?  (candidate for a fused pull + unpack kernel)
#include <cstdio>
#include <cuda_runtime.h>
template <int U>
__global__ void pull(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; u++) dst[i + u * stride] = v[u];
    }
}
int main() {
    const size_t bytes = 642ull << 20, n = bytes / 16;
    uint4 *h, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocMapped);
    for (size_t i = 0; i < n; i += 4096) h[i].x = (unsigned)i;
    cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    cudaEventRecord(e0); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventRecord(e0); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemcpyAsync: %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    const int grids[] = {37, 74, 148, 296, 592};
    for (int g : grids) {
        for (int u = 1; u <= 8; u *= 2) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (u == 1) pull<1><<<g, 128>>>(h, d, n); else if (u == 2) pull<2><<<g, 128>>>(h, d, n); else if (u == 4) pull<4><<<g, 128>>>(h, d, n); else pull<8><<<g, 128>>>(h, d, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            cudaEventElapsedTime(&ms, e0, e1);
            printf("pull kernel %3d blocks x 128 threads, %d x 16 B in flight per thread: %.2f ms  %.1f GB/s\n", g, u, ms, bytes / ms / 1e6);
        }
    }
    cudaError_t e = cudaGetLastError(); if (e) printf("err %s\n", cudaGetErrorString(e));
    return 0;
}

// Device side of ntt.hpp: one block per tile of a pass, 512 threads, the tile in 32 KB of static shared memory.
#pragma once
#include "ntt.hpp"

namespace svb {

#define SVB_NTT_BLOCK 512

struct DevSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// grid = (blocks per polynomial, n_polys): polynomial p at data + p * 2^k
__global__ void __launch_bounds__(SVB_NTT_BLOCK) ntt_pass_kernel(u64* __restrict__ data, const u64* __restrict__ tw, NttPass P) {
    __shared__ u64 tile[1 << NTT_TILE_LOG];
    u64* poly = data + ((size_t)blockIdx.y << P.k);
    ntt_pass_block(P, poly, tw, tile, (u64)blockIdx.x, threadIdx.x, SVB_NTT_BLOCK, DevSync());
}

// x[i] *= s for every element (the 1/n of the inverse transform)
__global__ void ntt_scale_kernel(u64* __restrict__ data, size_t total, u64 s) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < total) data[i] = mulc(data[i], s);
}

// LDE prelude: out[p][j] = coeffs[p][j] * shift^j (j < n), 0 (n <= j < N)
__global__ void lde_scale_pad_kernel(const u64* __restrict__ coeffs, u64* __restrict__ out, u32 log_n, u32 log_N, size_t n_polys, u64 shift) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (n_polys << log_N)) return;
    size_t p = i >> log_N, j = i & (((size_t)1 << log_N) - 1);
    out[i] = lde_scaled_coeff(coeffs + (p << log_n), (u64)1 << log_n, shift, j);
}

// [rows][cols] -> [cols][rows] through a 32 x 33 shared tile (coalesced on both sides): LDE output (one row per
// polynomial) -> Merkle leaves (one row per evaluation point).  grid = (ceil(cols/32), ceil(rows/32)), block = (32, 8).
__global__ void __launch_bounds__(256) transpose_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t rows, size_t cols) {
    __shared__ u64 tile[32][33];
    const size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
    for (u32 j = threadIdx.y; j < 32; j += 8) {
        const size_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (u32 j = threadIdx.y; j < 32; j += 8) {
        const size_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
    }
}

}  // namespace svb

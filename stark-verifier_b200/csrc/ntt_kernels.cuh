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

}  // namespace svb

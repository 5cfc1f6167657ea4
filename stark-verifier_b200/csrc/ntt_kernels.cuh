// Device side of ntt.hpp: one block of 512 threads per tile of a pass, the padded tile (up to 72 KB) in dynamic shared
// memory, the phases of ntt.hpp separated by barriers.
#pragma once
#include "ntt.hpp"

namespace svb {

// grid.x = tiles per polynomial * items (item = polynomial, or (polynomial, coset) for an LDE); item-major
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(const u64* __restrict__ src, size_t src_stride, u64* __restrict__ dst,
                                                               size_t dst_stride, const u64* __restrict__ tw,
                                                               const u64* __restrict__ scale_lo, const u64* __restrict__ scale_hi,
                                                               NttPass P, NttRounds Q) {
    extern __shared__ u64 ntt_tile[];
    const u64 tiles = ntt_tiles_per_poly(P);
    const u64 item = blockIdx.x >> (P.k - P.logT), t = blockIdx.x & (tiles - 1);
    const NttTileMap M = ntt_tile_map(P, t);
    const u64* s = src + (item >> P.coset_bits) * src_stride;
    u64* d = dst + item * dst_stride;
    ntt_tile_load(P, M, s, scale_lo, scale_hi, (u32)(item & ((1u << P.coset_bits) - 1)), ntt_tile, threadIdx.x, NTT_THREADS);
    __syncthreads();
    for (u32 i = 0; i < Q.n; i++) {
        ntt_tile_round_any(P, M, tw, ntt_tile, Q.b[i], Q.r[i], threadIdx.x, NTT_THREADS);
        __syncthreads();
    }
    ntt_tile_store(P, M, d, ntt_tile, threadIdx.x, NTT_THREADS);
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

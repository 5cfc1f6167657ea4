// Goldilocks NTT / low-degree extension (the commit-phase library next to sv_merkle_tree_build).
//
// plonky2 commits to a polynomial by evaluating it on the coset 7 * <omega_N> and putting the value at
// 7 * omega^bitrev(i) into leaf i (the order the FRI verifier assumes: chip/fri_chip.rs:152-166,262-264).  A
// decimation-in-frequency NTT takes coefficients in natural order and leaves the evaluations in exactly that
// bit-reversed order, so LDE = scale coefficient j by shift^j, zero-pad to N, run the DIF stages -- no permutation pass.
// The inverse direction (values in bit-reversed order -> coefficients) is the mirrored decimation-in-time network with
// inverse twiddles and a final 1/n.
//
// HBM traffic is what bounds an NTT, so the log2(n) stages are not log2(n) passes: a pass keeps a tile in shared
// memory and runs up to NTT_MAX_R consecutive stages on it.  In a pass over stages [s0, s0 + R) the 2^R elements
//     idx = hi * 2^(k - s0) + m * 2^(k - s0 - R) + lo,   m = 0 .. 2^R - 1
// only talk to each other; a block takes 2^logc consecutive `lo` values (8 * 2^logc-byte rows: 64-128 B) of one `hi`.
// k = 15 (shape A's LDE) is 2 passes, k = 22 (shape B's) is 3: 32-48 bytes of traffic per element instead of 240-352.
//
// One body for host and device (ntt_pass_block): the kernel runs it with the tile in shared memory and
// __syncthreads(); the host twin (sv_ntt_host, used by the CPU tests) runs the same code with one "thread".
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"

namespace svb {

#define NTT_MAX_R 9          // stages per pass: tile of 2^R rows x 2^logc columns, <= 2^12 elements = 32 KB of shared memory
#define NTT_TILE_LOG 12

struct NttPass {
    u32 k;        // log2(n)
    u32 s0, R;    // this pass runs DIF stages s0 .. s0+R-1 (DIT: the same stages in reverse order)
    u32 logc;     // log2 of the consecutive `lo` values per block
    u32 inverse;  // 0: DIF with omega, 1: DIT with omega^-1
};

// element (m, c) of block `blk` of a pass -> index inside one polynomial
SVB_HD u64 ntt_tile_index(const NttPass& P, u64 blk, u32 m, u32 c) {
    const u32 lo_bits = P.k - P.s0 - P.R;                 // bits below the pass's R bits
    const u64 lo_blocks = 1ull << (lo_bits - P.logc);     // blocks per `hi`
    const u64 hi = blk / lo_blocks, lo = (blk % lo_blocks) << P.logc;
    return (hi << (P.k - P.s0)) + ((u64)m << lo_bits) + lo + c;
}
SVB_HD u64 ntt_blocks_per_poly(const NttPass& P) { return 1ull << (P.k - P.R - P.logc); }

// One block of one pass over one polynomial.  tile: 2^(R + logc) words; tw: omega^t (or omega^-t), t < n/2.
// DIF butterfly of stage s on (i0, i1 = i0 + half), half = n >> (s + 1):  a' = a + b,  b' = (a - b) * w^((i0 mod half) << s)
// DIT (inverse) of the same stage:                                        b'' = b * w^-(...),  a' = a + b'',  b' = a - b''
template <class Sync>
SVB_HD void ntt_pass_block(const NttPass& P, u64* __restrict__ data, const u64* __restrict__ tw, u64* tile, u64 blk, u32 tid,
                           u32 nthreads, Sync sync) {
    const u32 rows = 1u << P.R, cols = 1u << P.logc, elems = rows * cols;
    for (u32 e = tid; e < elems; e += nthreads) tile[e] = data[ntt_tile_index(P, blk, e >> P.logc, e & (cols - 1))];
    sync();
    const u32 lo_bits = P.k - P.s0 - P.R;
    for (u32 step = 0; step < P.R; step++) {
        const u32 s = P.inverse ? P.s0 + P.R - 1 - step : P.s0 + step;   // global stage
        const u32 bit = P.s0 + P.R - 1 - s;                              // which bit of m the stage pairs on
        const u64 half = 1ull << (P.k - 1 - s);
        for (u32 b = tid; b < elems / 2; b += nthreads) {
            const u32 c = b & (cols - 1), mm = b >> P.logc;              // mm: m with the paired bit removed
            const u32 m0 = ((mm >> bit) << (bit + 1)) | (mm & ((1u << bit) - 1)), m1 = m0 | (1u << bit);
            const u64 i0 = ntt_tile_index(P, blk, m0, c);
            const u64 w = tw[(i0 & (half - 1)) << s];
            u64 x = tile[(m0 << P.logc) + c], y = tile[(m1 << P.logc) + c];
            if (P.inverse) {
                y = mulc(y, w);
                tile[(m0 << P.logc) + c] = add(x, y);
                tile[(m1 << P.logc) + c] = sub(x, y);
            } else {
                tile[(m0 << P.logc) + c] = add(x, y);
                tile[(m1 << P.logc) + c] = mulc(sub(x, y), w);
            }
        }
        sync();
    }
    (void)lo_bits;
    for (u32 e = tid; e < elems; e += nthreads) data[ntt_tile_index(P, blk, e >> P.logc, e & (cols - 1))] = tile[e];
}

// The passes of a size-2^k transform: stages split into ceil(k / NTT_MAX_R) runs of nearly equal length; a pass whose
// `lo` part is wide enough takes 16 consecutive columns per block (128-byte rows), the last pass (lo_bits = 0) one.
static inline int ntt_plan(u32 k, bool inverse, NttPass out[8]) {
    if (k == 0 || k > 32) return -1;
    const u32 n_pass = (k + NTT_MAX_R - 1) / NTT_MAX_R;
    u32 s0 = 0;
    for (u32 p = 0; p < n_pass; p++) {
        const u32 R = (k - s0 + (n_pass - p) - 1) / (n_pass - p);
        const u32 lo_bits = k - s0 - R;
        u32 logc = lo_bits < 4 ? lo_bits : 4;
        if (R + logc > NTT_TILE_LOG) logc = NTT_TILE_LOG - R;
        NttPass q = {k, s0, R, logc, inverse ? 1u : 0u};
        out[inverse ? n_pass - 1 - p : p] = q;     // the inverse network runs the stages, hence the passes, backwards
        s0 += R;
    }
    return (int)n_pass;
}

// twiddles omega_n^t (inverse: omega_n^-t), t < n/2, omega_n = 7^((p-1)/n)
static inline void ntt_twiddles(u32 k, bool inverse, u64* out) {
    u64 w = pow(7, (GL_P - 1) >> k);
    if (inverse) w = inv(w);
    u64 cur = 1;
    for (u64 t = 0; t < (1ull << k) / 2; t++) { out[t] = cur; cur = mulc(cur, w); }
    if (k == 0) out[0] = 1;
}

// LDE prelude: out[j] = coeffs[j] * shift^j for j < n, 0 for n <= j < N
SVB_HD u64 lde_scaled_coeff(const u64* __restrict__ coeffs, u64 n, u64 shift, u64 j) { return j < n ? mulc(coeffs[j], pow(shift, j)) : 0; }

}  // namespace svb

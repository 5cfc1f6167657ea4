// Goldilocks NTT / low-degree extension (the commit-phase library next to sv_merkle_tree_build).
//
// plonky2 commits to a polynomial by evaluating it on the coset 7 * <omega_N> and putting the value at
// 7 * omega^bitrev(i) into leaf i (the order the FRI verifier assumes: chip/fri_chip.rs:152-166,262-264).  A
// decimation-in-frequency NTT takes coefficients in natural order and leaves the evaluations in exactly that
// bit-reversed order, so no permutation pass is ever needed.  The inverse direction (values in bit-reversed order ->
// coefficients) is the mirrored decimation-in-time network with inverse twiddles and a final 1/n.
//
// Round 2 design (round 1: radix-2 stages through shared memory, one butterfly per thread and stage, 64-bit divisions in
// the index math, 128-element tiles in the last pass -- 9 % of the HBM roofline, issue-bound; profiles/ncu_r2head_*):
//  * a PASS keeps a tile of 2^13 elements (64 KB) in shared memory and runs up to 12 consecutive stages on it, so a
//    transform is 1 pass up to 2^13 points, 2 up to 2^22 (shape B's LDE), 3 beyond -- 16 bytes of HBM traffic per
//    element and pass.  Non-final passes take 2^R rows x >= 8 consecutive columns (64-byte runs); the final pass takes a
//    contiguous 64 KB chunk (whole sub-transforms of 2^12 points);
//  * inside a pass the stages run as ROUNDS of 3 (radix 8) in registers: a thread loads the 8 elements of a group from
//    shared memory, does the three butterfly levels with the fixed 8th roots of unity, multiplies output q by w1^q --
//    ONE twiddle-table entry w1 = omega^(u << s) per group, its powers by 6 multiplications -- and writes them back:
//    one shared-memory round trip and one barrier per 3 stages, all index math in shifts and masks;
//  * the tile is padded (word e lives at e + e/8) so that the strided group accesses of every round are conflict-free;
//  * an LDE never materialises the zero-padded input: with N = n * 2^rate_bits the first rate_bits stages of the size-N
//    transform only replicate the coefficients, block c receiving a_j * (shift * omega_N^bitrev(c))^j.  So
//    LDE = 2^rate_bits independent size-n NTTs whose first pass scales on load (two-level power table, both levels small);
//    shape A's 2^12 -> 2^15 is ONE pass: read 4 KB x 8 (L2 hits), write 256 KB per polynomial.
//
// One body for host and device: the tile routine below is phase-structured (load / round / store), every phase a loop
// over "threads"; the kernel runs the phases with __syncthreads() between them, the host twin (sv_ntt_host /
// sv_lde_host, what the CPU tests check against naive big-integer evaluation) runs each phase for tid = 0 .. T-1.
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"
#include <string.h>

namespace svb {

#define NTT_TL 13                       // log2 of the tile: 2^13 elements = 64 KB
#define NTT_THREADS 512
#define NTT_SMEM_WORDS ((1u << NTT_TL) + (1u << (NTT_TL - 3)))   // padded tile: word e at e + (e >> 3)
#define NTT_MAX_PASSES 4

struct NttPass {
    u32 k;           // log2 of the transform size n
    u32 s0, R;       // this pass runs DIF stages s0 .. s0+R-1 (inverse: the same stages of the DIT network, backwards)
    u32 logT;        // log2 of the tile = min(NTT_TL, k)
    u32 inverse;     // 0: DIF with omega, 1: DIT with omega^-1
    u32 coset_bits;  // LDE: polynomial index p of the grid = (source polynomial p >> coset_bits, coset p & mask); else 0
    u32 load_scaled; // LDE first pass: element j of the source is multiplied by scale(coset, j) on load
    u32 scale_h;     // scale(c, j) = lo[c][j & (2^h - 1)] * hi[c][j >> h]   (hi is skipped when h == k)
    u64 out_scale;   // multiply on store when != 1 (the 1/n of the inverse transform, in its last pass)
};

SVB_HD u32 ntt_pad(u32 e) { return e + (e >> 3); }
SVB_HD u64 ntt_tiles_per_poly(const NttPass& P) { return 1ull << (P.k - P.logT); }

// index inside the polynomial of tile element e.  Non-final pass: 2^R rows (bits [logC, logT) of e) x 2^logC columns;
// final pass (no bits below the pass's R bits): contiguous.
struct NttTileMap {
    u64 base;
    u32 lo_bits, mb;
    SVB_HD u64 operator()(u32 e) const { return base + (e & ((1u << mb) - 1)) + ((u64)(e >> mb) << lo_bits); }
};
SVB_HD NttTileMap ntt_tile_map(const NttPass& P, u64 tile) {
    NttTileMap M;
    M.lo_bits = P.k - P.s0 - P.R;
    const u32 logC = P.logT - P.R;
    if (M.lo_bits == 0) {
        M.mb = 0;
        M.base = tile << P.logT;
    } else {
        M.mb = logC;
        const u32 lb = M.lo_bits - logC;                       // log2 of the column blocks per `hi`
        M.base = ((tile >> lb) << (P.k - P.s0)) + ((tile & ((1ull << lb) - 1)) << logC);
    }
    return M;
}

// the fixed roots of the radix-8 / radix-4 butterflies are powers of two: omega_8 = 7^((p-1)/8) = -2^24, omega_4 = omega_8^2 = 2^48.
// x * 2^(24 j) mod p, canonical, j = 1..3: the magnitude of omega_8^j (omega_8^j = (-1)^j 2^(24 j); inverse: omega_8^-j =
// (-1)^(j+1) 2^(24 (4 - j))) -- the callers fold the sign into the subtraction in front.  On the device the product is three
// funnel shifts into the limbs of red5 (1 IMAD.WIDE + 6 ALU) instead of a full modular multiplication (5 IMAD.WIDE + 9 ALU).
template <int J>
SVB_HD u64 ntt_mul_pow24(u64 x) {
#if defined(__CUDA_ARCH__)
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    if (J == 1) return canon(red5(lo << 24, __funnelshift_l(lo, hi, 24), hi >> 8, 0, 0));
    if (J == 2) return canon(red5(0, lo << 16, __funnelshift_l(lo, hi, 16), hi >> 16, 0));
    return canon(red5(0, 0, lo << 8, __funnelshift_l(lo, hi, 8), hi >> 24));
#else
    return mulc(x, J == 1 ? (1ull << 24) : J == 2 ? (1ull << 48) : 0x000000ffffffff00ull /* 2^72 mod p */);
#endif
}

// ---- phases of one tile ------------------------------------------------------------------------------------
// load: global -> padded tile (LDE first pass: scaled on the way)
SVB_HD void ntt_tile_load(const NttPass& P, const NttTileMap& M, const u64* __restrict__ src, const u64* __restrict__ scale_lo,
                          const u64* __restrict__ scale_hi, u32 coset, u64* tile, u32 tid, u32 nthreads) {
    const u32 T = 1u << P.logT;
    for (u32 e = tid; e < T; e += nthreads) {
        const u64 j = M(e);
        u64 v = src[j];
        if (P.load_scaled) {
            v = mul(v, scale_lo[((u64)coset << P.scale_h) + (j & ((1ull << P.scale_h) - 1))]);
            if (P.scale_h < P.k) v = mul(v, scale_hi[((u64)coset << (P.k - P.scale_h)) + (j >> P.scale_h)]);
            v = canon(v);
        }
        tile[ntt_pad(e)] = v;
    }
}
SVB_HD void ntt_tile_store(const NttPass& P, const NttTileMap& M, u64* __restrict__ dst, const u64* tile, u32 tid, u32 nthreads) {
    const u32 T = 1u << P.logT;
    for (u32 e = tid; e < T; e += nthreads) {
        u64 v = tile[ntt_pad(e)];
        if (P.out_scale != 1) v = mulc(v, P.out_scale);
        dst[M(e)] = v;
    }
}

// One round: r <= 3 stages on the bits [b, b + r) of the pass's R bits.  tw: omega_n^t (inverse: omega_n^-t), t < n/2.
template <int r>
SVB_HD void ntt_tile_round(const NttPass& P, const NttTileMap& M, const u64* __restrict__ tw, u64* tile, u32 b, u32 tid, u32 nthreads) {
    const u32 T = 1u << P.logT, pos = M.mb + b;
    const u32 s_a = P.s0 + (P.R - b - r);                      // first global stage of the round
    const u64 umask = (1ull << (P.k - s_a - r)) - 1;           // indices below the round's bits: u = index mod (N_A / 2^r)
    const bool inv_ = P.inverse != 0;
    u64 u_prev = ~0ull, w1 = 1, w2 = 1, w3 = 1, w4 = 1, w5 = 1, w6 = 1, w7 = 1;   // the groups of one thread often share u (other hi-block / coset, same position)
    for (u32 g = tid; g < (T >> r); g += nthreads) {
        const u32 e0 = ((g >> pos) << (pos + r)) | (g & ((1u << pos) - 1));
        const u64 u = M(e0) & umask;
        if (u != u_prev) {
            u_prev = u;
            w1 = tw[u << s_a];                                 // omega_(N_A)^u; its powers stay LOOSE (they only feed multiplications)
            if (r >= 2) { w2 = mul(w1, w1); w3 = mul(w2, w1); }
            if (r == 3) { w4 = mul(w2, w2); w5 = mul(w4, w1); w6 = mul(w3, w3); w7 = mul(w6, w1); }
        }
        u64 x[1 << r];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < (1 << r); j++) x[j] = tile[ntt_pad(e0 | ((u32)j << pos))];
        if (r == 3) {
            if (!inv_) {
                // DIF: three butterfly levels, then output p (frequency bitrev(p)) times w1^bitrev(p)
                u64 a0 = add(x[0], x[4]), a4 = sub(x[0], x[4]);
                // omega_8 = -2^24, omega_8^2 = 2^48, omega_8^3 = -2^72: shifts, the sign in the order of the subtraction
                u64 a1 = add(x[1], x[5]), a5 = ntt_mul_pow24<1>(sub(x[5], x[1]));
                u64 a2 = add(x[2], x[6]), a6 = ntt_mul_pow24<2>(sub(x[2], x[6]));
                u64 a3 = add(x[3], x[7]), a7 = ntt_mul_pow24<3>(sub(x[7], x[3]));
                u64 b0 = add(a0, a2), b2 = sub(a0, a2), b1 = add(a1, a3), b3 = ntt_mul_pow24<2>(sub(a1, a3));
                u64 b4 = add(a4, a6), b6 = sub(a4, a6), b5 = add(a5, a7), b7 = ntt_mul_pow24<2>(sub(a5, a7));
                x[0] = add(b0, b1);
                x[1] = mulc(sub(b0, b1), w4);
                x[2] = mulc(add(b2, b3), w2);
                x[3] = mulc(sub(b2, b3), w6);
                x[4] = mulc(add(b4, b5), w1);
                x[5] = mulc(sub(b4, b5), w5);
                x[6] = mulc(add(b6, b7), w3);
                x[7] = mulc(sub(b6, b7), w7);
            } else {
                // DIT: input p times w1^bitrev(p) (inverse table), then the three levels from the bottom
                u64 y1 = mulc(x[1], w4), y2 = mulc(x[2], w2), y3 = mulc(x[3], w6), y4 = mulc(x[4], w1), y5 = mulc(x[5], w5),
                    y6 = mulc(x[6], w3), y7 = mulc(x[7], w7);
                // omega_8^-1 = 2^72, omega_8^-2 = -2^48, omega_8^-3 = 2^24
                u64 b0 = add(x[0], y1), b1 = sub(x[0], y1), b2 = add(y2, y3), b3 = ntt_mul_pow24<2>(sub(y3, y2));
                u64 b4 = add(y4, y5), b5 = sub(y4, y5), b6 = add(y6, y7), b7 = ntt_mul_pow24<2>(sub(y7, y6));
                u64 a0 = add(b0, b2), a2 = sub(b0, b2), a1 = add(b1, b3), a3 = sub(b1, b3);
                u64 a4 = add(b4, b6), a6 = sub(b4, b6), a5 = add(b5, b7), a7 = sub(b5, b7);
                a5 = ntt_mul_pow24<3>(a5);
                a6 = ntt_mul_pow24<2>(a6);                         // magnitude of omega_8^-2 = -2^48: the sign swaps add and sub below
                a7 = ntt_mul_pow24<1>(a7);
                x[0] = add(a0, a4); x[4] = sub(a0, a4);
                x[1] = add(a1, a5); x[5] = sub(a1, a5);
                x[2] = sub(a2, a6); x[6] = add(a2, a6);
                x[3] = add(a3, a7); x[7] = sub(a3, a7);
            }
        } else if (r == 2) {
            if (!inv_) {
                u64 a0 = add(x[0], x[2]), a2 = sub(x[0], x[2]), a1 = add(x[1], x[3]), a3 = ntt_mul_pow24<2>(sub(x[1], x[3]));
                x[0] = add(a0, a1);
                x[1] = mulc(sub(a0, a1), w2);
                x[2] = mulc(add(a2, a3), w1);
                x[3] = mulc(sub(a2, a3), w3);
            } else {
                u64 y1 = mulc(x[1], w2), y2 = mulc(x[2], w1), y3 = mulc(x[3], w3);
                u64 a0 = add(x[0], y1), a1 = sub(x[0], y1), a2 = add(y2, y3), a3 = ntt_mul_pow24<2>(sub(y3, y2));
                x[0] = add(a0, a2); x[2] = sub(a0, a2);
                x[1] = add(a1, a3); x[3] = sub(a1, a3);
            }
        } else {
            if (!inv_) {
                u64 a0 = add(x[0], x[1]), a1 = mulc(sub(x[0], x[1]), w1);
                x[0] = a0; x[1] = a1;
            } else {
                u64 y1 = mulc(x[1], w1);
                u64 a0 = add(x[0], y1), a1 = sub(x[0], y1);
                x[0] = a0; x[1] = a1;
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < (1 << r); j++) tile[ntt_pad(e0 | ((u32)j << pos))] = x[j];
    }
}

// the rounds of a pass in execution order: (b, r) pairs; DIF walks the R bits from the top, the inverse from the bottom
struct NttRounds {
    u32 n;
    u32 b[5], r[5];
};
SVB_HD NttRounds ntt_rounds(const NttPass& P) {
    NttRounds Q;
    Q.n = 0;
    u32 top = P.R;
    while (top > 0) {
        const u32 r = top >= 3 ? 3 : top;
        Q.b[Q.n] = top - r;
        Q.r[Q.n] = r;
        Q.n++;
        top -= r;
    }
    if (P.inverse)
        for (u32 i = 0; i < Q.n / 2; i++) {
            u32 t = Q.b[i]; Q.b[i] = Q.b[Q.n - 1 - i]; Q.b[Q.n - 1 - i] = t;
            t = Q.r[i]; Q.r[i] = Q.r[Q.n - 1 - i]; Q.r[Q.n - 1 - i] = t;
        }
    return Q;
}
SVB_HD void ntt_tile_round_any(const NttPass& P, const NttTileMap& M, const u64* __restrict__ tw, u64* tile, u32 b, u32 r, u32 tid,
                               u32 nthreads) {
    if (r == 3) ntt_tile_round<3>(P, M, tw, tile, b, tid, nthreads);
    else if (r == 2) ntt_tile_round<2>(P, M, tw, tile, b, tid, nthreads);
    else ntt_tile_round<1>(P, M, tw, tile, b, tid, nthreads);
}

// The passes of a size-2^k transform.  k <= 13: one; else a final pass of 12 stages (two whole sub-transforms per tile)
// preceded by passes of at most 10 stages (>= 8 columns = 64-byte runs).  The inverse network runs them backwards.
static inline int ntt_plan(u32 k, bool inverse, NttPass out[NTT_MAX_PASSES]) {
    if (k == 0 || k > 32) return -1;
    NttPass q;
    memset(&q, 0, sizeof q);
    q.k = k;
    q.logT = k < NTT_TL ? k : NTT_TL;
    q.inverse = inverse ? 1u : 0u;
    q.out_scale = 1;
    int np = 0;
    if (k <= NTT_TL) {
        q.s0 = 0; q.R = k;
        out[np++] = q;
    } else {
        const u32 last = NTT_TL - 1, rest = k - last, per = NTT_TL - 3;
        const u32 n_first = (rest + per - 1) / per;
        u32 s0 = 0;
        for (u32 p = 0; p < n_first; p++) {
            q.s0 = s0;
            q.R = (rest - s0 + (n_first - p) - 1) / (n_first - p);
            out[np++] = q;
            s0 += q.R;
        }
        q.s0 = s0; q.R = last;
        out[np++] = q;
    }
    if (inverse) {
        for (int i = 0; i < np / 2; i++) { NttPass t = out[i]; out[i] = out[np - 1 - i]; out[np - 1 - i] = t; }
        u64 n_mod = ((u64)1 << k) % GL_P;
        out[np - 1].out_scale = inv(n_mod);
    }
    return np;
}

// twiddles omega_n^t (inverse: omega_n^-t), t < n/2, omega_n = 7^((p-1)/n)
static inline void ntt_twiddles(u32 k, bool inverse, u64* out) {
    u64 w = pow(7, (GL_P - 1) >> k);
    if (inverse) w = inv(w);
    u64 cur = 1;
    for (u64 t = 0; t < (1ull << k) / 2; t++) { out[t] = cur; cur = mulc(cur, w); }
    if (k == 0) out[0] = 1;
}

// LDE scale tables for n = 2^log_n coefficients onto shift * <omega_N>, N = n * 2^rate_bits: coset c (the block at position
// c of the bit-reversed output) is the size-n transform of a_j * s_c^j with s_c = shift * omega_N^bitrev(c).
// lo[c][j] = s_c^j (j < 2^h), hi[c][j] = (s_c^(2^h))^j (j < 2^(log_n - h)).
static inline u32 lde_scale_h(u32 log_n) { return log_n <= 16 ? log_n : (log_n + 1) / 2; }
static inline void lde_scale_tables(u32 log_n, u32 rate_bits, u64 shift, u64* lo, u64* hi) {
    const u32 h = lde_scale_h(log_n), log_N = log_n + rate_bits;
    const u64 wN = pow(7, (GL_P - 1) >> log_N);
    for (u32 c = 0; c < (1u << rate_bits); c++) {
        u32 rc = 0;
        for (u32 i = 0; i < rate_bits; i++) rc |= ((c >> i) & 1u) << (rate_bits - 1 - i);
        const u64 sc = mulc(shift, pow(wN, rc));
        u64 cur = 1;
        for (u64 j = 0; j < (1ull << h); j++) { lo[((u64)c << h) + j] = cur; cur = mulc(cur, sc); }
        const u64 sh = cur;                                     // s_c^(2^h)
        cur = 1;
        for (u64 j = 0; j < (1ull << (log_n - h)); j++) { hi[((u64)c << (log_n - h)) + j] = cur; cur = mulc(cur, sh); }
    }
}

}  // namespace svb

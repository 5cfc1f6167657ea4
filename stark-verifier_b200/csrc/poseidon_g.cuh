// Poseidon permutation over Goldilocks (hash family "G"): width 12, S-box x^7, 4 + 22 + 4 rounds.
//
// Replaces (semantics): plonky2's PoseidonPermutation as restated by the reference at
//   chip/plonk/gates/poseidon.rs:634-686 (fast form: full rounds, FAST_PARTIAL_FIRST_ROUND_CONSTANT,
//   mds_partial_layer_init :504-537, 22 x { x^7 on lane 0, + FAST_PARTIAL_ROUND_CONSTANTS[r],
//   mds_partial_layer_fast :539-589 }, full rounds), MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20)
//   + diag(8,0,...) (:321-322, row formula :450-479).
//
// One thread owns one permutation; the 12-word state lives in registers (24 x 32-bit).  Why not the
// 12-lane cooperative layout: the 22 partial rounds are serial in lane 0, so a lane-per-word layout
// idles 11/12 lanes for ~40% of the instruction stream; thread-per-permutation keeps every lane busy
// and needs no shuffles.  The inside of the permutation works on LOOSE u64 values (see
// goldilocks.cuh); sums of products are accumulated unreduced and reduced once per output word.
#pragma once
#include "goldilocks.cuh"

namespace svb {

// ---- tables -------------------------------------------------------------------------------------
#define SVB_TABLE(name, n) static const uint64_t h_##name[n]
#include "poseidon_g_constants.inc"
#undef SVB_TABLE
#if defined(__CUDACC__)
#define SVB_TABLE(name, n) __constant__ uint64_t d_##name[n]
#include "poseidon_g_constants.inc"
#undef SVB_TABLE
#endif

#if defined(__CUDA_ARCH__)
#define SVB_T(name) d_##name
#else
#define SVB_T(name) h_##name
#endif

// Coefficient of state[j] in output row r of the MDS layer: CIRC[(j - r) mod 12] (+ DIAG[r] if j == r).
// Small compile-time integers (<= 41): become IMAD immediates on the device.
SVB_HD constexpr u32 mds_coeff(int r, int j) {
    constexpr u32 circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    return circ[(j - r + 12) % 12] + ((r == j && r == 0) ? 8u : 0u);
}

SVB_HD u64 sbox7(u64 x) {
    u64 x2 = sqr(x);
    u64 x4 = sqr(x2);
    u64 x3 = mul(x, x2);
    return mul(x3, x4);
}

// Dense MDS layer.  Split every word into 32-bit halves, accumulate sum_j m_rj * half_j in a u64
// (< 2^41, no carries), recombine lo + hi*2^32 and reduce once.
SVB_HD void mds_layer(u64 s[12]) {
    u32 l[12], h[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        l[j] = (u32)s[j];
        h[j] = (u32)(s[j] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            al += (u64)mds_coeff(r, j) * l[j];
            ah += (u64)mds_coeff(r, j) * h[j];
        }
        u64 t = al + (ah << 32);
        u32 top = (u32)(ah >> 32) + (t < al ? 1u : 0u);
        s[r] = reduce96(t, top);
    }
}

// 160-bit accumulator for sums of 64x64 products.
struct acc160 {
    u64 lo, hi;
    u32 top;
};
SVB_HD void acc_mul(acc160& a, u64 x, u64 y) {
    u64 l, h;
    mul_wide(x, y, l, h);
#if defined(__CUDA_ARCH__)
    asm("add.cc.u64 %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+l"(a.lo), "+l"(a.hi), "+r"(a.top)
        : "l"(l), "l"(h));
#else
    unsigned __int128 s = ((unsigned __int128)a.hi << 64 | a.lo);
    unsigned __int128 p = ((unsigned __int128)h << 64 | l);
    unsigned __int128 t = s + p;
    a.top += (t < s);
    a.lo = (u64)t;
    a.hi = (u64)(t >> 64);
#endif
}
// lo + hi*2^64 + top*2^128, with 2^128 = -2^32 (mod p); top < 2^31.  LOOSE result.
SVB_HD u64 acc_reduce(const acc160& a) {
    u64 r = reduce128(a.lo, a.hi);
    u64 sub = (u64)a.top << 32;   // canonical (< 2^63)
    u64 d = r - sub;
    return r < sub ? d - GL_EPS : d;   // wrapped value is >= 2^64 - 2^63, so - EPS cannot wrap again
}

SVB_HD void poseidon_g(u64 s[12]) {
    // first half of the full rounds (poseidon.rs:637-650)
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(add_lc(s[i], SVB_T(ALL_ROUND_CONSTANTS)[12 * r + i]));
        mds_layer(s);
    }
    // partial_first_constant_layer + mds_partial_layer_init (:652-653, :504-537)
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], SVB_T(FAST_PARTIAL_FIRST_ROUND_CONSTANT)[i]);
    {
        u64 t[12];
        t[0] = s[0];
#pragma unroll
        for (int c = 1; c < 12; c++) {
            acc160 a = {0, 0, 0};
#pragma unroll
            for (int r = 1; r < 12; r++) acc_mul(a, SVB_T(FAST_PARTIAL_ROUND_INITIAL_MATRIX)[(r - 1) * 11 + (c - 1)], s[r]);
            t[c] = acc_reduce(a);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = t[i];
    }
    // 22 partial rounds (:654-672)
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        u64 s0 = sbox7(s[0]);
        s0 = add_lc(s0, SVB_T(FAST_PARTIAL_ROUND_CONSTANTS)[r]);   // entry 21 is 0 (:140), same as skipping it
        // mds_partial_layer_fast (:539-589): d = (CIRC[0]+DIAG[0]) * s0 + sum_i W_HAT[r][i-1] * s[i]
        acc160 a = {0, 0, 0};
        acc_mul(a, s0, 25);
#pragma unroll
        for (int i = 1; i < 12; i++) acc_mul(a, SVB_T(FAST_PARTIAL_ROUND_W_HATS)[r * 11 + i - 1], s[i]);
#pragma unroll
        for (int i = 1; i < 12; i++) s[i] = mul_add(SVB_T(FAST_PARTIAL_ROUND_VS)[r * 11 + i - 1], s0, s[i]);
        s[0] = acc_reduce(a);
    }
    // second half of the full rounds (:675-686), round constants 26..29
#pragma unroll 1
    for (int r = 26; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(add_lc(s[i], SVB_T(ALL_ROUND_CONSTANTS)[12 * r + i]));
        mds_layer(s);
    }
}

// canonical in, canonical out
SVB_HD void poseidon_g_canonical(u64 s[12]) {
    poseidon_g(s);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}

}  // namespace svb

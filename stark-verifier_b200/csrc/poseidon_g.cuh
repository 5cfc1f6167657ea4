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

// ================================================================================================
// Portable path: plain, readable (host: transcript and synthetic prover).
// ================================================================================================
SVB_HD void mds_layer_ref(u64 s[12]) {
    u64 out[12];
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
        for (int j = 0; j < 12; j++) {
            al += (u64)mds_coeff(r, j) * (u32)s[j];
            ah += (u64)mds_coeff(r, j) * (u32)(s[j] >> 32);
        }
        u64 t = al + (ah << 32);
        u32 top = (u32)(ah >> 32) + (t < al ? 1u : 0u);
        out[r] = reduce96(t, top);
    }
    for (int r = 0; r < 12; r++) s[r] = out[r];
}

// 160-bit accumulator for sums of 64x64 products (host + generic code).
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
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(add_lc(s[i], SVB_T(ALL_ROUND_CONSTANTS)[12 * r + i]));
        mds_layer_ref(s);
    }
    for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], SVB_T(FAST_PARTIAL_FIRST_ROUND_CONSTANT)[i]);
    {
        u64 t[12];
        t[0] = s[0];
        for (int c = 1; c < 12; c++) {
            acc160 a = {0, 0, 0};
            for (int r = 1; r < 12; r++) acc_mul(a, SVB_T(FAST_PARTIAL_ROUND_INITIAL_MATRIX)[(r - 1) * 11 + (c - 1)], s[r]);
            t[c] = acc_reduce(a);
        }
        for (int i = 0; i < 12; i++) s[i] = t[i];
    }
    for (int r = 0; r < 22; r++) {
        u64 s0 = sbox7(s[0]);
        s0 = add_lc(s0, SVB_T(FAST_PARTIAL_ROUND_CONSTANTS)[r]);   // entry 21 is 0 (:140), same as skipping it
        acc160 a = {0, 0, 0};
        acc_mul(a, s0, 25);
        for (int i = 1; i < 12; i++) acc_mul(a, SVB_T(FAST_PARTIAL_ROUND_W_HATS)[r * 11 + i - 1], s[i]);
        for (int i = 1; i < 12; i++) s[i] = mul_add(SVB_T(FAST_PARTIAL_ROUND_VS)[r * 11 + i - 1], s0, s[i]);
        s[0] = acc_reduce(a);
    }
    for (int r = 26; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(add_lc(s[i], SVB_T(ALL_ROUND_CONSTANTS)[12 * r + i]));
        mds_layer_ref(s);
    }
}

#if defined(__CUDACC__)
// ================================================================================================
// Device path (sm_100a).  Design constraints measured with ncu on B200 (profiles/):
//  * the first version unrolled everything (90 KB of SASS per permutation) and spent half of its
//    issue slots stalled on `no_instruction` (instruction-cache misses, icc hit rate 67%).  The code
//    below keeps the hot loop bodies small: ONE full-round body (S-box on 4 lanes x 3 rotations +
//    one unrolled MDS) shared by both halves, a looped initial matrix, one partial-round body;
//  * IMAD (fma pipe) and IADD3/LOP3 (alu pipe) each issue at half rate per scheduler, so the
//    arithmetic is written as mad.wide / carry chains that ptxas maps to IMAD.WIDE.U32[.X] with
//    predicate carries, keeping the two pipes roughly balanced.
// ================================================================================================
SVB_D u64 mad_wide(u32 a, u32 b, u64 c) {
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}

// out_r = sum_j m_rj * s_j + rc_r, for all 12 rows; rc (canonical) is folded into the accumulators.
SVB_D void mds_layer_rc(u64 s[12], const u64* __restrict__ rc) {
    u32 l[12], h[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        l[j] = (u32)s[j];
        h[j] = (u32)(s[j] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 c = rc[r];
        u32 al0 = (u32)c, al1 = 0, ah0 = (u32)(c >> 32), ah1 = 0;
        // carry-chain form: ptxas keeps each pair as ONE accumulating IMAD.WIDE.U32 (a plain mad.wide
        // gets re-associated into IMAD.WIDE + IADD3 + IADD3.X, one extra issue slot per product)
#pragma unroll
        for (int j = 0; j < 12; j++) {
            asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(al0), "+r"(al1) : "r"(l[j]), "r"(mds_coeff(r, j)));
            asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(ah0), "+r"(ah1) : "r"(h[j]), "r"(mds_coeff(r, j)));
        }
        u64 al = ((u64)al1 << 32) | al0, ah = ((u64)ah1 << 32) | ah0;
        // al + ah*2^32, al, ah < 2^42
        u32 t0 = (u32)al, t1, top, r0, r1;
        (void)t1; (void)top;
        asm("{\n\t"
            ".reg .u32 cy;\n\t"
            "add.cc.u32 %2, %4, %5;\n\t"
            "addc.u32 %3, %6, 0;\n\t"
            "mad.lo.cc.u32 %0, %3, 0xFFFFFFFF, %7;\n\t"
            "madc.hi.cc.u32 %1, %3, 0xFFFFFFFF, %2;\n\t"
            "addc.u32 cy, 0, 0;\n\t"
            "sub.u32 cy, 0, cy;\n\t"
            "add.cc.u32 %0, %0, cy;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1), "=&r"(t1), "=&r"(top)
            : "r"((u32)(al >> 32)), "r"((u32)ah), "r"((u32)(ah >> 32)), "r"(t0));
        s[r] = ((u64)r1 << 32) | r0;
    }
}

// The same layer on the FP64 pipe.  Measured on B200 (tools/microbench/pipes.cu): IMAD.WIDE.U32 issues
// at 1/4 rate per scheduler (4 cycles per warp instruction on the fmaheavy pipe, which the S-boxes
// already saturate), DFMA at 1/2 rate on the otherwise idle FP64 pipe.  Every product here is
// (coefficient <= 49) x (32-bit half) and every row sum is < 2^42, so the arithmetic is EXACT in
// binary64: the accumulator starts at 2^52 + rc_half and stays inside [2^52, 2^53), where doubles are
// consecutive integers, and the integer is read back from the mantissa bits.
SVB_D void mds_layer_rc_f64(u64 s[12], const u64* __restrict__ rc) {
    const double TWO52 = 4503599627370496.0;
    double d[12];
    u32 h[12];
    u64 al[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        d[j] = __hiloint2double(0x43300000, (int)(u32)s[j]) - TWO52;
        h[j] = (u32)(s[j] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double acc = __hiloint2double(0x43300000, (int)(u32)rc[r]);
#pragma unroll
        for (int j = 0; j < 12; j++) acc = __fma_rn(d[j], (double)mds_coeff(r, j), acc);
        al[r] = (u64)__double_as_longlong(acc) & 0xFFFFFFFFFFFFFull;
    }
#pragma unroll
    for (int j = 0; j < 12; j++) d[j] = __hiloint2double(0x43300000, (int)h[j]) - TWO52;
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double acc = __hiloint2double(0x43300000, (int)(u32)(rc[r] >> 32));
#pragma unroll
        for (int j = 0; j < 12; j++) acc = __fma_rn(d[j], (double)mds_coeff(r, j), acc);
        u64 ah = (u64)__double_as_longlong(acc) & 0xFFFFFFFFFFFFFull;
        // al + ah*2^32, al, ah < 2^42
        u32 t0 = (u32)al[r], t1, top, r0, r1;
        (void)t1; (void)top;
        asm("{\n\t"
            ".reg .u32 cy;\n\t"
            "add.cc.u32 %2, %4, %5;\n\t"
            "addc.u32 %3, %6, 0;\n\t"
            "mad.lo.cc.u32 %0, %3, 0xFFFFFFFF, %7;\n\t"
            "madc.hi.cc.u32 %1, %3, 0xFFFFFFFF, %2;\n\t"
            "addc.u32 cy, 0, 0;\n\t"
            "sub.u32 cy, 0, cy;\n\t"
            "add.cc.u32 %0, %0, cy;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "}"
            : "=&r"(r0), "=&r"(r1), "=&r"(t1), "=&r"(top)
            : "r"((u32)(al[r] >> 32)), "r"((u32)ah), "r"((u32)(ah >> 32)), "r"(t0));
        s[r] = ((u64)r1 << 32) | r0;
    }
}

// Unreduced sum of 64x64 products: even columns (a0*b0, a1*b1) in e0..e4, odd columns
// (a0*b1 + a1*b0, weight 2^32) in o0..o2.  Up to 2^4 terms.
struct dot_acc {
    u32 e0, e1, e2, e3, e4, o0, o1, o2;
};
SVB_D void dot_init(dot_acc& a) { a.e0 = a.e1 = a.e2 = a.e3 = a.e4 = a.o0 = a.o1 = a.o2 = 0; }
SVB_D void dot_mac(dot_acc& a, u64 x, u64 y) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    asm("mad.lo.cc.u32 %0, %8, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.hi.cc.u32 %6, %8, %11, %6;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
        "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(a.e0), "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4), "+r"(a.o0), "+r"(a.o1), "+r"(a.o2)
        : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
}
// small constant k (< 2^32) times y
SVB_D void dot_mac_small(dot_acc& a, u32 k, u64 y) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32);
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        : "+r"(a.e0), "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4)
        : "r"(k), "r"(y0));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(a.o0), "+r"(a.o1), "+r"(a.o2)
        : "r"(k), "r"(y1));
}
SVB_D u64 dot_reduce(dot_acc a) {
    // T = E + (O << 32), then lo + hi*2^64 + top*2^128 with 2^128 = -2^32
    asm("add.cc.u32 %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, %5;\n\t"
        "addc.cc.u32 %2, %2, %6;\n\t"
        "addc.u32 %3, %3, 0;"
        : "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4)
        : "r"(a.o0), "r"(a.o1), "r"(a.o2));
    u64 r = reduce128(((u64)a.e1 << 32) | a.e0, ((u64)a.e3 << 32) | a.e2);
    u64 sub = (u64)a.e4 << 32;
    u64 d = r - sub;
    return r < sub ? d - GL_EPS : d;
}

// v*x + c -> LOOSE, one carry chain for the product and the addend
SVB_D u64 mul_add_dev(u64 v, u64 x, u64 c) {
    u32 v0 = (u32)v, v1 = (u32)(v >> 32), x0 = (u32)x, x1 = (u32)(x >> 32);
    u32 c0 = (u32)c, c1 = (u32)(c >> 32);
    u32 r0, r1, r2, r3;
    asm("{\n\t"
        ".reg .u32 m0, m1, m2;\n\t"
        "mad.lo.cc.u32 %0, %4, %6, %8;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %9;\n\t"
        "madc.lo.cc.u32 %2, %5, %7, 0;\n\t"
        "madc.hi.u32 %3, %5, %7, 0;\n\t"
        "mul.lo.u32 m0, %4, %7;\n\t"
        "mul.hi.u32 m1, %4, %7;\n\t"
        "mad.lo.cc.u32 m0, %5, %6, m0;\n\t"
        "madc.hi.cc.u32 m1, %5, %6, m1;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "add.cc.u32 %1, %1, m0;\n\t"
        "addc.cc.u32 %2, %2, m1;\n\t"
        "addc.u32 %3, %3, m2;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(v0), "r"(v1), "r"(x0), "r"(x1), "r"(c0), "r"(c1));
    return reduce128(((u64)r1 << 32) | r0, ((u64)r3 << 32) | r2);
}

// Rotate the 12-word state left by 4 words (register moves).
SVB_D void rot4(u64 s[12]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 t = s[i];
        s[i] = s[i + 4];
        s[i + 4] = s[i + 8];
        s[i + 8] = t;
    }
}

// d_FULL_RC_NEXT (derived table, poseidon_g_constants.inc): constants folded into the MDS layer that
// FOLLOWS full round f: f = 0..2 -> ALL_ROUND_CONSTANTS of round f+1, f = 3 ->
// FAST_PARTIAL_FIRST_ROUND_CONSTANT, f = 4..6 -> rounds 27..29, f = 7 -> 0.

SVB_D void poseidon_g_dev(u64 s[12], u64* __restrict__ scratch /* 11 words of per-thread shared memory, stride blockDim.x */,
                      u32 scratch_stride) {
    // round constants of round 0 (poseidon.rs:637-640); later constants ride on the MDS accumulators
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], d_ALL_ROUND_CONSTANTS[i]);
#pragma unroll 1
    for (int f = 0; f < 8; f++) {
        // S-box layer: 4 lanes per iteration, state rotated by 4 between iterations (:438-448)
#pragma unroll 1
        for (int g = 0; g < 3; g++) {
#pragma unroll
            for (int i = 0; i < 4; i++) s[i] = sbox7(s[i]);
            rot4(s);
        }
        mds_layer_rc_f64(s, d_FULL_RC_NEXT + 12 * f);   // (:450-502) + next constant layer
        if (f == 3) {
            // mds_partial_layer_init (:504-537): t[c] = sum_{r=1..11} INIT[r-1][c-1] * s[r]
#pragma unroll 1
            for (int c = 0; c < 11; c++) {
                dot_acc a;
                dot_init(a);
#pragma unroll
                for (int r = 1; r < 12; r++) dot_mac(a, d_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + c], s[r]);
                scratch[c * scratch_stride] = dot_reduce(a);
            }
#pragma unroll
            for (int c = 0; c < 11; c++) s[c + 1] = scratch[c * scratch_stride];
            // 22 partial rounds (:654-672)
#pragma unroll 1
            for (int r = 0; r < 22; r++) {
                u64 s0 = sbox7(s[0]);
                s0 = add_lc(s0, d_FAST_PARTIAL_ROUND_CONSTANTS[r]);   // entry 21 is 0 (:140)
                // mds_partial_layer_fast (:539-589)
                dot_acc a;
                dot_init(a);
                dot_mac_small(a, 25, s0);   // MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0]
#pragma unroll
                for (int i = 1; i < 12; i++) dot_mac(a, d_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1], s[i]);
#pragma unroll
                for (int i = 1; i < 12; i++) s[i] = mul_add_dev(d_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1], s0, s[i]);
                s[0] = dot_reduce(a);
            }
            // constant layer of full round 26 (:675-678)
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], d_ALL_ROUND_CONSTANTS[12 * 26 + i]);
        }
    }
}
#endif

// canonical in, canonical out
SVB_HD void poseidon_g_canonical(u64 s[12]) {
    poseidon_g(s);
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}

}  // namespace svb

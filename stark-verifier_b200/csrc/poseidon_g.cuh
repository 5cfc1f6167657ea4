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
    unsigned __int128 s = ((unsigned __int128)a.hi << 64 | a.lo);
    unsigned __int128 p = ((unsigned __int128)h << 64 | l);
    unsigned __int128 t = s + p;
    a.top += (t < s);
    a.lo = (u64)t;
    a.hi = (u64)(t >> 64);
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

// ------------------------------------------------------------------------------------------------
// The circulant MDS layer in the frequency domain, on exact small integers held in binary64.
// y_i = sum_k CIRC[k] x_{(i+k) mod 12} (+ 8 x_0 on row 0) + rc_i  (poseidon.rs:450-502 computes the same
// rows directly).  Index i = b + 3a: a 4-point DFT over a (roots 1, i, -1, -i) turns the 12 x 12
// circulant into three 3 x 3 blocks, one per frequency:
//     f = 0: [16 16 32; 32 16 16; 16 32 16] * 4     f = 2: [-1 -2 8; -8 -1 -2; 2 -8 -1] * 4
//     f = 1: [2+i 1+16i 1-4i; -4-i 2+i 1+16i; 16-i -4-i 2+i] * 2      (f = 3 is its conjugate)
// and the factors 4, 4, 2 are exactly what the inverse DFT divides by, so everything stays an integer
// (the reason plonky2 chose this MDS vector).  90 additions / FMAs instead of 144 + nothing for rc.
// Inputs < 2^32 give |intermediates| < 2^41: exact in binary64, also for subnormal operands (the device
// feeds u32 bit patterns, i.e. multiples of 2^-1074, and reads the result bits back as the integer).
// RC = 12 adds rc[0..11] to the rows, RC = 1 adds rc[0] to row 0 only, RC = 0 adds nothing.  An all-zero
// input comes out as +0 in every row, never -0 (16 * (+0) + (+-0) = +0 in round-to-nearest), so the result
// bits are the integer in all cases.
SVB_HD double svb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
template <int RC>
SVB_HD void mds_freq_half(const double x[12], const double* rc, double y[12]) {
    double U0[3], U2[3], Ur[3], Ui[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 3; b++) {
        const double s02 = x[b] + x[b + 6], s13 = x[b + 3] + x[b + 9];
        U0[b] = s02 + s13;
        U2[b] = s02 - s13;
        Ur[b] = x[b] - x[b + 6];
        Ui[b] = x[b + 3] - x[b + 9];
    }
    const double S = (U0[0] + U0[1]) + U0[2];
    const double T[3] = {S + U0[2], S + U0[0], S + U0[1]};                 // A_b = 16 T_b
    const double B[3] = {svb_fma(8.0, U2[2], svb_fma(-2.0, U2[1], -U2[0])),
                         svb_fma(-8.0, U2[0], svb_fma(-2.0, U2[2], -U2[1])),
                         svb_fma(2.0, U2[0], svb_fma(-8.0, U2[1], -U2[2]))};
    const double R[3] = {svb_fma(4.0, Ui[2], svb_fma(-16.0, Ui[1], svb_fma(2.0, Ur[0], (Ur[1] + Ur[2]) - Ui[0]))),
                         svb_fma(-16.0, Ui[2], svb_fma(2.0, Ur[1], svb_fma(-4.0, Ur[0], (Ui[0] - Ui[1]) + Ur[2]))),
                         svb_fma(2.0, Ur[2], svb_fma(-4.0, Ur[1], svb_fma(16.0, Ur[0], (Ui[0] + Ui[1]) - Ui[2])))};
    const double I[3] = {svb_fma(-4.0, Ur[2], svb_fma(16.0, Ur[1], svb_fma(2.0, Ui[0], (Ur[0] + Ui[1]) + Ui[2]))),
                         svb_fma(16.0, Ur[2], svb_fma(2.0, Ui[1], svb_fma(-4.0, Ui[0], (Ur[1] - Ur[0]) + Ui[2]))),
                         svb_fma(2.0, Ui[2], svb_fma(-4.0, Ui[1], svb_fma(16.0, Ui[0], (Ur[2] - Ur[0]) - Ur[1])))};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 3; b++) {
        const double P = svb_fma(16.0, T[b], B[b]), Q = svb_fma(16.0, T[b], -B[b]);
        double y0 = P + R[b];
        if (b == 0) y0 = svb_fma(8.0, x[0], y0);                            // MDS_MATRIX_DIAG[0]
        if (RC == 12) {
            y[b] = y0 + rc[b];
            y[b + 3] = (Q + I[b]) + rc[b + 3];
            y[b + 6] = (P - R[b]) + rc[b + 6];
            y[b + 9] = (Q - I[b]) + rc[b + 9];
        } else {
            y[b] = (RC == 1 && b == 0) ? y0 + rc[0] : y0;
            y[b + 3] = Q + I[b];
            y[b + 6] = P - R[b];
            y[b + 9] = Q - I[b];
        }
    }
}

#if defined(__CUDACC__)
// ================================================================================================
// Device path (sm_100a).  What the measurements on B200 said (profiles/, tools/lab/NOTES.md):
//  * the first version unrolled everything (90 KB of SASS per permutation) and spent half of its
//    issue slots stalled on instruction-cache misses; the code below is ~35 KB: ONE full-round body
//    (12 unrolled S-boxes + one MDS layer) shared by both halves, a looped initial matrix, one
//    partial-round body;
//  * IMAD.WIDE.U32 holds the fmaheavy pipe 4.24 cycles per warp and the other pipes overlap with it
//    only partially, so every instruction counts: a modular multiplication is 14 instructions
//    (goldilocks.cuh), dot products keep unreduced limbs (5.5 instructions per term), and the MDS layers
//    run as exact integer arithmetic on the otherwise idle FP64 pipe, fed with subnormals so that no
//    int<->double conversion is needed, in the frequency domain (90 instead of 144 operations per half).
// ================================================================================================
// The MDS layer on the FP64 pipe.  Measured on B200 (tools/microbench/pipes.cu): IMAD.WIDE.U32 holds
// the fmaheavy pipe 4 cycles per warp instruction and the S-boxes already saturate it; DFMA issues
// every 2 cycles on the otherwise idle FP64 pipe.  Every product here is (coefficient <= 49) x (32-bit
// half) and every row sum is < 2^42, so the arithmetic is EXACT in binary64: the accumulator starts at
// 2^52 + rc_half and stays inside [2^52, 2^53), where doubles are consecutive integers, and the
// integer is read back from the mantissa bits (the exponent pattern 0x433 of the high words is
// subtracted inside the carry chain that recombines the two halves).
#ifndef SVB_MDS_CVT
#define SVB_MDS_CVT 2   // 2: subnormal trick (no conversion at all); 1: I2F.F64.U32; 0: mantissa trick (2 MOV + DADD)
#endif
#if SVB_MDS_CVT == 2
// A u32 x placed in the low word of a binary64 with a zero high word IS the subnormal x * 2^-1074: no
// conversion instruction at all (I2F.F64.U32 costs 8.5 cycles per warp).  Subnormal arithmetic is plain
// fixed point with ulp 2^-1074, exact while the integers stay below 2^52 -- here every row sum is < 2^42 --
// and FP64 has no flush-to-zero mode on NVIDIA GPUs.  The accumulator's bit pattern is the integer itself.
SVB_D double u32_to_f64(u32 x) { return __hiloint2double(0, (int)x); }
#elif SVB_MDS_CVT == 1
SVB_D double u32_to_f64(u32 x) { return __uint2double_rn(x); }
#else
SVB_D double u32_to_f64(u32 x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
#endif
// V = al + ah * 2^32 with al, ah < 2^42 (the two half-sums of one row) -> LOOSE u64
SVB_D u64 mds_recombine(u32 al0, u32 al1, u32 ah0, u32 ah1) {
    // V = v0 + v1 W + v2 W^2 (v2 < 2^11); result = v2*EPS + (v1:v0) + cy*EPS
    u32 lo, hi;
    asm("{\n\t.reg .u32 x1, v1, v2, t0, t1, cy, h2;\n\t"
#if SVB_MDS_CVT == 2
        "add.cc.u32 v1, %3, %4;\n\t addc.u32 v2, %5, 0;\n\t"
#else
        "sub.u32 x1, %3, 0x43300000;\n\t"
        "add.cc.u32 v1, x1, %4;\n\t addc.u32 v2, %5, 0xBCD00000;\n\t"
#endif
        "mad.lo.cc.u32 t0, v2, %6, %2;\n\t madc.hi.cc.u32 t1, v2, %6, v1;\n\t addc.u32 cy, 0, 0;\n\t"
        "sub.cc.u32 %0, t0, cy;\n\t subc.u32 h2, t1, 0;\n\t add.u32 %1, h2, cy;\n\t"
        "}" : "=r"(lo), "=r"(hi) : "r"(al0), "r"(al1), "r"(ah0), "r"(ah1), "r"(d_EPS32));
    return ((u64)hi << 32) | lo;
}
#ifndef SVB_MDS_FFT
#define SVB_MDS_FFT 1   // 1: frequency-domain form (mds_freq_half: 180 FP64 operations per layer, +3.3 % perms/s); 0: 288 DFMA
#endif
#if SVB_MDS_FFT && SVB_MDS_CVT != 2
#error "SVB_MDS_FFT needs the subnormal operand form (SVB_MDS_CVT == 2)"
#endif
// RC = 12: the layer adds the 12 constants rcf (2 subnormal words per row); RC = 0: no constants (rcf unused)
template <int RC = 12>
SVB_D void mds_layer_rc_f64(u64 s[12], const u64* __restrict__ rcf /* FULL_RC_NEXT_{F64,SUBNORMAL}: 2 words per row */) {
#if SVB_MDS_FFT
    double x[12], rc[12], yl[12], yh[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        x[j] = u32_to_f64((u32)s[j]);
        if (RC) rc[j] = __longlong_as_double((long long)rcf[2 * j]);
    }
    mds_freq_half<RC>(x, rc, yl);
#pragma unroll
    for (int j = 0; j < 12; j++) {
        x[j] = u32_to_f64((u32)(s[j] >> 32));
        if (RC) rc[j] = __longlong_as_double((long long)rcf[2 * j + 1]);
    }
    mds_freq_half<RC>(x, rc, yh);
#pragma unroll
    for (int r = 0; r < 12; r++)
        s[r] = mds_recombine((u32)__double2loint(yl[r]), (u32)__double2hiint(yl[r]), (u32)__double2loint(yh[r]), (u32)__double2hiint(yh[r]));
#else
    double d[12];
    u32 h[12];
    u32 al0[12], al1[12];
#pragma unroll
    for (int j = 0; j < 12; j++) {
        d[j] = u32_to_f64((u32)s[j]);
        h[j] = (u32)(s[j] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double acc = __longlong_as_double((long long)rcf[2 * r]);       // the starting value carries the round constant
#pragma unroll
        for (int j = 0; j < 12; j++) acc = __fma_rn(d[j], (double)mds_coeff(r, j), acc);
        al0[r] = (u32)__double2loint(acc);
        al1[r] = (u32)__double2hiint(acc);
    }
#pragma unroll
    for (int j = 0; j < 12; j++) d[j] = u32_to_f64(h[j]);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        double acc = __longlong_as_double((long long)rcf[2 * r + 1]);
#pragma unroll
        for (int j = 0; j < 12; j++) acc = __fma_rn(d[j], (double)mds_coeff(r, j), acc);
        s[r] = mds_recombine(al0[r], al1[r], (u32)__double2loint(acc), (u32)__double2hiint(acc));
    }
#endif
}

// Unreduced sum of up to 16 products of 64-bit words: E = e0..e4 collects x0*y0 (weight 1) and x1*y1
// (weight W^2), chained through the IMAD.WIDE carry; O = o1..o3 collects the cross products at weight
// W.  ptxas maps each term to 4 IMAD.WIDE.U32 and merges the carry captures into dual-carry-in
// IADD3.X (5.5 instructions per term).
struct dot_acc {
    u32 e0, e1, e2, e3, e4, o1, o2, o3;
};
SVB_D void dot_init(dot_acc& a) { a.e0 = a.e1 = a.e2 = a.e3 = a.e4 = a.o1 = a.o2 = a.o3 = 0; }
SVB_D void dot_mac(dot_acc& a, u64 x, u64 y) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    asm("mad.lo.cc.u32 %0, %8, %10, %0;\n\t madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t madc.hi.cc.u32 %3, %9, %11, %3;\n\t addc.u32 %4, %4, 0;\n\t"
        "mad.lo.cc.u32 %5, %8, %11, %5;\n\t madc.hi.cc.u32 %6, %8, %11, %6;\n\t addc.u32 %7, %7, 0;\n\t"
        "mad.lo.cc.u32 %5, %9, %10, %5;\n\t madc.hi.cc.u32 %6, %9, %10, %6;\n\t addc.u32 %7, %7, 0;"
        : "+r"(a.e0), "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4), "+r"(a.o1), "+r"(a.o2), "+r"(a.o3)
        : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
}
// small constant k (< 2^32) times y
SVB_D void dot_mac_small(dot_acc& a, u32 k, u64 y) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32);
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t addc.cc.u32 %3, %3, 0;\n\t addc.u32 %4, %4, 0;"
        : "+r"(a.e0), "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4)
        : "r"(k), "r"(y0));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t addc.u32 %2, %2, 0;"
        : "+r"(a.o1), "+r"(a.o2), "+r"(a.o3)
        : "r"(k), "r"(y1));
}
SVB_D u64 dot_reduce(dot_acc a) {
    // T = E + O*W (5 limbs), then red5 folds W^4 = -W
    asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.u32 %3, %3, 0;"
        : "+r"(a.e1), "+r"(a.e2), "+r"(a.e3), "+r"(a.e4)
        : "r"(a.o1), "r"(a.o2), "r"(a.o3));
    return red5(a.e0, a.e1, a.e2, a.e3, a.e4);
}

// x^7 + c: the additive constant rides on the last product
SVB_D u64 sbox7_add(u64 x, u64 c) {
    u64 x2 = mul(x, x);
    u64 x4 = mul(x2, x2);
    u64 x3 = mul(x, x2);
    return mul_add(x3, x4, c);
}

// S-box lanes per loop iteration of the full-round S-box layer: 12 = fully unrolled (no register
// rotation, largest code), 6 or 4 = looped with the state rotated between iterations.
#ifndef SVB_SBOX_LANES
#define SVB_SBOX_LANES 12
#endif
// Rotate the 12-word state left by SVB_SBOX_LANES words (register moves).
SVB_D void rot_lanes(u64 s[12]) {
    u64 t[12];
#pragma unroll
    for (int i = 0; i < 12; i++) t[i] = s[(i + SVB_SBOX_LANES) % 12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = t[i];
}

// d_FULL_RC_NEXT (derived table, poseidon_g_constants.inc): constants folded into the MDS layer that
// FOLLOWS full round f: f = 0..2 -> ALL_ROUND_CONSTANTS of round f+1, f = 3 ->
// FAST_PARTIAL_FIRST_ROUND_CONSTANT, f = 4..6 -> rounds 27..29, f = 7 -> 0.

// SVB_PARTIAL_NAIVE = 1: the 22 partial rounds in the NAIVE form of the permutation (constant layer, x^7 on
// lane 0, full MDS -- poseidon_spec/constants.rs:7-443 with the rounds of poseidon.rs:634-686 before the
// sparse-matrix optimisation; the two forms are the same function, pinned by the known-answer tests).
// On a CPU the sparse form wins (22 multiplications instead of 144 per round); here the dense MDS costs
// 180 operations on the otherwise idle FP64 pipe, while the sparse form costs 120 IMAD.WIDE on the pipe
// that binds: the naive form HALVES the multiplier load of a permutation (5 198 -> 2 7xx IMAD.WIDE) and
// needs no initial matrix and no scratch.  Measured on B200: 1 206 -> 1 394 M perms/s.  Two refinements:
//  * lanes 1..11 see no S-box in rounds 4..25, so their round constants commute with the S-box layer and are
//    pushed through the linear MDS into the next round (tools/gen_poseidon_constants.py derives d_r): a
//    partial round adds ONE scalar, to lane 0 (NAIVE_LANE0_RC), and the accumulated vector d_26 is added
//    once before round 26 (NAIVE_PRE_ROUND26);
//  * the partial rounds run as 11 DOUBLE layers: after the first MDS only lane 0 is recombined and reduced
//    (it feeds the next S-box); lanes 1..11 stay as their two exact half-sums (< 2^41) and go straight
//    into the second MDS, whose sums (< 2^49) are still exact in binary64.  That removes 11 of every 24
//    recombinations (one IMAD.WIDE + 8 add/sub each) and the matching integer -> double operand moves.
#ifndef SVB_PARTIAL_NAIVE
#define SVB_PARTIAL_NAIVE 1
#endif
#ifndef SVB_RC_PRE_MDS
#define SVB_RC_PRE_MDS 1   // full rounds: next round's constants as M^-1 c on the S-box outputs (x3 * x4 + c) instead of 24 DADD in the MDS layer
#endif
#if SVB_PARTIAL_NAIVE && !(SVB_MDS_CVT == 2 && SVB_MDS_FFT)
#error "SVB_PARTIAL_NAIVE needs the subnormal, frequency-domain MDS (SVB_MDS_CVT == 2, SVB_MDS_FFT == 1)"
#endif
SVB_D u64 mds_recombine_d(double yl, double yh) {
    return mds_recombine((u32)__double2loint(yl), (u32)__double2hiint(yl), (u32)__double2loint(yh), (u32)__double2hiint(yh));
}

SVB_D void poseidon_g_dev(u64 s[12], u64* __restrict__ scratch /* 11 words of per-thread shared memory, stride blockDim.x */,
                      u32 scratch_stride) {
#if SVB_PARTIAL_NAIVE
    (void)scratch; (void)scratch_stride;
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        // constants of round 0 (poseidon.rs:637-640) / accumulated constants d_26 of round 26
        const u64* __restrict__ pre = phase ? d_NAIVE_PRE_ROUND26 : d_ALL_ROUND_CONSTANTS;
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], pre[i]);
        // four full rounds: S-box layer, MDS + constants of the next round (:641-650, :675-684)
#pragma unroll 1
        for (int f = 0; f < 4; f++) {
#if SVB_RC_PRE_MDS
            // the constants of the next round, moved in front of the MDS layer (M^-1 c): they ride on the last product of x^7
            const u64* __restrict__ cpre = d_NAIVE_FULL_RC_PRE_MDS + 12 * (4 * phase + f);
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = sbox7_add(s[i], cpre[i]);
            mds_layer_rc_f64<0>(s, nullptr);
#else
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
            mds_layer_rc_f64(s, d_NAIVE_FULL_RC_NEXT_SUBNORMAL + 24 * (4 * phase + f));
#endif
        }
        if (phase == 0) {
            // rounds 4..25 as 11 double layers
#pragma unroll 1
            for (int k = 0; k < 11; k++) {
                const u64* __restrict__ l0 = d_NAIVE_LANE0_RC_SUBNORMAL + 4 * k;
                double x[12], yl[12], yh[12], rc;
                s[0] = sbox7(s[0]);
#pragma unroll
                for (int j = 0; j < 12; j++) x[j] = u32_to_f64((u32)s[j]);
                rc = __longlong_as_double((long long)l0[0]);
                mds_freq_half<1>(x, &rc, yl);
#pragma unroll
                for (int j = 0; j < 12; j++) x[j] = u32_to_f64((u32)(s[j] >> 32));
                rc = __longlong_as_double((long long)l0[1]);
                mds_freq_half<1>(x, &rc, yh);
                const u64 s0 = sbox7(mds_recombine_d(yl[0], yh[0]));
                yl[0] = u32_to_f64((u32)s0);
                yh[0] = u32_to_f64((u32)(s0 >> 32));
                rc = __longlong_as_double((long long)l0[2]);
                mds_freq_half<1>(yl, &rc, x);
#pragma unroll
                for (int j = 0; j < 12; j++) yl[j] = x[j];
                rc = __longlong_as_double((long long)l0[3]);
                mds_freq_half<1>(yh, &rc, x);
#pragma unroll
                for (int j = 0; j < 12; j++) s[j] = mds_recombine_d(yl[j], x[j]);
            }
        }
    }
    return;
#else
    // round constants of round 0 (poseidon.rs:637-640); later constants ride on the MDS accumulators
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], d_ALL_ROUND_CONSTANTS[i]);
    // SVB_ABLATE (tools/lab only; the result is WRONG when set): bit 0 skips the full-round S-boxes, 1 the
    // MDS layers, 2 the initial matrix, 3 the partial-round dot product, 4 the partial-round rank-1
    // update, 5 the partial-round S-box.  Timing differences give the cost of each section.
#ifndef SVB_ABLATE
#define SVB_ABLATE 0
#endif
#pragma unroll 1
    for (int f = 0; f < 8; f++) {
        // S-box layer: 4 lanes per iteration, state rotated by 4 between iterations (:438-448)
#if SVB_ABLATE & 1
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] ^= s[(i + 1) % 12] >> 3;
#elif SVB_SBOX_LANES == 12
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = sbox7(s[i]);
#else
#pragma unroll 1
        for (int g = 0; g < 12 / SVB_SBOX_LANES; g++) {
#pragma unroll
            for (int i = 0; i < SVB_SBOX_LANES; i++) s[i] = sbox7(s[i]);
            rot_lanes(s);
        }
#endif
#if SVB_ABLATE & 2
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] += s[(i + 5) % 12] ^ d_FULL_RC_NEXT[12 * f + i];
#else
#if SVB_MDS_CVT == 2
        mds_layer_rc_f64(s, d_FULL_RC_NEXT_SUBNORMAL + 24 * f);   // (:450-502) + next constant layer
#else
        mds_layer_rc_f64(s, d_FULL_RC_NEXT_F64 + 24 * f);   // (:450-502) + next constant layer
#endif
#endif
        if (f == 3) {
            // mds_partial_layer_init (:504-537): t[c] = sum_{r=1..11} INIT[r-1][c-1] * s[r]
#if !(SVB_ABLATE & 4)
#pragma unroll 1
            for (int c = 0; c < 11; c++) {
                dot_acc a;
                dot_init(a);
#pragma unroll
                for (int r = 1; r < 12; r++) dot_mac(a, s[r], d_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + c]);
                scratch[c * scratch_stride] = dot_reduce(a);
            }
#pragma unroll
            for (int c = 0; c < 11; c++) s[c + 1] = scratch[c * scratch_stride];
#endif
            // 22 partial rounds (:654-672)
#pragma unroll 1
            for (int r = 0; r < 22; r++) {
#if SVB_ABLATE & 32
                u64 s0 = s[0] + d_FAST_PARTIAL_ROUND_CONSTANTS[r];
#else
                u64 s0 = sbox7_add(s[0], d_FAST_PARTIAL_ROUND_CONSTANTS[r]);   // entry 21 is 0 (:140)
#endif
                // mds_partial_layer_fast (:539-589)
                dot_acc a;
                dot_init(a);
                dot_mac_small(a, 25, s0);   // MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0]
#if !(SVB_ABLATE & 8)
#pragma unroll
                for (int i = 1; i < 12; i++) dot_mac(a, s[i], d_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1]);
#endif
#if SVB_ABLATE & 16
#pragma unroll
                for (int i = 1; i < 12; i++) s[i] += s0 ^ d_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1];
#else
#pragma unroll
                for (int i = 1; i < 12; i++) s[i] = mul_add(d_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1], s0, s[i]);
#endif
                s[0] = dot_reduce(a);
            }
            // constant layer of full round 26 (:675-678)
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = add_lc(s[i], d_ALL_ROUND_CONSTANTS[12 * 26 + i]);
        }
    }
#endif   // SVB_PARTIAL_NAIVE
}
#endif

// canonical in, canonical out
SVB_HD void poseidon_g_canonical(u64 s[12]) {
    poseidon_g(s);
    for (int i = 0; i < 12; i++) s[i] = canon(s[i]);
}

}  // namespace svb

// Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[X]/(X^2 - 7).
//
// Semantics replaced: GoldilocksChip / GoldilocksExtensionChip of the reference
//   chip/goldilocks_chip.rs:105-404, chip/goldilocks_extension_chip.rs:49-400,
//   gate r = a*b + c mod p: native_chip/arithmetic_chip.rs:19,98-132.
// One implementation for host (synthetic prover, transcript) and device (sm_100a kernels).
//
// Representation: a field element travels as a u64 that is either CANONICAL (< p; everything in
// memory, everything compared) or LOOSE (any u64, congruent mod p; the inside of the Poseidon
// permutation).  2^64 = 2^32 - 1 =: EPS and 2^96 = -1 (mod p) make the reduction of a 128-bit product
// three add/sub steps; on the device they are carry-chain PTX so that ptxas emits
// IMAD.WIDE.U32 / IADD3.X with predicate carries (checked with cuobjdump -sass).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SVB_HD __host__ __device__ __forceinline__
#define SVB_D __device__ __forceinline__
#else
#define SVB_HD inline
#define SVB_D inline
#endif

namespace svb {

typedef uint64_t u64;
typedef uint32_t u32;

static constexpr u64 GL_P = 0xFFFFFFFF00000001ull;
static constexpr u64 GL_EPS = 0xFFFFFFFFull;

// ---- 64x64 -> 128 --------------------------------------------------------------------------
// Portable forms (host: transcript, synthetic prover; the device kernels use the limb primitives below).
SVB_HD void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (u64)p;
    hi = (u64)(p >> 64);
#endif
}
SVB_HD void sqr_wide(u64 a, u64& lo, u64& hi) { mul_wide(a, a, lo, hi); }

// (hi:lo) mod p as a LOOSE u64.  lo + hi_lo*EPS - hi_hi with the two wrap corrections
// (each can fire at most once: after a wrap the value is small).
SVB_HD u64 reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
#if defined(__CUDA_ARCH__)
    if (lo < hh) t0 -= GL_EPS;
    u64 t1 = hl * GL_EPS;
    u64 t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
#else
    // host: the wraps are data-dependent coin flips; as masks they cost nothing, as branches a misprediction each
    t0 -= GL_EPS & (0 - (u64)(lo < hh));
    u64 t1 = hl * GL_EPS;
    u64 t2 = t0 + t1;
    t2 += GL_EPS & (0 - (u64)(t2 < t1));
#endif
    return t2;
}

// lo + hi32 * 2^64 (hi32 < 2^32) mod p as a LOOSE u64.
SVB_HD u64 reduce96(u64 lo, u32 hi32) {
    u64 t1 = (u64)hi32 * GL_EPS;
    u64 t2 = lo + t1;
    if (t2 < t1) t2 += GL_EPS;
    return t2;
}

#if defined(__CUDACC__)
// ---- device limb primitives (sm_100a) -----------------------------------------------------------
// Cost model measured on B200 (tools/microbench/pipes2.cu, profiles/pipes2_b200_r1.txt): IMAD.WIDE.U32 (32x32+64 -> 64, optional
// carry-out predicate, .X = carry-in) occupies the fmaheavy pipe for 4 cycles per warp, IMAD.HI 4,
// other IMAD forms 2; IADD3 issues every cycle, IADD3.X / LOP3 / SHF / SEL every 2.  The sequences
// below are written so that ptxas maps every mad.lo.cc/madc.hi pair onto ONE IMAD.WIDE.U32 and
// merges carry captures into dual-carry IADD3.X (checked with cuobjdump -sass; a*b mod p is 5
// IMAD.WIDE + 9 ALU instructions).
//
// 0xFFFFFFFF lives in constant memory on purpose: as an immediate ptxas strength-reduces the
// multiplication by EPS into IMAD.HI + IMAD.IADD (6 fmaheavy cycles instead of 4).
__constant__ u32 d_EPS32 = 0xFFFFFFFFu;

// a*b = r0 + r1 W + r2 W^2 + r3 W^3  (W = 2^32).  The carry of the cross-term sum N = a1*b0 + a0*b1 and the
// carry of the column additions both land in r3 through `addc ..., 0` instructions that ptxas merges
// into ONE IADD3.X with two carry-in predicates (no SEL to materialise a flag).
SVB_D void mulw4(u64 a, u64 b, u32& r0, u32& r1, u32& r2, u32& r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m0, m1, n0, n1, r3p, p1, q0, q1;\n\t.reg .u64 P, Q, M;\n\t"
        "mul.wide.u32 Q, %5, %7;\n\t mov.b64 {q0, q1}, Q;\n\t"
        "mul.wide.u32 M, %4, %7;\n\t mov.b64 {m0, m1}, M;\n\t"
        "mad.lo.cc.u32 n0, %5, %6, m0;\n\t madc.hi.cc.u32 n1, %5, %6, m1;\n\t addc.u32 r3p, q1, 0;\n\t"
        "mul.wide.u32 P, %4, %6;\n\t mov.b64 {%0, p1}, P;\n\t"
        "add.cc.u32 %1, p1, n0;\n\t"
        "addc.cc.u32 %2, q0, n1;\n\t"
        "addc.u32 %3, r3p, 0;\n\t"
        "}" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// a*b + c (c any u64): the addend rides on the a0*b0 product, its carry on the a1*b1 product
SVB_D void muladdw4(u64 a, u64 b, u64 cc, u32& r0, u32& r1, u32& r2, u32& r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), c0 = (u32)cc, c1 = (u32)(cc >> 32);
    asm("{\n\t.reg .u32 m0, m1, n0, n1, r3p, p1, q0, q1;\n\t.reg .u64 M;\n\t"
        "mad.lo.cc.u32 %0, %4, %6, %8;\n\t madc.hi.cc.u32 p1, %4, %6, %9;\n\t"
        "madc.lo.cc.u32 q0, %5, %7, 0;\n\t madc.hi.u32 q1, %5, %7, 0;\n\t"
        "mul.wide.u32 M, %4, %7;\n\t mov.b64 {m0, m1}, M;\n\t"
        "mad.lo.cc.u32 n0, %5, %6, m0;\n\t madc.hi.cc.u32 n1, %5, %6, m1;\n\t addc.u32 r3p, q1, 0;\n\t"
        "add.cc.u32 %1, p1, n0;\n\t"
        "addc.cc.u32 %2, q0, n1;\n\t"
        "addc.u32 %3, r3p, 0;\n\t"
        "}" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c0), "r"(c1));
}
// r0 + r1 W + r2 W^2 + r3 W^3 + r4 W^4 as a LOOSE u64 (W^2 = W - 1, W^3 = -1, W^4 = -W mod p).
//   T = r2 * EPS + (r1:r0)      one IMAD.WIDE with carry-out cy
//   u = T - (r4:r3)             borrow b
//   k = cy - b in {-1, 0, 1}    number of 2^64 wraps; 2^64 = EPS, so the result is u + k*EPS, computed as
//                               (u1:u0) - sign_extend(k) + (k << 32); it cannot wrap again.
// 1 IMAD.WIDE + 6 ALU instructions (a*b mod p = 5 IMAD.WIDE + 9 ALU in total).  (A 7-instruction variant that feeds the borrow of a sub.cc chain
// into madc.cc was tried: ptxas passes the inverted flag there, see tools/lab/NOTES.md.)
// Exhaustive corner-limb model: tools/lab/reduce_model.py; on the device:
// tests/test_gpu_parity.py::test_field_corner_cases.
#ifndef SVB_RED_ALU
#define SVB_RED_ALU 0   // 1: reduction without the IMAD.WIDE by EPS (12 add/sub instructions)
#endif
SVB_D u64 red5(u32 r0, u32 r1, u32 r2, u32 r3, u32 r4) {
    u32 lo, hi;
#if SVB_RED_ALU
    asm("{\n\t.reg .u32 t0, t1, k, kh, h2;\n\t"
        "sub.cc.u32 t0, %2, %4;\n\t subc.cc.u32 t1, %3, 0;\n\t subc.u32 k, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, %5;\n\t subc.cc.u32 t1, t1, %6;\n\t subc.u32 k, k, 0;\n\t"
        "add.cc.u32 t1, t1, %4;\n\t addc.u32 k, k, 0;\n\t"
        "shr.s32 kh, k, 31;\n\t"
        "sub.cc.u32 %0, t0, k;\n\t subc.u32 h2, t1, kh;\n\t add.u32 %1, h2, k;\n\t"
        "}" : "=r"(lo), "=r"(hi) : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4));
#else
    // k = -1 + cy + (1 - b): two `..c k, ..` instructions that ptxas merges into one dual-carry-in IADD3.X
    asm("{\n\t.reg .u32 t0, t1, u0, u1, k, kh, h2;\n\t"
        "mad.lo.cc.u32 t0, %4, %7, %2;\n\t madc.hi.cc.u32 t1, %4, %7, %3;\n\t addc.u32 k, 0xFFFFFFFF, 0;\n\t"
        "sub.cc.u32 u0, t0, %5;\n\t subc.cc.u32 u1, t1, %6;\n\t subc.u32 k, k, 0xFFFFFFFF;\n\t"
        "shr.s32 kh, k, 31;\n\t"
        "sub.cc.u32 %0, u0, k;\n\t subc.u32 h2, u1, kh;\n\t add.u32 %1, h2, k;\n\t"
        "}" : "=r"(lo), "=r"(hi) : "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(d_EPS32));
#endif
    return ((u64)hi << 32) | lo;
}
SVB_D u64 red4(u32 r0, u32 r1, u32 r2, u32 r3) { return red5(r0, r1, r2, r3, 0); }
#endif

SVB_HD u64 canon(u64 a) { return a >= GL_P ? a - GL_P : a; }
SVB_HD bool is_canonical(u64 a) { return a < GL_P; }

// LOOSE x LOOSE -> LOOSE
SVB_HD u64 mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 r0, r1, r2, r3;
    mulw4(a, b, r0, r1, r2, r3);
    return red4(r0, r1, r2, r3);
#else
    u64 lo, hi; mul_wide(a, b, lo, hi); return reduce128(lo, hi);
#endif
}
SVB_HD u64 sqr(u64 a) {
#if defined(__CUDA_ARCH__)
    return mul(a, a);
#else
    u64 lo, hi; sqr_wide(a, lo, hi); return reduce128(lo, hi);
#endif
}
// a*b + c, all LOOSE
SVB_HD u64 mul_add(u64 a, u64 b, u64 c) {
#if defined(__CUDA_ARCH__)
    u32 r0, r1, r2, r3;
    muladdw4(a, b, c, r0, r1, r2, r3);
    return red4(r0, r1, r2, r3);
#else
    u64 lo, hi;
    mul_wide(a, b, lo, hi);
    u64 l2 = lo + c;
    hi += (l2 < lo);   // a*b + c < 2^128: no overflow
    return reduce128(l2, hi);
#endif
}
// LOOSE + CANONICAL -> LOOSE   (after a wrap the sum is < b < p, so + EPS cannot wrap again)
SVB_HD u64 add_lc(u64 a, u64 b_canonical) {
    u64 s = a + b_canonical;
    return s < a ? s + GL_EPS : s;
}
// canonical ops (inputs canonical, output canonical)
// Device: borrow-mask forms, 5 instructions for a - b (t = a - b; a borrow means t is 2^64 too large, i.e. EPS = 2^32 - 1
// too large mod p, and the mask the borrow chain leaves in a register IS that EPS) against 8 for compare + select; a + b
// as a - (p - b), 7 against 10 (p - 0 = p is a harmless non-canonical subtrahend: a - p borrows and comes back as a).  The
// NTT butterflies and the FRI algebra are made of these.
SVB_HD u64 sub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1;
    asm("{\n\t.reg .u32 t0, t1, m;\n\t"
        "sub.cc.u32 t0, %2, %4;\n\t subc.cc.u32 t1, %3, %5;\n\t subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, t0, m;\n\t subc.u32 %1, t1, 0;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((u64)r1 << 32) | r0;
#else
    return a >= b ? a - b : a + (GL_P - b);
#endif
}
SVB_HD u64 add(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return sub(a, GL_P - b);
#else
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
#endif
}
SVB_HD u64 neg(u64 a) { return a ? GL_P - a : 0; }
SVB_HD u64 mulc(u64 a, u64 b) { return canon(mul(a, b)); }

SVB_HD u64 pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = mulc(r, b);
        b = mulc(b, b);
        e >>= 1;
    }
    return r;
}
// Fermat inverse with the standard Goldilocks addition chain for p - 2 = 2^64 - 2^32 - 1
// (72 multiplications).  a != 0.
SVB_HD u64 sqn(u64 a, int n) { for (int i = 0; i < n; i++) a = sqr(a); return a; }
SVB_HD u64 inv(u64 a) {
    // p - 2 = 0xFFFFFFFE_FFFFFFFF : 31 ones, a zero, 32 ones
    u64 t2 = mul(sqr(a), a);              // 2 ones
    u64 t3 = mul(sqr(t2), a);             // 3 ones
    u64 t6 = mul(sqn(t3, 3), t3);         // 6
    u64 t12 = mul(sqn(t6, 6), t6);        // 12
    u64 t24 = mul(sqn(t12, 12), t12);     // 24
    u64 t30 = mul(sqn(t24, 6), t6);       // 30
    u64 t31 = mul(sqr(t30), a);           // 31
    u64 t32 = mul(sqr(t31), a);           // 32 ones
    u64 t63 = mul(sqn(t31, 33), t32);     // ones at [63:33], zero at bit 32, ones at [31:0]
    return canon(t63);
}

// ---- quadratic extension -----------------------------------------------------------------------
struct fp2 {
    u64 c0, c1;
};
SVB_HD fp2 mk2(u64 a, u64 b) { fp2 r; r.c0 = a; r.c1 = b; return r; }
SVB_HD fp2 add2(fp2 a, fp2 b) { return mk2(add(a.c0, b.c0), add(a.c1, b.c1)); }
SVB_HD fp2 sub2(fp2 a, fp2 b) { return mk2(sub(a.c0, b.c0), sub(a.c1, b.c1)); }
SVB_HD bool eq2(fp2 a, fp2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }
SVB_HD bool is_zero2(fp2 a) { return (a.c0 | a.c1) == 0; }
// (a0 b0 + 7 a1 b1, a0 b1 + a1 b0): arithmetic_chip.rs:109-132.  Inputs canonical.
SVB_HD fp2 mul2(fp2 a, fp2 b) {
    u64 t = mul(a.c1, b.c1);
    u64 c0 = mul_add(t, 7, mul(a.c0, b.c0));
    u64 c1 = mul_add(a.c0, b.c1, mul(a.c1, b.c0));
    return mk2(canon(c0), canon(c1));
}
SVB_HD fp2 mul_add2(fp2 a, fp2 b, fp2 c) { return add2(mul2(a, b), c); }
// fp2 * base-field scalar
SVB_HD fp2 scale2(fp2 a, u64 s) { return mk2(mulc(a.c0, s), mulc(a.c1, s)); }
// (a0 + a1 X)^-1 = (a0 - a1 X)/(a0^2 - 7 a1^2); a != 0 (norm != 0 because 7 is a non-residue)
SVB_HD fp2 inv2(fp2 a) {
    u64 n = sub(mulc(a.c0, a.c0), mulc(7, mulc(a.c1, a.c1)));
    u64 ni = inv(n);
    return mk2(mulc(a.c0, ni), mulc(neg(a.c1), ni));
}

}  // namespace svb

// The 2^k-ary FRI fold: P'(x^r) from the r = 2^k values of P on the coset of x.
//
// Reference: next_eval (chip/fri_chip.rs:168-226) interpolates {(coset_start * g^i, evals_rev[i])} and evaluates at
// beta, but only implements r = 2 (:211); plonky2's native verifier (compute_evaluation, which the reference's demo
// relies on for its ConstantArityBits(3, 5) proofs, plonky2_semaphore/access_set.rs:124,174) does the same
// interpolation for any r.  The value is  sum_t beta^t P_t(x^r)  where P(X) = sum_t X^t P_t(X^r): a unique field
// element, so HOW it is computed is free.  Here: k successive 2-ary folds with beta, beta^2, beta^4, ... --
//     f'(y^2) = (f(y) + f(-y))/2 + b (f(y) - f(-y))/(2 y)
// -- because in LEAF order (reverse_index_bits, :188-189) the points of entries 2m and 2m+1 are y and -y at every
// level, and the folded values come out in the next level's leaf order.  One inversion of x per query (carried along
// by squaring) replaces the per-step division; the powers of the 2^k-th root are a 16-entry table.  For r = 2 this is
// exactly the two-point formula of :212-224.  The test suite's two CPU restatements compute the same value by barycentric
// and by Lagrange interpolation instead -- three routes to one number.
#pragma once
#include "goldilocks.cuh"

namespace svb {

// omega_16^j, omega_16 = 7^((p-1)/16); the primitive 2^k-th root the reference uses is 7^((p-1)/2^k) = omega_16^(16/2^k)
SVB_HD u64 w16_pow(u32 j) {
    constexpr u64 T[16] = {0x0000000000000001ull, 0xefffffff00000001ull, 0xfffffffeff000001ull, 0x000ffffffff00000ull,
                           0x0001000000000000ull, 0x0000000000001000ull, 0xfffffeff00000101ull, 0xffffffef00000001ull,
                           0xffffffff00000000ull, 0x1000000000000000ull, 0x0000000001000000ull, 0xffefffff00100001ull,
                           0xfffeffff00000001ull, 0xfffffffefffff001ull, 0x000000ffffffff00ull, 0x0000001000000000ull};
    return T[j & 15];
}
SVB_HD u32 bitrev_small(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// ev: 2^ab Fp2 values (2 words each, canonical) in leaf order; x: the query's point on this layer, inv_x = 1/x;
// within = x_index & (2^ab - 1).  1 <= ab <= 4.
SVB_HD fp2 fri_fold(u32 ab, const u64* __restrict__ ev, u64 x, u64 inv_x, u32 within, fp2 beta) {
    (void)x;
    const u32 r = 1u << ab;
    // coset_start = x * g^-rev(within) (:191-200), so 1/coset_start = inv_x * g^rev(within)
    u64 inv_cs = mulc(inv_x, w16_pow(bitrev_small(within, ab) << (4 - ab)));
    fp2 v[16];
    for (u32 j = 0; j < r; j++) v[j] = mk2(ev[2 * j], ev[2 * j + 1]);
    fp2 b = beta;
    for (u32 lb = ab; lb >= 1; lb--) {                   // a level of 2^lb values, root g_l = omega_16^(16 >> lb)
        const u32 half = 1u << (lb - 1);
        for (u32 m = 0; m < half; m++) {
            // y_m = cs_l * g_l^bitrev_lb(2m)  =>  1/y_m = inv_cs_l * omega_16^(-(bitrev << (4 - lb)))
            const u32 e = bitrev_small(2 * m, lb) << (4 - lb);
            const u64 iy = mulc(inv_cs, w16_pow(16 - e));
            const fp2 s = add2(v[2 * m], v[2 * m + 1]), d = sub2(v[2 * m], v[2 * m + 1]);
            v[m] = add2(s, mul2(b, scale2(d, iy)));
        }
        inv_cs = mulc(inv_cs, inv_cs);
        b = mul2(b, b);
    }
    return scale2(v[0], GL_P - ((GL_P - 1) >> ab));     // the halvings, once: 2^-ab = p - (p-1)/2^ab
}

}  // namespace svb

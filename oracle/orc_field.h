/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the Goldilocks base field F_p, p = 2^64 - 2^32 + 1, and of the quadratic
 * extension F_p[X]/(X^2 - 7) exactly as the reference constrains them:
 *   modulus ............ native_chip/arithmetic_chip.rs:19  (GOLDILOCKS_MODULUS)
 *   r = a*b + c mod p .. native_chip/arithmetic_chip.rs:98-107 (base gate)
 *   ext gate ........... native_chip/arithmetic_chip.rs:109-132 (a0b0 + 7 a1b1 + c0, a0b1 + a1b0 + c1)
 *   W = 7 .............. chip/goldilocks_extension_chip.rs:49
 * The native arithmetic itself lives in the un-vendored dependency plonky2_field 0.1.0
 * (DoHoonKim8/plonky2 @ 72229c47, Cargo.lock:1573-1612); this file restates its published
 * algorithm with 128-bit integers, always returning the canonical representative in [0, p).
 *
 * Parity status: pinned by the field KATs of SURVEY.md section 8c (7 generates F_p^*,
 * 7^((p-1)/2^32) = 1753635133440165772, 7^((p-1)/2) = p-1) -- see tests/test_oracle_kat.py.
 */
#ifndef ORC_FIELD_H
#define ORC_FIELD_H
#include <stdint.h>
#include <stddef.h>

#define ORC_P 0xFFFFFFFF00000001ULL
typedef unsigned __int128 orc_u128;
typedef struct { uint64_t c[2]; } orc_fp2;

/* Reference reduction (definition): x mod p with the compiler's 128-bit remainder. */
static inline uint64_t orc_red128_slow(orc_u128 x) { return (uint64_t)(x % ORC_P); }
/* plonky2_field's published reduce128 (goldilocks_field.rs, `reduce128`): with EPS = 2^32 - 1,
 * 2^64 = EPS and 2^96 = -1 (mod p), so x = lo + hi_lo*EPS - hi_hi; each wrap is fixed by +-EPS.
 * Canonicalised at the end.  tests/test_oracle_kat.py checks it against orc_red128_slow. */
static inline uint64_t orc_red128(orc_u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & 0xFFFFFFFFULL;
    /* the two wrap corrections and the final canonicalisation as masks: the wraps are data-dependent coin flips, and a
     * mispredicted branch costs more than the whole reduction */
    uint64_t t0 = lo - hh;
    t0 -= 0xFFFFFFFFULL & (0 - (uint64_t)(lo < hh));
    uint64_t t1 = hl * 0xFFFFFFFFULL;
    uint64_t t2 = t0 + t1;
    t2 += 0xFFFFFFFFULL & (0 - (uint64_t)(t2 < t1));
    return t2 - (ORC_P & (0 - (uint64_t)(t2 >= ORC_P)));
}
/* a, b canonical */
static inline uint64_t orc_add(uint64_t a, uint64_t b) { orc_u128 s = (orc_u128)a + b; return (uint64_t)(s >= ORC_P ? s - ORC_P : s); }
/* a, b canonical */
static inline uint64_t orc_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (ORC_P - b); }
static inline uint64_t orc_neg(uint64_t a) { return orc_sub(0, a); }
static inline uint64_t orc_mul(uint64_t a, uint64_t b) { return orc_red128((orc_u128)a * b); }
/* r = a*b + c  (arithmetic_chip.rs:98-107) */
static inline uint64_t orc_mul_add(uint64_t a, uint64_t b, uint64_t c) { return orc_red128((orc_u128)a * b + c); }

static inline uint64_t orc_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = orc_mul(r, b); b = orc_mul(b, b); e >>= 1; }
    return r;
}
/* Field::inverse for a != 0 (Fermat). */
static inline uint64_t orc_inv(uint64_t a) { return orc_pow(a, ORC_P - 2); }

/* ---- quadratic extension, X^2 = 7 ---- */
static inline orc_fp2 orc2(uint64_t a, uint64_t b) { orc_fp2 r = {{a, b}}; return r; }
static inline orc_fp2 orc2_add(orc_fp2 a, orc_fp2 b) { return orc2(orc_add(a.c[0], b.c[0]), orc_add(a.c[1], b.c[1])); }
static inline orc_fp2 orc2_sub(orc_fp2 a, orc_fp2 b) { return orc2(orc_sub(a.c[0], b.c[0]), orc_sub(a.c[1], b.c[1])); }
static inline orc_fp2 orc2_mul(orc_fp2 a, orc_fp2 b) {
    uint64_t c0 = orc_add(orc_mul(a.c[0], b.c[0]), orc_mul(7, orc_mul(a.c[1], b.c[1])));
    uint64_t c1 = orc_add(orc_mul(a.c[0], b.c[1]), orc_mul(a.c[1], b.c[0]));
    return orc2(c0, c1);
}
/* mul_add_extension (goldilocks_extension_chip.rs:56-69) */
static inline orc_fp2 orc2_mul_add(orc_fp2 a, orc_fp2 b, orc_fp2 c) { return orc2_add(orc2_mul(a, b), c); }
static inline int orc2_is_zero(orc_fp2 a) { return a.c[0] == 0 && a.c[1] == 0; }
static inline int orc2_eq(orc_fp2 a, orc_fp2 b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1]; }
/* QuadraticExtension::inverse (witnessed at goldilocks_extension_chip.rs:83-96):
 * (a0 + a1 X)^-1 = (a0 - a1 X) / (a0^2 - 7 a1^2).  Caller guarantees a != 0. */
static inline orc_fp2 orc2_inv(orc_fp2 a) {
    uint64_t norm = orc_sub(orc_mul(a.c[0], a.c[0]), orc_mul(7, orc_mul(a.c[1], a.c[1])));
    uint64_t ni = orc_inv(norm);
    return orc2(orc_mul(a.c[0], ni), orc_mul(orc_neg(a.c[1]), ni));
}
#endif

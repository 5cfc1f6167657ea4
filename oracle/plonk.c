/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Statement-by-statement restatement of the reference's plonk-level verifier checks, materialising the same vectors
 * the reference builds (paths relative to src/plonky2_verifier/):
 *   verify_proof_with_challenges  chip/plonk/plonk_verifier_chip.rs:156-210
 *   eval_vanishing_poly           chip/plonk/vanishing_poly.rs:18-124
 *   eval_gate_constraints         vanishing_poly.rs:126-154
 *   eval_l_0_x                    vanishing_poly.rs:156-181
 *   check_partial_products        vanishing_poly.rs:183-218
 *   eval_filtered_constraint      chip/plonk/gates/mod.rs:86-134
 *   reduce_extension              chip/goldilocks_extension_chip.rs:331-342 (terms.rev().fold(0, acc*base + term))
 *   gates: noop.rs, constant.rs:18-37, public_input.rs:22-40, arithmetic.rs:38-72, arithmetic_extension.rs:40-84,
 *          multiplication_extension.rs:34-71, base_sum.rs:29-62, reducing.rs:54-86, reducing_extension.rs:57-88,
 *          random_access.rs:84-166, poseidon_mds.rs:35-125, poseidon.rs:324-700
 *   extension algebra             chip/goldilocks_extension_algebra_chip.rs:34-171
 * Parity unpinned against a real plonky2 proof (none can be produced here); pinned instead by an independent
 * pure-Python prover (tests/plonk_prover.py) whose proofs this restatement must accept. */
#include "oracle.h"
#include "orc_field.h"
#include "poseidon_g_constants.h"
#include <stdlib.h>
#include <string.h>

#define UNUSED_SELECTOR 0xFFFFFFFFull /* gates/mod.rs:30 */

static orc_fp2 at(const uint64_t *v, size_t i) { return orc2(v[2 * i], v[2 * i + 1]); }
static orc_fp2 lift(uint64_t a) { return orc2(a, 0); }
static orc_fp2 scalar_mul(orc_fp2 a, uint64_t s) { return orc2(orc_mul(a.c[0], s), orc_mul(a.c[1], s)); }

/* ExtensionAlgebra: chip/goldilocks_extension_algebra_chip.rs */
typedef struct { orc_fp2 e[2]; } ext_alg;
static ext_alg get_local_ext_algebra(const uint64_t *local_wires, size_t start) { /* gates/mod.rs:50-58 */
    ext_alg r = {{at(local_wires, start), at(local_wires, start + 1)}};
    return r;
}
static ext_alg zero_ext_algebra(void) { ext_alg r = {{orc2(0, 0), orc2(0, 0)}}; return r; }
static ext_alg convert_to_ext_algebra(orc_fp2 et) { ext_alg r = {{et, orc2(0, 0)}}; return r; }
/* inner_product_extension :59-82: acc += constant * a * b over the pairs */
static ext_alg mul_add_ext_algebra(ext_alg a, ext_alg b, ext_alg c) { /* :112-147 */
    ext_alg res;
    for (int out = 0; out < 2; out++) {
        orc_fp2 acc = c.e[out];
        for (int i = 0; i < 2; i++)                 /* inner_w: pairs with i + j >= 2, constant w = 7 */
            for (int j = 2 - i; j < 2; j++)
                if ((i + j) % 2 == out) acc = orc2_add(scalar_mul(orc2_mul(a.e[i], b.e[j]), 7), acc);
        for (int i = 0; i < 2; i++)                 /* inner: pairs with i + j < 2, constant 1 */
            for (int j = 0; j < 2 - i; j++)
                if ((i + j) % 2 == out) acc = orc2_add(orc2_mul(a.e[i], b.e[j]), acc);
        res.e[out] = acc;
    }
    return res;
}
static ext_alg mul_ext_algebra(ext_alg a, ext_alg b) { return mul_add_ext_algebra(a, b, zero_ext_algebra()); }
static ext_alg scalar_mul_add_ext_algebra(orc_fp2 a, ext_alg b, ext_alg c) { /* :85-99 */
    ext_alg r;
    for (int i = 0; i < 2; i++) r.e[i] = orc2_mul_add(a, b.e[i], c.e[i]);
    return r;
}
static ext_alg scalar_mul_ext_algebra(orc_fp2 a, ext_alg b) { return scalar_mul_add_ext_algebra(a, b, zero_ext_algebra()); }
static ext_alg sub_ext_algebra(ext_alg a, ext_alg b) {
    ext_alg r;
    for (int i = 0; i < 2; i++) r.e[i] = orc2_sub(a.e[i], b.e[i]);
    return r;
}

/* ---- PoseidonGateConstrainer helpers, gates/poseidon.rs:383-585 ---- */
#define T 12
static void constant_layer(orc_fp2 *state, size_t round_ctr) {
    for (int i = 0; i < T; i++) state[i] = orc2_add(state[i], lift(ORC_ALL_ROUND_CONSTANTS[i + T * round_ctr]));
}
static void partial_first_constant_layer(orc_fp2 *state) {
    for (int i = 0; i < T; i++) state[i] = orc2_add(state[i], lift(ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
}
static orc_fp2 sbox(orc_fp2 x) { /* exp(element, 7) */
    orc_fp2 r = x;
    for (int i = 0; i < 6; i++) r = orc2_mul(r, x);
    return r;
}
static void sbox_layer(orc_fp2 *state) { for (int i = 0; i < T; i++) state[i] = sbox(state[i]); }
static orc_fp2 mds_row_shf(int row, const orc_fp2 *state) {
    orc_fp2 res = orc2(0, 0);
    for (int i = 0; i < T; i++) res = orc2_mul_add(lift(ORC_MDS_MATRIX_CIRC[i]), state[(i + row) % T], res);
    return orc2_mul_add(lift(ORC_MDS_MATRIX_DIAG[row]), state[row], res);
}
static void mds_layer(orc_fp2 *state) {
    orc_fp2 result[T];
    for (int i = 0; i < T; i++) result[i] = mds_row_shf(i, state);
    memcpy(state, result, sizeof result);
}
static void mds_partial_layer_init(orc_fp2 *state) {
    orc_fp2 result[T];
    for (int i = 0; i < T; i++) result[i] = orc2(0, 0);
    result[0] = state[0];
    for (int r = 1; r < T; r++)
        for (int c = 1; c < T; c++)
            result[c] = orc2_mul_add(lift(ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)]), state[r], result[c]);
    memcpy(state, result, sizeof result);
}
static void mds_partial_layer_fast(orc_fp2 *state, int r) {
    orc_fp2 s0 = state[0];
    uint64_t mds0to0 = ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0];
    orc_fp2 d = scalar_mul(s0, mds0to0);
    for (int i = 1; i < T; i++) d = orc2_mul_add(lift(ORC_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1]), state[i], d);
    orc_fp2 result[T];
    result[0] = d;
    for (int i = 1; i < T; i++) result[i] = orc2_mul_add(lift(ORC_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1]), state[0], state[i]);
    memcpy(state, result, sizeof result);
}
/* eval_unfiltered_constraint of the PoseidonGate, gates/poseidon.rs:588-699; returns the number of constraints */
static size_t poseidon_gate(const uint64_t *local_wires, orc_fp2 *constraints) {
    enum { R_F_HALF = 4, R_P = 22, WIRE_SWAP = 2 * T, START_DELTA = 2 * T + 1, START_FULL_0 = START_DELTA + 4,
           START_PARTIAL = START_FULL_0 + T * (R_F_HALF - 1), START_FULL_1 = START_PARTIAL + R_P };
    size_t n = 0;
    orc_fp2 swap = at(local_wires, WIRE_SWAP);
    constraints[n++] = orc2_sub(orc2_mul(swap, swap), swap);
    for (int i = 0; i < 4; i++) {
        orc_fp2 diff = orc2_sub(at(local_wires, i + 4), at(local_wires, i));
        constraints[n++] = orc2_sub(orc2_mul(swap, diff), at(local_wires, START_DELTA + i));
    }
    orc_fp2 state[T];
    for (int i = 0; i < 4; i++) {
        orc_fp2 delta_i = at(local_wires, START_DELTA + i);
        state[i] = orc2_add(at(local_wires, i), delta_i);
        state[i + 4] = orc2_sub(at(local_wires, i + 4), delta_i);
    }
    for (int i = 8; i < T; i++) state[i] = at(local_wires, i);
    size_t round_ctr = 0;
    for (int r = 0; r < R_F_HALF; r++) {
        constant_layer(state, round_ctr);
        if (r != 0)
            for (int i = 0; i < T; i++) {
                orc_fp2 sbox_in = at(local_wires, START_FULL_0 + T * (r - 1) + i);
                constraints[n++] = orc2_sub(state[i], sbox_in);
                state[i] = sbox_in;
            }
        sbox_layer(state);
        mds_layer(state);
        round_ctr++;
    }
    partial_first_constant_layer(state);
    mds_partial_layer_init(state);
    for (int r = 0; r < R_P - 1; r++) {
        orc_fp2 sbox_in = at(local_wires, START_PARTIAL + r);
        constraints[n++] = orc2_sub(state[0], sbox_in);
        state[0] = orc2_add(sbox(sbox_in), lift(ORC_FAST_PARTIAL_ROUND_CONSTANTS[r]));
        mds_partial_layer_fast(state, r);
    }
    {
        orc_fp2 sbox_in = at(local_wires, START_PARTIAL + R_P - 1);
        constraints[n++] = orc2_sub(state[0], sbox_in);
        state[0] = sbox(sbox_in);
        mds_partial_layer_fast(state, R_P - 1);
    }
    round_ctr += R_P;
    for (int r = 0; r < R_F_HALF; r++) {
        constant_layer(state, round_ctr);
        for (int i = 0; i < T; i++) {
            orc_fp2 sbox_in = at(local_wires, START_FULL_1 + T * r + i);
            constraints[n++] = orc2_sub(state[i], sbox_in);
            state[i] = sbox_in;
        }
        sbox_layer(state);
        mds_layer(state);
        round_ctr++;
    }
    for (int i = 0; i < T; i++) constraints[n++] = orc2_sub(state[i], at(local_wires, T + i));
    return n;
}

static orc_fp2 reduce_extension(orc_fp2 base, const orc_fp2 *terms, size_t n) {
    orc_fp2 acc = orc2(0, 0);
    for (size_t k = n; k-- > 0;) acc = orc2_add(orc2_mul(acc, base), terms[k]);
    return acc;
}

/* returns 1 iff accepted; <0 on a malformed circuit description */
int orc_plonk_check(const orc_plonk_circuit *C, const uint64_t *open0, const uint64_t *open1, const uint64_t pi_hash[4],
                    const uint64_t *chal, const uint64_t zeta_w[2]) {
    const orc_common *c = &C->common;
    const uint32_t nch = c->num_challenges, nr = c->num_routed_wires, qdf = c->quotient_degree_factor,
                   npp = c->num_partial_products;
    if (nch == 0 || qdf == 0 || (nr + qdf - 1) / qdf != npp + 1 || C->num_selectors == 0 || C->num_selectors > c->num_constants) return -1;
    /* OpeningSetValues in FRI batch order (types/assigned.rs:26-40) */
    const uint64_t *local_constants = open0;
    const uint64_t *s_sigmas = local_constants + 2 * (size_t)c->num_constants;
    const uint64_t *local_wires = s_sigmas + 2 * (size_t)nr;
    const uint64_t *local_zs = local_wires + 2 * (size_t)c->num_wires;
    const uint64_t *partial_products = local_zs + 2 * (size_t)nch;
    const uint64_t *quotient_polys = partial_products + 2 * (size_t)nch * npp;
    const uint64_t *next_zs = open1;
    const uint64_t *betas = chal, *gammas = chal + nch, *alphas = chal + 2 * nch;
    /* range checks of assign_value (native_chip/arithmetic_chip.rs:256-268) */
    size_t n0 = 2 * (size_t)(c->num_constants + nr + c->num_wires + nch + nch * npp + nch * qdf);
    for (size_t k = 0; k < n0; k++) if (open0[k] >= ORC_P) return 0;
    for (size_t k = 0; k < 2 * (size_t)nch; k++) if (open1[k] >= ORC_P) return 0;
    for (size_t k = 0; k < 3 * (size_t)nch; k++) if (chal[k] >= ORC_P) return 0;
    for (int k = 0; k < 4; k++) if (pi_hash[k] >= ORC_P) return 0;
    if (zeta_w[0] >= ORC_P || zeta_w[1] >= ORC_P) return 0;
    const orc_fp2 x = orc2(zeta_w[0], zeta_w[1]), one = orc2(1, 0);

    /* plonk_verifier_chip.rs:174-178: exp_power_of_2_extension(zeta, degree_bits) */
    orc_fp2 x_pow_deg = x;
    for (uint32_t i = 0; i < C->degree_bits; i++) x_pow_deg = orc2_mul(x_pow_deg, x_pow_deg);

    /* ---- eval_gate_constraints ---- */
    size_t ngc = C->num_gate_constraints;
    orc_fp2 *constraint_terms = calloc(ngc ? ngc : 1, sizeof(orc_fp2));
    for (uint32_t i = 0; i < C->num_gates; i++) {
        uint32_t selector_index = C->gates[i].selector_index;
        if (selector_index >= C->num_selectors) { free(constraint_terms); return -2; }
        orc_fp2 f_zeta = at(local_constants, selector_index);
        orc_fp2 filter = one;                                   /* mul_many_extension(terms) */
        for (uint32_t k = C->group_lo[selector_index]; k < C->group_hi[selector_index]; k++)
            if (k != i) filter = orc2_mul(filter, orc2_sub(lift(k), f_zeta));
        if (C->num_selectors > 1) filter = orc2_mul(filter, orc2_sub(lift(UNUSED_SELECTOR), f_zeta));
        const uint64_t *gate_constants = local_constants + 2 * (size_t)C->num_selectors;   /* gates/mod.rs:122 */
        orc_fp2 gc[128];
        size_t n_gc = 0;
        switch (C->gates[i].kind) {
            case 0: break;                                       /* NoopGate */
            case 1:                                              /* ConstantGate: constants[i] - wires[i] */
                for (uint32_t k = 0; k < C->gates[i].param; k++) gc[n_gc++] = orc2_sub(at(gate_constants, k), at(local_wires, k));
                break;
            case 2:                                              /* PublicInputGate: wires[0..4] - hash */
                for (uint32_t k = 0; k < 4; k++) gc[n_gc++] = orc2_sub(at(local_wires, k), lift(pi_hash[k]));
                break;
            case 3: {                                            /* ArithmeticGate */
                orc_fp2 const_0 = at(gate_constants, 0), const_1 = at(gate_constants, 1);
                for (uint32_t k = 0; k < C->gates[i].param; k++) {
                    orc_fp2 term1 = orc2_mul(orc2_mul(at(local_wires, 4 * k), at(local_wires, 4 * k + 1)), const_0);
                    orc_fp2 term2 = orc2_mul(at(local_wires, 4 * k + 2), const_1);
                    gc[n_gc++] = orc2_sub(at(local_wires, 4 * k + 3), orc2_add(term1, term2));
                }
                break;
            }
            case 4: {                                            /* ArithmeticExtensionGate, arithmetic_extension.rs:40-84 */
                orc_fp2 const_0 = at(gate_constants, 0), const_1 = at(gate_constants, 1);
                for (uint32_t k = 0; k < C->gates[i].param; k++) {
                    ext_alg multiplicand_0 = get_local_ext_algebra(local_wires, 8 * k);
                    ext_alg multiplicand_1 = get_local_ext_algebra(local_wires, 8 * k + 2);
                    ext_alg addend = get_local_ext_algebra(local_wires, 8 * k + 4);
                    ext_alg output = get_local_ext_algebra(local_wires, 8 * k + 6);
                    ext_alg scaled_mul = scalar_mul_ext_algebra(const_0, mul_ext_algebra(multiplicand_0, multiplicand_1));
                    ext_alg diff = sub_ext_algebra(output, scalar_mul_add_ext_algebra(const_1, addend, scaled_mul));
                    gc[n_gc++] = diff.e[0]; gc[n_gc++] = diff.e[1];
                }
                break;
            }
            case 5: {                                            /* MulExtensionGate, multiplication_extension.rs:34-71 */
                orc_fp2 const_0 = at(gate_constants, 0);
                for (uint32_t k = 0; k < C->gates[i].param; k++) {
                    ext_alg mul = mul_ext_algebra(get_local_ext_algebra(local_wires, 6 * k), get_local_ext_algebra(local_wires, 6 * k + 2));
                    ext_alg diff = sub_ext_algebra(get_local_ext_algebra(local_wires, 6 * k + 4), scalar_mul_ext_algebra(const_0, mul));
                    gc[n_gc++] = diff.e[0]; gc[n_gc++] = diff.e[1];
                }
                break;
            }
            case 6: {                                            /* BaseSumGate<2>, base_sum.rs:29-62 */
                uint32_t num_limbs = C->gates[i].param;
                orc_fp2 limbs[127];
                if (num_limbs > 127) { free(constraint_terms); return -3; }
                for (uint32_t k = 0; k < num_limbs; k++) limbs[k] = at(local_wires, 1 + k);
                orc_fp2 computed_sum = reduce_extension(lift(2), limbs, num_limbs);
                gc[n_gc++] = orc2_sub(computed_sum, at(local_wires, 0));
                for (uint32_t k = 0; k < num_limbs; k++) {
                    orc_fp2 acc = one;
                    for (uint64_t b = 0; b < 2; b++)             /* acc' = acc * limb + (-b) * acc */
                        acc = orc2_add(orc2_mul(acc, limbs[k]), scalar_mul(acc, orc_neg(b)));
                    gc[n_gc++] = acc;
                }
                break;
            }
            case 7: case 8: {                                    /* ReducingGate reducing.rs:54-86, ReducingExtensionGate reducing_extension.rs:57-88 */
                uint32_t num_coeffs = C->gates[i].param;
                int is_ext = C->gates[i].kind == 8;
                size_t start_accs = 6 + (is_ext ? 2 * (size_t)num_coeffs : num_coeffs);
                ext_alg alpha = get_local_ext_algebra(local_wires, 2), acc = get_local_ext_algebra(local_wires, 4);
                if (2 * (size_t)num_coeffs > 128) { free(constraint_terms); return -3; }
                for (uint32_t k = 0; k < num_coeffs; k++) {
                    ext_alg coeff = is_ext ? get_local_ext_algebra(local_wires, 6 + 2 * k) : convert_to_ext_algebra(at(local_wires, 6 + k));
                    ext_alg accs_k = get_local_ext_algebra(local_wires, k == num_coeffs - 1 ? 0 : start_accs + 2 * k);
                    ext_alg tmp = sub_ext_algebra(mul_add_ext_algebra(acc, alpha, coeff), accs_k);
                    gc[n_gc++] = tmp.e[0]; gc[n_gc++] = tmp.e[1];
                    acc = accs_k;
                }
                break;
            }
            case 9: {                                            /* RandomAccessGate, random_access.rs:84-166 */
                uint32_t bits = C->gates[i].param, num_copies = C->gates[i].param2, num_extra_constants = C->gates[i].param3;
                size_t vec_size = (size_t)1 << bits, num_routed = (2 + vec_size) * num_copies + num_extra_constants;
                if (bits > 6 || num_copies * (bits + 2) + num_extra_constants > 128) { free(constraint_terms); return -3; }
                for (uint32_t copy = 0; copy < num_copies; copy++) {
                    orc_fp2 access_index = at(local_wires, (2 + vec_size) * copy);
                    orc_fp2 claimed_element = at(local_wires, (2 + vec_size) * copy + 1);
                    orc_fp2 list_items[64], bitv[6];
                    for (size_t k = 0; k < vec_size; k++) list_items[k] = at(local_wires, (2 + vec_size) * copy + 2 + k);
                    for (uint32_t k = 0; k < bits; k++) bitv[k] = at(local_wires, num_routed + copy * bits + k);
                    for (uint32_t k = 0; k < bits; k++) gc[n_gc++] = orc2_sub(orc2_mul(bitv[k], bitv[k]), bitv[k]);
                    gc[n_gc++] = orc2_sub(reduce_extension(lift(2), bitv, bits), access_index);
                    size_t len = vec_size;
                    for (uint32_t k = 0; k < bits; k++) {        /* tuples().map(|(x, y)| select(b, y, x)) */
                        for (size_t j = 0; j < len / 2; j++) {
                            orc_fp2 x = list_items[2 * j], y = list_items[2 * j + 1];
                            list_items[j] = orc2_add(orc2_mul(bitv[k], orc2_sub(y, x)), x);
                        }
                        len /= 2;
                    }
                    gc[n_gc++] = orc2_sub(list_items[0], claimed_element);
                }
                for (uint32_t k = 0; k < num_extra_constants; k++)
                    gc[n_gc++] = orc2_sub(at(gate_constants, k), at(local_wires, (2 + vec_size) * num_copies + k));
                break;
            }
            case 10: {                                           /* PoseidonMdsGate, poseidon_mds.rs:35-125 */
                for (int row = 0; row < T; row++) {
                    ext_alg res = zero_ext_algebra();
                    for (int k = 0; k < T; k++)
                        res = scalar_mul_add_ext_algebra(lift(ORC_MDS_MATRIX_CIRC[k]), get_local_ext_algebra(local_wires, 2 * (size_t)((k + row) % T)), res);
                    res = scalar_mul_add_ext_algebra(lift(ORC_MDS_MATRIX_DIAG[row]), get_local_ext_algebra(local_wires, 2 * (size_t)row), res);
                    ext_alg diff = sub_ext_algebra(get_local_ext_algebra(local_wires, 2 * (size_t)(T + row)), res);
                    gc[n_gc++] = diff.e[0]; gc[n_gc++] = diff.e[1];
                }
                break;
            }
            case 11: n_gc = poseidon_gate(local_wires, gc); break;  /* PoseidonGate */
            default: free(constraint_terms); return -3;         /* unimplemented!() in the reference for unknown ids */
        }
        if (n_gc > ngc) { free(constraint_terms); return -4; }
        for (size_t k = 0; k < n_gc; k++) constraint_terms[k] = orc2_mul_add(filter, gc[k], constraint_terms[k]);
    }

    /* ---- eval_vanishing_poly ---- */
    size_t n_terms = nch + (size_t)nch * (npp + 1) + ngc, nt = 0;
    orc_fp2 *vanishing_terms = calloc(n_terms, sizeof(orc_fp2));
    /* eval_l_0_x: (x^n - 1) / (n*x - n) */
    uint64_t n_f = (uint64_t)1 << C->degree_bits;
    orc_fp2 zero_poly = orc2_sub(x_pow_deg, one);
    orc_fp2 denominator = orc2_add(scalar_mul(x, n_f % ORC_P), lift(orc_neg(n_f % ORC_P)));
    int ok = 1;
    if (orc2_is_zero(denominator)) ok = 0;                       /* div_extension by zero */
    orc_fp2 l_0_x = ok ? orc2_mul(zero_poly, orc2_inv(denominator)) : orc2(0, 0);
    for (uint32_t i = 0; i < nch; i++)                           /* vanishing_z_1_terms */
        vanishing_terms[nt++] = orc2_sub(orc2_mul(l_0_x, at(local_zs, i)), l_0_x);
    orc_fp2 *numerator_values = calloc(nr, sizeof(orc_fp2)), *denominator_values = calloc(nr, sizeof(orc_fp2));
    for (uint32_t i = 0; i < nch; i++) {
        orc_fp2 beta = lift(betas[i]), gamma = lift(gammas[i]);
        for (uint32_t j = 0; j < nr; j++) {
            orc_fp2 s_id = scalar_mul(x, C->k_is[j]);
            orc_fp2 wire_value_plus_gamma = orc2_add(at(local_wires, j), gamma);
            numerator_values[j] = orc2_mul_add(beta, s_id, wire_value_plus_gamma);
            denominator_values[j] = orc2_mul_add(beta, at(s_sigmas, j), wire_value_plus_gamma);
        }
        /* check_partial_products: product_accs = [z_x, partials.., z_gx], chunk_size = max_degree */
        for (uint32_t w = 0, c0 = 0; c0 < nr; c0 += qdf, w++) {
            orc_fp2 nume_product = one, denom_product = one;
            for (uint32_t j = c0; j < nr && j < c0 + qdf; j++) {
                nume_product = orc2_mul(nume_product, numerator_values[j]);
                denom_product = orc2_mul(denom_product, denominator_values[j]);
            }
            orc_fp2 prev_acc = w == 0 ? at(local_zs, i) : at(partial_products, (size_t)i * npp + w - 1);
            orc_fp2 next_acc = w == npp ? at(next_zs, i) : at(partial_products, (size_t)i * npp + w);
            vanishing_terms[nt++] = orc2_sub(orc2_mul(prev_acc, nume_product), orc2_mul(next_acc, denom_product));
        }
    }
    for (size_t k = 0; k < ngc; k++) vanishing_terms[nt++] = constraint_terms[k];

    /* plonk_verifier_chip.rs:194-209 */
    orc_fp2 z_h_zeta = orc2_sub(x_pow_deg, one);
    for (uint32_t i = 0; i < nch && ok; i++) {
        orc_fp2 vanishing_poly_zeta = reduce_extension(lift(alphas[i]), vanishing_terms, nt);
        orc_fp2 chunk[64];
        if (qdf > 64) { ok = 0; break; }
        for (uint32_t j = 0; j < qdf; j++) chunk[j] = at(quotient_polys, (size_t)i * qdf + j);
        orc_fp2 recombined_quotient = reduce_extension(x_pow_deg, chunk, qdf);
        if (!orc2_eq(vanishing_poly_zeta, orc2_mul(z_h_zeta, recombined_quotient))) ok = 0;
    }
    free(numerator_values); free(denominator_values); free(vanishing_terms); free(constraint_terms);
    return ok;
}

// The O(1)-per-proof plonk checks of the verifier (second half): the vanishing-polynomial identity at
// zeta.  One function for host and device, like wire.hpp: the kernel in plonk_kernels.cuh runs one thread per proof,
// the host entry point sv_plonk_check_host runs the same code for the CPU tests and for callers with a handful of proofs.
//
// Reference restated (paths relative to src/plonky2_verifier/):
//   PlonkVerifierChip::verify_proof_with_challenges .... chip/plonk/plonk_verifier_chip.rs:156-210
//     zeta^n, Z_H(zeta) = zeta^n - 1, per challenge i: vanishing_i(zeta) == Z_H(zeta) * sum_j quotient[i*qdf + j] * zeta^(n*j)
//   eval_vanishing_poly ................................ chip/plonk/vanishing_poly.rs:18-124
//     terms = [L_0(zeta)(Z_i(zeta) - 1)]_i ++ [partial-product checks]_i ++ gate constraints; vanishing_i = sum_k alpha_i^k term_k
//   eval_l_0_x ......................................... vanishing_poly.rs:156-181      L_0(x) = (x^n - 1) / (n (x - 1))
//   check_partial_products ............................. vanishing_poly.rs:183-218
//   eval_gate_constraints / eval_filtered_constraint ... vanishing_poly.rs:126-154, chip/plonk/gates/mod.rs:86-134
//     filter_i = prod_{k in group(i), k != i} (k - s(zeta)) * [num_selectors > 1] (UNUSED_SELECTOR - s(zeta))
//   gates: NoopGate gates/noop.rs, ConstantGate gates/constant.rs:18-37, PublicInputGate gates/public_input.rs:22-40,
//          ArithmeticGate gates/arithmetic.rs:38-72, ArithmeticExtensionGate gates/arithmetic_extension.rs:21-84,
//          MulExtensionGate gates/multiplication_extension.rs:21-71, BaseSumGate (base 2) gates/base_sum.rs:18-62,
//          ReducingGate gates/reducing.rs:19-86, ReducingExtensionGate gates/reducing_extension.rs:19-88; the extension
//          algebra they compute in: chip/goldilocks_extension_algebra_chip.rs:34-171; RandomAccessGate
//          gates/random_access.rs:84-166, PoseidonMdsGate gates/poseidon_mds.rs:35-125, PoseidonGate gates/poseidon.rs:324-700
//          (the permutation in its fast form with every S-box input a wire: 135 wires, 123 constraints).
// That is every gate the reference knows (gates/mod.rs:138-196); any other kind is refused by sv_plonk_circuit_check
// (an error, never a silent accept).
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"
#include "poseidon_g.cuh"   // the Poseidon-Goldilocks tables (SVB_T), shared with the hash itself

namespace svb {

SVB_HD fp2 ext_at(const u64* v, u32 i) { return mk2(v[2 * i], v[2 * i + 1]); }
SVB_HD fp2 lift(u64 a) { return mk2(a, 0); }

// ExtensionAlgebra (Fp2 (x) Fp2, chip/goldilocks_extension_algebra_chip.rs): a pair of Fp2 values standing for the two
// base-field limbs of a D = 2 wire element, each evaluated at zeta.
struct alg2 { fp2 a, b; };
SVB_HD alg2 alg_at(const u64* wires, u32 first) { alg2 r; r.a = ext_at(wires, first); r.b = ext_at(wires, first + 1); return r; }
SVB_HD alg2 alg_lift(fp2 x) { alg2 r; r.a = x; r.b = mk2(0, 0); return r; }                  // convert_to_ext_algebra :46-57
SVB_HD alg2 alg_sub(alg2 x, alg2 y) { alg2 r; r.a = sub2(x.a, y.a); r.b = sub2(x.b, y.b); return r; }
// mul_add_ext_algebra :112-147: (x.a y.a + W x.b y.b + c.a, x.a y.b + x.b y.a + c.b), W = 7
SVB_HD alg2 alg_mul_add(alg2 x, alg2 y, alg2 c) {
    alg2 r;
    r.a = add2(add2(c.a, scale2(mul2(x.b, y.b), 7)), mul2(x.a, y.a));
    r.b = add2(add2(c.b, mul2(x.a, y.b)), mul2(x.b, y.a));
    return r;
}
// scalar_mul_add_ext_algebra :85-99: s * y + c with s in Fp2
SVB_HD alg2 alg_scalar_mul_add(fp2 s, alg2 y, alg2 c) { alg2 r; r.a = add2(mul2(s, y.a), c.a); r.b = add2(mul2(s, y.b), c.b); return r; }

// wires used / constraints produced by a gate (0 wires = unknown kind)
SVB_HD void plonk_gate_dims(u32 kind, u32 param, u32 param2, u32 param3, u32& wires, u32& constraints, u32& constants) {
    wires = constraints = constants = 0;
    // A gate parameter comes straight from a gate-id string (sv_plonk_gate_from_id): bound it BEFORE the u32 products below
    // can wrap (ArithmeticExtensionGate { num_ops: 0x80000001 } would otherwise look like 8 wires / 2 constraints and then
    // run its loops 2^31 times over the openings).  No gate of the reference has more than 135 wires; 0 wires = refused.
    if (param > SV_MAX_GATE_CONSTRAINTS || param2 > 64 || param3 > 64) return;
    switch (kind) {
        case SV_GATE_NOOP: wires = 1; break;
        case SV_GATE_CONSTANT: wires = param; constraints = param; constants = param; break;
        case SV_GATE_PUBLIC_INPUT: wires = 4; constraints = 4; break;
        case SV_GATE_ARITHMETIC: wires = 4 * param; constraints = param; constants = 2; break;
        case SV_GATE_ARITHMETIC_EXT: wires = 8 * param; constraints = 2 * param; constants = 2; break;
        case SV_GATE_MUL_EXT: wires = 6 * param; constraints = 2 * param; constants = 1; break;
        case SV_GATE_BASE_SUM: wires = 1 + param; constraints = 1 + param; break;
        case SV_GATE_REDUCING: wires = param ? 3 * param + 4 : 0; constraints = 2 * param; break;
        case SV_GATE_REDUCING_EXT: wires = param ? 4 * param + 4 : 0; constraints = 2 * param; break;
        case SV_GATE_RANDOM_ACCESS:
            if (param == 0 || param > 6 || param2 == 0 || param2 > 64 || param3 > 16) break;
            wires = (2 + (1u << param)) * param2 + param3 + param * param2;
            constraints = param2 * (param + 2) + param3;
            constants = param3;
            break;
        case SV_GATE_POSEIDON_MDS: wires = 48; constraints = 24; break;
        case SV_GATE_POSEIDON: wires = 135; constraints = 123; break;
        default: break;
    }
}

// ---- PoseidonGate (gates/poseidon.rs): the permutation over Fp2 values, fast form ----------------------------
SVB_HD fp2 pg_sbox(fp2 x) {                       // exp(element, 7), :420-429
    const fp2 x2 = mul2(x, x), x4 = mul2(x2, x2);
    return mul2(mul2(x2, x), x4);
}
SVB_HD void pg_mds_layer(fp2 st[12]) {            // mds_row_shf / mds_layer :443-502
    fp2 r[12];
    for (int row = 0; row < 12; row++) {
        fp2 acc = mk2(0, 0);
        for (int i = 0; i < 12; i++) acc = add2(acc, scale2(st[(i + row) % 12], SVB_T(MDS_MATRIX_CIRC)[i]));
        r[row] = add2(acc, scale2(st[row], SVB_T(MDS_MATRIX_DIAG)[row]));
    }
    for (int i = 0; i < 12; i++) st[i] = r[i];
}
// Adds filter * constraint_k to gate_c[k] for the 123 constraints of the gate, in the reference's order (:596-699).
SVB_HD void poseidon_gate_constraints(const u64* wires, fp2 filter, fp2* gate_c) {
    u32 k = 0;
    auto emit = [&](fp2 c) { gate_c[k] = add2(gate_c[k], mul2(filter, c)); k++; };
    const u32 WIRE_SWAP = 24, START_DELTA = 25, START_FULL_0 = 29, START_PARTIAL = 29 + 36, START_FULL_1 = 29 + 36 + 22;
    const fp2 swap = ext_at(wires, WIRE_SWAP);
    emit(sub2(mul2(swap, swap), swap));                                   // swap is binary
    for (u32 i = 0; i < 4; i++)                                           // delta_i = swap * (rhs - lhs)
        emit(sub2(mul2(swap, sub2(ext_at(wires, i + 4), ext_at(wires, i))), ext_at(wires, START_DELTA + i)));
    fp2 st[12];
    for (u32 i = 0; i < 4; i++) {
        const fp2 d = ext_at(wires, START_DELTA + i);
        st[i] = add2(ext_at(wires, i), d);
        st[i + 4] = sub2(ext_at(wires, i + 4), d);
    }
    for (u32 i = 8; i < 12; i++) st[i] = ext_at(wires, i);
    u32 round_ctr = 0;
    for (u32 r = 0; r < 4; r++) {                                         // first set of full rounds
        for (u32 i = 0; i < 12; i++) st[i] = add2(st[i], lift(SVB_T(ALL_ROUND_CONSTANTS)[i + 12 * round_ctr]));
        if (r != 0)
            for (u32 i = 0; i < 12; i++) {
                const fp2 sbox_in = ext_at(wires, START_FULL_0 + 12 * (r - 1) + i);
                emit(sub2(st[i], sbox_in));
                st[i] = sbox_in;
            }
        for (u32 i = 0; i < 12; i++) st[i] = pg_sbox(st[i]);
        pg_mds_layer(st);
        round_ctr++;
    }
    for (u32 i = 0; i < 12; i++) st[i] = add2(st[i], lift(SVB_T(FAST_PARTIAL_FIRST_ROUND_CONSTANT)[i]));
    {                                                                      // mds_partial_layer_init :503-535
        fp2 res[12];
        for (u32 c = 0; c < 12; c++) res[c] = mk2(0, 0);
        res[0] = st[0];
        for (u32 r = 1; r < 12; r++)
            for (u32 c = 1; c < 12; c++)
                res[c] = add2(res[c], scale2(st[r], SVB_T(FAST_PARTIAL_ROUND_INITIAL_MATRIX)[(r - 1) * 11 + (c - 1)]));
        for (u32 c = 0; c < 12; c++) st[c] = res[c];
    }
    for (u32 r = 0; r < 22; r++) {                                        // partial rounds :651-670
        const fp2 sbox_in = ext_at(wires, START_PARTIAL + r);
        emit(sub2(st[0], sbox_in));
        st[0] = pg_sbox(sbox_in);
        if (r != 21) st[0] = add2(st[0], lift(SVB_T(FAST_PARTIAL_ROUND_CONSTANTS)[r]));
        // mds_partial_layer_fast :537-585
        fp2 d = scale2(st[0], SVB_T(MDS_MATRIX_CIRC)[0] + SVB_T(MDS_MATRIX_DIAG)[0]);
        for (u32 i = 1; i < 12; i++) d = add2(d, scale2(st[i], SVB_T(FAST_PARTIAL_ROUND_W_HATS)[r * 11 + i - 1]));
        for (u32 i = 1; i < 12; i++) st[i] = add2(scale2(st[0], SVB_T(FAST_PARTIAL_ROUND_VS)[r * 11 + i - 1]), st[i]);
        st[0] = d;
    }
    round_ctr += 22;
    for (u32 r = 0; r < 4; r++) {                                         // second set of full rounds
        for (u32 i = 0; i < 12; i++) st[i] = add2(st[i], lift(SVB_T(ALL_ROUND_CONSTANTS)[i + 12 * round_ctr]));
        for (u32 i = 0; i < 12; i++) {
            const fp2 sbox_in = ext_at(wires, START_FULL_1 + 12 * r + i);
            emit(sub2(st[i], sbox_in));
            st[i] = pg_sbox(sbox_in);
        }
        pg_mds_layer(st);
        round_ctr++;
    }
    for (u32 i = 0; i < 12; i++) emit(sub2(st[i], ext_at(wires, 12 + i)));
}

// 0 = usable; < 0 = why not
static inline int plonk_circuit_check(const sv_plonk_circuit& C) {
    const sv_plonk_common& c = C.common;
    if (C.num_gates == 0 || C.num_gates > SV_MAX_GATES) return -1;
    if (C.num_selectors == 0 || C.num_selectors > SV_MAX_SELECTORS || C.num_selectors > c.num_constants) return -2;
    if (c.num_routed_wires == 0 || c.num_routed_wires > SV_MAX_ROUTED_WIRES || c.num_routed_wires > c.num_wires) return -3;
    if (c.num_challenges == 0 || c.num_challenges > SV_MAX_PLONK_CHALLENGES) return -4;
    if (c.quotient_degree_factor == 0) return -5;
    // check_partial_products zips the chunks with the windows of [z, partials.., z_next] (zip_eq): they must agree
    if ((c.num_routed_wires + c.quotient_degree_factor - 1) / c.quotient_degree_factor != c.num_partial_products + 1) return -6;
    if (C.num_gate_constraints > SV_MAX_GATE_CONSTRAINTS) return -7;
    if (C.degree_bits > 40) return -8;
    const u32 gate_consts = c.num_constants - C.num_selectors;
    for (u32 i = 0; i < C.num_gates; i++) {
        const sv_plonk_gate& g = C.gates[i];
        if (g.selector_index >= C.num_selectors) return -9;
        if (i < C.group_lo[g.selector_index] || i >= C.group_hi[g.selector_index] || C.group_hi[g.selector_index] > C.num_gates) return -10;
        u32 nwires, ncons, nconst;
        plonk_gate_dims(g.kind, g.param, g.param2, g.param3, nwires, ncons, nconst);
        if (nwires == 0) return -12;   // a gate this library does not evaluate yet, or an empty one
        if (nwires > c.num_wires || nconst > gate_consts) return -11;
        if (ncons > C.num_gate_constraints) return -13;
    }
    for (u32 j = 0; j < c.num_routed_wires; j++)
        if (!is_canonical(C.k_is[j])) return -14;
    return 0;
}

// open0: constants, sigmas, wires, zs, partial_products, quotient (FriOpenings batch 0, types/assigned.rs:26-37);
// open1: zs_next; all Fp2, canonical.  chal: betas[nch], gammas[nch], alphas[nch] (base field, canonical).
// Returns true iff the identity holds for every challenge.  The circuit must have passed plonk_circuit_check.
SVB_HD bool plonk_check_one(const sv_plonk_circuit& C, const u64* open0, const u64* open1, const u64 pi_hash[4],
                            const u64* chal, fp2 zeta) {
    const sv_plonk_common& c = C.common;
    const u32 nch = c.num_challenges, nr = c.num_routed_wires, qdf = c.quotient_degree_factor, npp = c.num_partial_products;
    const u64* constants = open0;
    const u64* sigmas = constants + 2 * c.num_constants;
    const u64* wires = sigmas + 2 * nr;
    const u64* zs = wires + 2 * c.num_wires;
    const u64* pps = zs + 2 * nch;
    const u64* quot = pps + 2 * nch * npp;
    const u64 *betas = chal, *gammas = chal + nch, *alphas = chal + 2 * nch;
    // every assigned value is range-checked by the reference (native_chip/arithmetic_chip.rs:256-268)
    {
        bool canon_ok = is_canonical(zeta.c0) && is_canonical(zeta.c1);
        const u32 n0_words = 2 * (c.num_constants + nr + c.num_wires + nch + nch * npp + nch * qdf);
        for (u32 k = 0; k < n0_words; k++) canon_ok = canon_ok && is_canonical(open0[k]);
        for (u32 k = 0; k < 2 * nch; k++) canon_ok = canon_ok && is_canonical(open1[k]);
        for (u32 k = 0; k < 3 * nch; k++) canon_ok = canon_ok && is_canonical(chal[k]);
        for (u32 k = 0; k < 4; k++) canon_ok = canon_ok && is_canonical(pi_hash[k]);
        if (!canon_ok) return false;
    }

    // plonk_verifier_chip.rs:174-178
    fp2 zeta_pow_deg = zeta;
    for (u32 i = 0; i < C.degree_bits; i++) zeta_pow_deg = mul2(zeta_pow_deg, zeta_pow_deg);
    const fp2 one = mk2(1, 0);
    const fp2 z_h = sub2(zeta_pow_deg, one);
    // eval_l_0_x: (x^n - 1) / (n x - n)
    const u64 n_f = (1ull << C.degree_bits) % GL_P;
    const fp2 l0_den = sub2(scale2(zeta, n_f), lift(n_f));
    if (is_zero2(l0_den)) return false;   // div_extension on zero (goldilocks_extension_chip.rs:72-101)
    const fp2 l_0 = mul2(z_h, inv2(l0_den));

    // sum_k alpha_i^k term_k, accumulated term by term in the order of `vanishing_terms` (vanishing_poly.rs:108-123)
    fp2 sum[SV_MAX_PLONK_CHALLENGES], apow[SV_MAX_PLONK_CHALLENGES];
    for (u32 i = 0; i < nch; i++) { sum[i] = mk2(0, 0); apow[i] = one; }
    auto push = [&](fp2 term) {
        for (u32 i = 0; i < nch; i++) {
            sum[i] = add2(sum[i], mul2(term, apow[i]));
            apow[i] = scale2(apow[i], alphas[i]);
        }
    };
    // vanishing_z_1_terms: L_0(x) Z(x) - L_0(x)   (:65-66)
    for (u32 i = 0; i < nch; i++) push(sub2(mul2(l_0, ext_at(zs, i)), l_0));
    // vanishing_partial_products_terms (:68-105, check_partial_products :183-218)
    for (u32 i = 0; i < nch; i++) {
        const fp2 beta = lift(betas[i]), gamma = lift(gammas[i]);
        fp2 prev = ext_at(zs, i);
        for (u32 ch0 = 0, w = 0; ch0 < nr; ch0 += qdf, w++) {
            fp2 nume = one, deno = one;
            for (u32 j = ch0; j < nr && j < ch0 + qdf; j++) {
                const fp2 wire_plus_gamma = add2(ext_at(wires, j), gamma);
                const fp2 s_id = scale2(zeta, C.k_is[j]);
                nume = mul2(nume, add2(mul2(beta, s_id), wire_plus_gamma));
                deno = mul2(deno, add2(mul2(beta, ext_at(sigmas, j)), wire_plus_gamma));
            }
            const fp2 next = w < npp ? ext_at(pps, i * npp + w) : ext_at(open1, i);
            push(sub2(mul2(prev, nume), mul2(next, deno)));
            prev = next;
        }
    }
    // eval_gate_constraints (:126-154)
    fp2 gate_c[SV_MAX_GATE_CONSTRAINTS];
    for (u32 k = 0; k < C.num_gate_constraints; k++) gate_c[k] = mk2(0, 0);
    const u64* gconst = constants + 2 * C.num_selectors;   // local_constants[num_selectors..] (gates/mod.rs:122)
    for (u32 i = 0; i < C.num_gates; i++) {
        const sv_plonk_gate& g = C.gates[i];
        const fp2 f_zeta = ext_at(constants, g.selector_index);
        fp2 filter = one;
        for (u32 k = C.group_lo[g.selector_index]; k < C.group_hi[g.selector_index]; k++)
            if (k != i) filter = mul2(filter, sub2(lift(k), f_zeta));
        if (C.num_selectors > 1) filter = mul2(filter, sub2(lift(0xFFFFFFFFull), f_zeta));   // UNUSED_SELECTOR = u32::MAX
        switch (g.kind) {
            case SV_GATE_CONSTANT:       // constants[i] - wires[i]
                for (u32 k = 0; k < g.param; k++)
                    gate_c[k] = add2(gate_c[k], mul2(filter, sub2(ext_at(gconst, k), ext_at(wires, k))));
                break;
            case SV_GATE_PUBLIC_INPUT:   // wires[0..4] - public_inputs_hash
                for (u32 k = 0; k < 4; k++)
                    gate_c[k] = add2(gate_c[k], mul2(filter, sub2(ext_at(wires, k), lift(pi_hash[k]))));
                break;
            case SV_GATE_ARITHMETIC: {   // output - (const_0 * m0 * m1 + const_1 * addend)
                const fp2 c0 = ext_at(gconst, 0), c1 = ext_at(gconst, 1);
                for (u32 k = 0; k < g.param; k++) {
                    const fp2 m0 = ext_at(wires, 4 * k), m1 = ext_at(wires, 4 * k + 1), ad = ext_at(wires, 4 * k + 2),
                              out = ext_at(wires, 4 * k + 3);
                    const fp2 computed = add2(mul2(mul2(m0, m1), c0), mul2(ad, c1));
                    gate_c[k] = add2(gate_c[k], mul2(filter, sub2(out, computed)));
                }
                break;
            }
            case SV_GATE_ARITHMETIC_EXT: {   // output - (const_0 * m0 * m1 + const_1 * addend), in the extension algebra
                const fp2 c0 = ext_at(gconst, 0), c1 = ext_at(gconst, 1);
                const alg2 zero = alg_lift(mk2(0, 0));
                for (u32 k = 0; k < g.param; k++) {
                    const alg2 mul = alg_mul_add(alg_at(wires, 8 * k), alg_at(wires, 8 * k + 2), zero);
                    const alg2 computed = alg_scalar_mul_add(c1, alg_at(wires, 8 * k + 4), alg_scalar_mul_add(c0, mul, zero));
                    const alg2 d = alg_sub(alg_at(wires, 8 * k + 6), computed);
                    gate_c[2 * k] = add2(gate_c[2 * k], mul2(filter, d.a));
                    gate_c[2 * k + 1] = add2(gate_c[2 * k + 1], mul2(filter, d.b));
                }
                break;
            }
            case SV_GATE_MUL_EXT: {          // output - const_0 * m0 * m1
                const fp2 c0 = ext_at(gconst, 0);
                const alg2 zero = alg_lift(mk2(0, 0));
                for (u32 k = 0; k < g.param; k++) {
                    const alg2 mul = alg_mul_add(alg_at(wires, 6 * k), alg_at(wires, 6 * k + 2), zero);
                    const alg2 d = alg_sub(alg_at(wires, 6 * k + 4), alg_scalar_mul_add(c0, mul, zero));
                    gate_c[2 * k] = add2(gate_c[2 * k], mul2(filter, d.a));
                    gate_c[2 * k + 1] = add2(gate_c[2 * k + 1], mul2(filter, d.b));
                }
                break;
            }
            case SV_GATE_BASE_SUM: {         // sum_k limb_k 2^k - sum; then limb (limb - 1) per limb
                fp2 computed = mk2(0, 0);
                for (u32 k = g.param; k-- > 0;) computed = add2(add2(computed, computed), ext_at(wires, 1 + k));
                gate_c[0] = add2(gate_c[0], mul2(filter, sub2(computed, ext_at(wires, 0))));
                for (u32 k = 0; k < g.param; k++) {
                    const fp2 limb = ext_at(wires, 1 + k);
                    gate_c[1 + k] = add2(gate_c[1 + k], mul2(filter, sub2(mul2(limb, limb), limb)));
                }
                break;
            }
            case SV_GATE_REDUCING:           // acc_i = acc_{i-1} * alpha + coeff_i; the last accumulator is the output
            case SV_GATE_REDUCING_EXT: {
                const bool ext = g.kind == SV_GATE_REDUCING_EXT;
                const u32 nco = g.param, start_accs = 6 + (ext ? 2 * nco : nco);
                const alg2 alpha = alg_at(wires, 2);
                alg2 acc = alg_at(wires, 4);
                for (u32 k = 0; k < nco; k++) {
                    const alg2 coeff = ext ? alg_at(wires, 6 + 2 * k) : alg_lift(ext_at(wires, 6 + k));
                    const alg2 acc_k = alg_at(wires, k == nco - 1 ? 0 : start_accs + 2 * k);
                    const alg2 d = alg_sub(alg_mul_add(acc, alpha, coeff), acc_k);
                    gate_c[2 * k] = add2(gate_c[2 * k], mul2(filter, d.a));
                    gate_c[2 * k + 1] = add2(gate_c[2 * k + 1], mul2(filter, d.b));
                    acc = acc_k;
                }
                break;
            }
            case SV_GATE_RANDOM_ACCESS: {    // gates/random_access.rs:84-166
                const u32 bits = g.param, copies = g.param2, extra = g.param3, vec = 1u << bits;
                const u32 routed = (2 + vec) * copies + extra;
                u32 k = 0;
                auto emit = [&](fp2 c) { gate_c[k] = add2(gate_c[k], mul2(filter, c)); k++; };
                for (u32 copy = 0; copy < copies; copy++) {
                    const u32 base = (2 + vec) * copy;
                    fp2 items[64];
                    for (u32 i = 0; i < vec; i++) items[i] = ext_at(wires, base + 2 + i);
                    fp2 reconstructed = mk2(0, 0);
                    for (u32 i = 0; i < bits; i++) {
                        const fp2 b = ext_at(wires, routed + copy * bits + i);
                        emit(sub2(mul2(b, b), b));
                    }
                    for (u32 i = bits; i-- > 0;) reconstructed = add2(add2(reconstructed, reconstructed), ext_at(wires, routed + copy * bits + i));
                    emit(sub2(reconstructed, ext_at(wires, base)));
                    for (u32 i = 0, len = vec; i < bits; i++, len >>= 1) {   // select(b, y, x) = b (y - x) + x
                        const fp2 b = ext_at(wires, routed + copy * bits + i);
                        for (u32 j = 0; j < len / 2; j++) items[j] = add2(mul2(b, sub2(items[2 * j + 1], items[2 * j])), items[2 * j]);
                    }
                    emit(sub2(items[0], ext_at(wires, base + 1)));
                }
                for (u32 i = 0; i < extra; i++) emit(sub2(ext_at(gconst, i), ext_at(wires, (2 + vec) * copies + i)));
                break;
            }
            case SV_GATE_POSEIDON_MDS:       // outputs - MDS(inputs) over the extension algebra, gates/poseidon_mds.rs:35-125
                for (u32 row = 0; row < 12; row++) {
                    alg2 res = alg_lift(mk2(0, 0));
                    for (u32 i = 0; i < 12; i++)
                        res = alg_scalar_mul_add(lift(SVB_T(MDS_MATRIX_CIRC)[i]), alg_at(wires, 2 * ((i + row) % 12)), res);
                    res = alg_scalar_mul_add(lift(SVB_T(MDS_MATRIX_DIAG)[row]), alg_at(wires, 2 * row), res);
                    const alg2 d = alg_sub(alg_at(wires, 2 * (12 + row)), res);
                    gate_c[2 * row] = add2(gate_c[2 * row], mul2(filter, d.a));
                    gate_c[2 * row + 1] = add2(gate_c[2 * row + 1], mul2(filter, d.b));
                }
                break;
            case SV_GATE_POSEIDON:
                poseidon_gate_constraints(wires, filter, gate_c);
                break;
            default: break;              // SV_GATE_NOOP: no constraints
        }
    }
    for (u32 k = 0; k < C.num_gate_constraints; k++) push(gate_c[k]);

    // plonk_verifier_chip.rs:194-209: Z_H(zeta) * reduce_with_powers(quotient chunk, zeta^n)
    bool ok = true;
    for (u32 i = 0; i < nch; i++) {
        fp2 acc = mk2(0, 0);
        for (u32 j = qdf; j-- > 0;) acc = add2(mul2(acc, zeta_pow_deg), ext_at(quot, i * qdf + j));
        ok = ok && eq2(sum[i], mul2(z_h, acc));
    }
    return ok;
}

}  // namespace svb

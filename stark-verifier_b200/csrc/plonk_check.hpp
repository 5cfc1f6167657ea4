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
//          algebra they compute in: chip/goldilocks_extension_algebra_chip.rs:34-171.
// The remaining gates of the recursion gate set (gates/mod.rs:138-196: Poseidon, PoseidonMds, RandomAccess) are not
// implemented yet: a circuit that uses one of them is refused by sv_plonk_circuit_check (an error, never a silent
// accept).
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"

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
SVB_HD void plonk_gate_dims(u32 kind, u32 param, u32& wires, u32& constraints, u32& constants) {
    wires = constraints = constants = 0;
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
        default: break;
    }
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
        plonk_gate_dims(g.kind, g.param, nwires, ncons, nconst);
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

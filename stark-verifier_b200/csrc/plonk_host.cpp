// Host side of the plonk-level checks: circuit validation, the plonk challenges of the transcript, and
// the CPU twin of the per-proof check (plonk_check.hpp) -- the function the device kernel runs, on CPU threads.
#include "plonk_check.hpp"
#include "host_util.hpp"
#include "layout.hpp"

#include <cstdio>
#include <string>

using namespace svb;

namespace svb {
// the shape must describe the same openings as the circuit's common data (CommonData::fri_oracles)
int plonk_shape_matches(const sv_fri_shape& s, const sv_plonk_circuit& C) {
    const sv_plonk_common& c = C.common;
    if (s.degree_bits != C.degree_bits) return -20;
    if (s.oracle_num_polys[0] != c.num_constants + c.num_routed_wires || s.oracle_num_polys[1] != c.num_wires ||
        s.oracle_num_polys[2] != c.num_challenges * (1 + c.num_partial_products) ||
        s.oracle_num_polys[3] != c.num_challenges * c.quotient_degree_factor || s.num_zs != c.num_challenges)
        return -21;
    return 0;
}
}  // namespace svb

// CustomGateRef::from (gates/mod.rs:138-196) matches `gate.0.id().trim_end()` against literal strings; the same
// strings here, with the numbers read instead of hard-coded.
extern "C" int sv_plonk_gate_from_id(const char* id, sv_plonk_gate* out) {
    if (!id || !out) return -1;
    std::string s(id);
    while (!s.empty() && (s.back() == ' ' || s.back() == '\n' || s.back() == '\t')) s.pop_back();
    sv_plonk_gate g;
    memset(&g, 0, sizeof g);
    unsigned a = 0, b = 0, c = 0;
    int used = 0;
    auto whole = [&](int n) { return n > 0 && (size_t)n == s.size(); };
    if (s == "NoopGate") g.kind = SV_GATE_NOOP;
    else if (s == "PublicInputGate") g.kind = SV_GATE_PUBLIC_INPUT;
    else if (sscanf(s.c_str(), "ArithmeticGate { num_ops: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_ARITHMETIC; g.param = a; }
    else if (sscanf(s.c_str(), "ConstantGate { num_consts: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_CONSTANT; g.param = a; }
    else if (sscanf(s.c_str(), "BaseSumGate { num_limbs: %u } + Base: %u%n", &a, &b, &used) == 2 && whole(used) && b == 2) { g.kind = SV_GATE_BASE_SUM; g.param = a; }
    else if (sscanf(s.c_str(), "ArithmeticExtensionGate { num_ops: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_ARITHMETIC_EXT; g.param = a; }
    else if (sscanf(s.c_str(), "MulExtensionGate { num_ops: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_MUL_EXT; g.param = a; }
    else if (sscanf(s.c_str(), "ReducingGate { num_coeffs: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_REDUCING; g.param = a; }
    else if (sscanf(s.c_str(), "ReducingExtensionGate { num_coeffs: %u }%n", &a, &used) == 1 && whole(used)) { g.kind = SV_GATE_REDUCING_EXT; g.param = a; }
    else if (sscanf(s.c_str(), "RandomAccessGate { bits: %u, num_copies: %u, num_extra_constants: %u, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>%n",
                    &a, &b, &c, &used) == 3 && whole(used)) { g.kind = SV_GATE_RANDOM_ACCESS; g.param = a; g.param2 = b; g.param3 = c; }
    else if (s == "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>") g.kind = SV_GATE_POSEIDON;
    else if (s == "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>") g.kind = SV_GATE_POSEIDON_MDS;
    else return -2;
    *out = g;
    return 0;
}

// CommonData::from (types/common_data.rs:224-270): field-by-field, with CustomGateRef::from (gates/mod.rs:138-196) per gate.
extern "C" int sv_circuit_from_common_data(const sv_common_circuit_data* cd, sv_fri_shape* shape_out, sv_plonk_circuit* circuit_out) {
    if (!cd || !shape_out || !circuit_out) return -1;
    if (cd->num_gates == 0 || cd->num_gates > SV_MAX_GATES || !cd->gate_ids || !cd->selector_indices) return -2;
    if (cd->num_selector_groups == 0 || cd->num_selector_groups > SV_MAX_SELECTORS || !cd->group_starts || !cd->group_ends) return -2;
    if (cd->num_k_is != cd->common.num_routed_wires || cd->num_k_is > SV_MAX_ROUTED_WIRES || !cd->k_is) return -2;
    if (cd->num_reduction_steps && !cd->reduction_arity_bits) return -2;
    sv_fri_shape s;
    if (int rc = sv_fri_shape_from_common(&cd->common, cd->degree_bits, cd->rate_bits, cd->cap_height, cd->num_query_rounds,
                                          cd->proof_of_work_bits, cd->num_reduction_steps, cd->reduction_arity_bits, cd->hiding, cd->hash_kind, &s))
        return rc < 0 ? rc - 10 : rc;
    sv_plonk_circuit C;
    memset(&C, 0, sizeof C);
    C.common = cd->common;
    C.degree_bits = cd->degree_bits;
    C.num_gate_constraints = cd->num_gate_constraints;
    C.num_selectors = cd->num_selector_groups;
    for (u32 g = 0; g < cd->num_selector_groups; g++) { C.group_lo[g] = cd->group_starts[g]; C.group_hi[g] = cd->group_ends[g]; }
    C.num_gates = cd->num_gates;
    for (u32 i = 0; i < cd->num_gates; i++) {
        if (sv_plonk_gate_from_id(cd->gate_ids[i], &C.gates[i])) return -3;          // unimplemented!() in the reference
        const u32 sel = cd->selector_indices[i];
        if (sel >= cd->num_selector_groups || i < cd->group_starts[sel] || i >= cd->group_ends[sel]) return -4;   // SelectorsInfo invariant
        C.gates[i].selector_index = sel;
    }
    for (u32 j = 0; j < cd->num_k_is; j++) {
        if (!is_canonical(cd->k_is[j])) return -5;
        C.k_is[j] = cd->k_is[j];
    }
    if (int rc = plonk_circuit_check(C)) return rc < 0 ? rc - 20 : rc;
    if (int rc = plonk_shape_matches(s, C)) return rc < 0 ? rc - 30 : rc;
    *shape_out = s;
    *circuit_out = C;
    return 0;
}

extern "C" int sv_plonk_circuit_check(const sv_plonk_circuit* circuit) {
    if (!circuit) return -1;
    return plonk_circuit_check(*circuit);
}

// Same duplex sponge as Challenger in host_side.cpp, up to the plonk alphas (plonk_verifier_chip.rs:65-103).
extern "C" int sv_plonk_challenges(const sv_fri_shape* shape, const uint64_t* rec, const uint64_t circuit_digest[4],
                                   const uint64_t pi_hash[4], uint32_t nc, uint64_t* out) {
    sv_fri_layout L;
    if (!shape || !rec || !circuit_digest || !pi_hash || !out || make_layout(*shape, L)) return -1;
    if (shape->hash_kind > SV_HASH_POSEIDON_BN254 || nc == 0 || nc > SV_MAX_PLONK_CHALLENGES) return -2;
    const u32 kind = shape->hash_kind;
    u64 st[12] = {0}, in[8], outb[8];
    int n_in = 0, n_out = 0;
    auto duplex = [&](int len) {
        for (int i = 0; i < len; i++) st[i] = in[i];
        permute_kind(kind, st);
        memcpy(outb, st, 64);
        n_out = 8;
        n_in = 0;
    };
    auto observe = [&](u64 v) {
        n_out = 0;
        in[n_in++] = v;
        if (n_in == 8) duplex(8);
    };
    auto squeeze = [&]() {
        if (n_in) duplex(n_in);
        if (n_out == 0) {
            permute_kind(kind, st);
            memcpy(outb, st, 64);
            n_out = 8;
        }
        return outb[--n_out];
    };
    const u32 capw = L.ncap * 4;
    const u64* caps = rec + L.off_init_caps;
    for (int i = 0; i < 4; i++) observe(circuit_digest[i]);
    for (int i = 0; i < 4; i++) observe(pi_hash[i]);
    for (u32 i = 0; i < capw; i++) observe(caps[1 * capw + i]);         // wires_cap
    for (u32 i = 0; i < nc; i++) out[i] = squeeze();                    // plonk_betas
    for (u32 i = 0; i < nc; i++) out[nc + i] = squeeze();               // plonk_gammas
    for (u32 i = 0; i < capw; i++) observe(caps[2 * capw + i]);         // plonk_zs_partial_products_cap
    for (u32 i = 0; i < nc; i++) out[2 * nc + i] = squeeze();           // plonk_alphas
    return 0;
}

extern "C" int sv_plonk_check_host(const sv_fri_shape* shape, const sv_plonk_circuit* C, size_t n, const uint64_t* records,
                                   const uint64_t* pi_hashes, const uint64_t* chal, uint32_t* bitmap, int nthreads) {
    sv_fri_layout L;
    if (!shape || !C || !bitmap || (n && (!records || !pi_hashes || !chal)) || make_layout(*shape, L)) return -1;
    if (int rc = plonk_circuit_check(*C)) return rc;
    if (int rc = plonk_shape_matches(*shape, *C)) return rc;
    const size_t words = (n + 31) / 32;
    const u32 nch = C->common.num_challenges;
    if (nthreads < 1) nthreads = 1;
    parallel_for(words, nthreads, [&](size_t b, size_t e) {
        for (size_t w = b; w < e; w++) {
            u32 bits = 0;
            for (size_t p = 32 * w; p < n && p < 32 * w + 32; p++) {
                const u64* rec = records + p * (size_t)L.record_words;
                bool ok = plonk_check_one(*C, rec + L.off_open0, rec + L.off_open1, pi_hashes + 4 * p, chal + 3 * (size_t)nch * p,
                                          mk2(rec[L.off_zeta], rec[L.off_zeta + 1]));
                bits |= (u32)ok << (p & 31);
            }
            bitmap[w] = bits;
        }
    });
    return 0;
}

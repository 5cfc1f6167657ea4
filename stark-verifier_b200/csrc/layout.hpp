// Flat per-proof record layout (see include/stark_verifier_b200.h for the field list).
// Mirrors the data contract of the reference's FriProofValues / FriQueryRoundValues /
// FriInitialTreeProofValues / FriQueryStepValues (types/proof.rs:143-377) and FriChallenges /
// FriOpenings (types/assigned.rs:117-135), flattened in declaration order.
#pragma once
#include "../../include/stark_verifier_b200.h"
#include <string.h>

#if defined(__CUDACC__)
#define SVB_LAYOUT_HD __host__ __device__ static inline
#else
#define SVB_LAYOUT_HD static inline
#endif

namespace svb {

SVB_LAYOUT_HD uint32_t up4(uint32_t x) { return (x + 3u) & ~3u; }

static inline int make_layout(const sv_fri_shape& s, sv_fri_layout& L) {
    memset(&L, 0, sizeof L);
    if (s.num_steps > SV_MAX_STEPS || s.cap_height > 16 || s.num_query_rounds == 0) return -1;
    // Goldilocks has 2-adicity 32: there is no domain of more than 2^32 points
    if (s.degree_bits > 32 || s.rate_bits > 32 || s.degree_bits + s.rate_bits > 32) return -1;
    L.ncap = 1u << s.cap_height;
    L.lde_bits = s.degree_bits + s.rate_bits;
    uint32_t total_arity_bits = 0;
    for (uint32_t i = 0; i < s.num_steps; i++) {
        if (s.reduction_arity_bits[i] < 1 || s.reduction_arity_bits[i] > SV_MAX_ARITY_BITS) return -1;
        total_arity_bits += s.reduction_arity_bits[i];
    }
    if (L.lde_bits < s.cap_height + total_arity_bits) return -1;   // every step tree is at least as tall as the cap
    if (s.num_zs > s.oracle_num_polys[2]) return -1;
    L.n0 = s.oracle_num_polys[0] + s.oracle_num_polys[1] + s.oracle_num_polys[2] + s.oracle_num_polys[3];
    L.n1 = s.num_zs;
    uint32_t o = 0;
    auto seg = [&o](uint32_t words) { uint32_t at = o; o = up4(o + words); return at; };
    L.off_init_caps = seg(4 * L.ncap * 4);
    L.off_step_caps = seg(s.num_steps * L.ncap * 4);
    L.off_open0 = seg(2 * L.n0);
    L.off_open1 = seg(2 * L.n1);
    L.off_final_poly = seg(2 * s.final_poly_len);
    L.off_pow_witness = seg(1);
    L.off_alpha = seg(2);
    L.off_betas = seg(2 * s.num_steps);
    L.off_pow_response = seg(1);
    L.off_indices = seg(s.num_query_rounds);
    L.off_zeta = seg(2);
    L.off_zeta_next = seg(2);
    L.header_words = o;
    o = 0;
    L.init_depth = L.lde_bits - s.cap_height;
    uint32_t algo_q = 0, perms = 0;
    for (int k = 0; k < 4; k++) {
        L.leaf_len[k] = s.oracle_num_polys[k] + ((s.hiding && s.oracle_blinding[k]) ? 4u : 0u);
        L.q_off_init_evals[k] = seg(L.leaf_len[k]);
        L.q_off_init_sibs[k] = seg(4 * L.init_depth);
        algo_q += 8 * L.leaf_len[k] + 32 * L.init_depth;
        perms += (L.leaf_len[k] > 4 ? (L.leaf_len[k] + 7) / 8 : 0) + L.init_depth;
    }
    uint32_t shift = 0;
    for (uint32_t i = 0; i < s.num_steps; i++) {
        const uint32_t ab = s.reduction_arity_bits[i], leaf = 2u << ab;   // 2^ab Fp2 evaluations = 2^(ab+1) words
        shift += ab;
        L.step_arity_bits[i] = ab;
        L.step_index_shift[i] = shift;
        L.step_depth[i] = L.lde_bits - shift - s.cap_height;
        L.q_off_step_evals[i] = seg(leaf);
        L.q_off_step_sibs[i] = seg(4 * L.step_depth[i]);
        algo_q += 8 * leaf + 32 * L.step_depth[i];
        perms += (leaf > 4 ? (leaf + 7) / 8 : 0) + L.step_depth[i];
    }
    L.query_words = o;
    L.record_words = L.header_words + s.num_query_rounds * L.query_words;
    L.algo_bytes_per_query = algo_q;
    // caps + openings + final poly + pow witness (SURVEY 8d "shared" bytes)
    L.algo_bytes_shared = 32 * L.ncap * (4 + s.num_steps) + 16 * (L.n0 + L.n1) + 16 * s.final_poly_len + 8;
    L.perms_per_query = perms;
    return 0;
}

}  // namespace svb
